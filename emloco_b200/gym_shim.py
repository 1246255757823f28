"""`gymapi` / `gymtorch`-shaped shim over the C ABI: the subset of the Isaac Gym tensor API that the PACER tasks call
(SURVEY 8b): isaacgym/python/isaacgym/gymtorch.py:61-106 and the call sites pacer/pacer/env/tasks/base_task.py:59,128,238,
258,792-797, humanoid.py:137-216,470-475,1202, humanoid_amp.py:565-583.  Asset loading / env creation are replaced by the
static MJCF table (`emloco_b200.model`), so `create_sim` takes the env count directly.

    gym = acquire_gym()
    sim = gym.create_sim(compute_device_id=0, num_envs=4096, sim_params=SimParams())
    root = wrap_tensor(gym.acquire_actor_root_state_tensor(sim))          # zero-copy alias, stable address
    gym.set_dof_position_target_tensor(sim, unwrap_tensor(targets)); gym.simulate(sim); gym.fetch_results(sim, True)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib
from .sim import EmlocoSim, _stream


@dataclass
class PhysxParams:                      # config.py:153-160, pacer.yaml:93-104 (recorded; the contact model is the builder's)
    solver_type: int = 1
    num_position_iterations: int = 4
    num_velocity_iterations: int = 0
    contact_offset: float = 0.02
    rest_offset: float = 0.0
    bounce_threshold_velocity: float = 0.2
    max_depenetration_velocity: float = 10.0


@dataclass
class SimParams:                        # gymapi.SimParams as filled by parse_sim_params (utils/config.py:141-174)
    dt: float = 1.0 / 60.0
    substeps: int = 2
    gravity_z: float = -9.81
    physx: PhysxParams = field(default_factory=PhysxParams)


class Tensor:
    """gymapi.Tensor descriptor: device ordinal, dtype code, shape, data_address, own_data (gymtorch.py:61-88)."""

    def __init__(self, device, dtype, shape, data_address, owner):
        self.device, self.dtype, self.shape, self.data_address, self.own_data, self._owner = device, dtype, list(shape), int(data_address), False, owner


def wrap_tensor(t: Tensor):
    from .sim import wrap_tensor as _wrap
    return _wrap(t.data_address, t.shape, t.dtype, t.device, owner=t._owner)


def unwrap_tensor(t: torch.Tensor):
    if not t.is_contiguous():
        raise Exception("Input tensor must be contiguous")          # gymtorch.py:91-92
    if not t.is_cuda:
        raise _lib.EmlocoError("emloco_b200 works on CUDA tensors only (pipeline=gpu); there is no CPU pipeline")
    return t


class Gym:
    def create_sim(self, compute_device_id=0, graphics_device_id=-1, engine=None, sim_params: SimParams = None, num_envs=1, **cfg):
        p = sim_params or SimParams()
        return EmlocoSim(num_envs, device=compute_device_id, sim_dt=p.dt, substeps=p.substeps, gravity_z=p.gravity_z,
                         contact_offset=p.physx.contact_offset, **cfg)

    def prepare_sim(self, sim):                                       # base_task.py:128
        return True

    def _acquire(self, sim, name):
        t = sim.tensor(name)
        code = {torch.float32: 0, torch.int64: 1, torch.int16: 2}[t.dtype]
        return Tensor(sim.device, code, t.shape, t.data_ptr(), sim)

    def acquire_actor_root_state_tensor(self, sim): return self._acquire(sim, "root_state")      # humanoid.py:137
    def acquire_dof_state_tensor(self, sim): return self._acquire(sim, "dof_state")              # :138
    def acquire_rigid_body_state_tensor(self, sim): return self._acquire(sim, "rb_state")        # :140
    def acquire_net_contact_force_tensor(self, sim): return self._acquire(sim, "contact")        # :141
    def acquire_dof_force_tensor(self, sim): return self._acquire(sim, "dof_force")              # :150
    def acquire_force_sensor_tensor(self, sim):                                                   # :139 (unused by the default task)
        raise NotImplementedError("force sensors are not on the hot path (SURVEY 8a)")

    # state tensors are written by the step kernel itself: refresh is a no-op (humanoid_amp.py:565-583)
    def refresh_dof_state_tensor(self, sim): return True
    refresh_actor_root_state_tensor = refresh_rigid_body_state_tensor = refresh_net_contact_force_tensor = \
        refresh_dof_force_tensor = refresh_force_sensor_tensor = refresh_dof_state_tensor

    def set_dof_position_target_tensor(self, sim, targets):                                       # humanoid.py:1202
        sim.set_pd_targets(unwrap_tensor(targets))
        return True

    def set_dof_actuation_force_tensor(self, sim, forces):                                        # :1207 (DOF_MODE_EFFORT)
        raise NotImplementedError("torque control is not the configured drive mode (pd_control, humanoid.py:905-910)")

    def simulate(self, sim):                                                                       # base_task.py:795
        sim.simulate()

    def fetch_results(self, sim, wait=True):                                                       # base_task.py:258
        if wait:
            torch.cuda.current_stream().synchronize()

    def _indexed(self, sim, ids, n):
        ids = unwrap_tensor(ids)
        if ids.dtype != torch.int32:
            raise _lib.EmlocoError("actor index tensors are int32 (humanoid.py:469)")
        sim.reset_indexed(ids[:n])
        return True

    def set_actor_root_state_tensor_indexed(self, sim, root_states, actor_ids, n):                 # humanoid.py:470-472
        return self._indexed(sim, actor_ids, n)

    def set_dof_state_tensor_indexed(self, sim, dof_states, actor_ids, n):                         # :473-475
        return self._indexed(sim, actor_ids, n)      # idempotent: re-reads both aliases, like the first call did

    def get_sim_params(self, sim):
        return SimParams(dt=sim.cfg.sim_dt, substeps=sim.cfg.substeps, gravity_z=sim.cfg.gravity_z)

    def get_asset_rigid_body_count(self, asset=None): return _lib.NB
    def get_asset_dof_count(self, asset=None): return _lib.ND
    def get_asset_joint_count(self, asset=None): return _lib.ND

    def find_actor_rigid_body_handle(self, sim, env=None, actor=None, name=None):                 # humanoid.py:917-944
        return sim.model_arrays["names"].index(name)

    def destroy_sim(self, sim):
        sim.close()


def acquire_gym():                                                                                 # base_task.py:59
    return Gym()
