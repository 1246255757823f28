"""`gymapi` / `gymtorch`-shaped shim over the C ABI: the subset of the Isaac Gym tensor API that the PACER tasks call
(SURVEY 8b): isaacgym/python/isaacgym/gymtorch.py:61-106 and the call sites pacer/pacer/env/tasks/base_task.py:59,128,238,
258,792-797, humanoid.py:137-216,470-475,1202, humanoid_amp.py:565-583, and the env-creation sequence of
humanoid.py:643-946 / humanoid_pedestrain_terrain.py:860-880 (load_asset, create_env, create_actor, *_properties,
add_ground / add_triangle_mesh): with `num_envs=None` the sim is built by `prepare_sim` from what those calls collected.

    gym = acquire_gym()
    sim = gym.create_sim(compute_device_id=0, num_envs=4096, sim_params=SimParams())     # or the reference's creation calls
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


# ---- small value types of gymapi that the creation code fills in (humanoid.py:643-946) -----------------------
DOF_MODE_NONE, DOF_MODE_POS, DOF_MODE_VEL, DOF_MODE_EFFORT = 0, 1, 2, 3
MESH_VISUAL, MESH_COLLISION, MESH_VISUAL_AND_COLLISION = 1, 2, 3
SIM_PHYSX = 1


class _Bag:
    """Attribute bag: gymapi option / parameter structs are filled attribute by attribute by the task code."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class Vec3(_Bag):
    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(x=float(x), y=float(y), z=float(z))


class Quat(_Bag):
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        super().__init__(x=float(x), y=float(y), z=float(z), w=float(w))


class Transform(_Bag):
    def __init__(self, p=None, r=None):
        super().__init__(p=p or Vec3(), r=r or Quat())


class AssetOptions(_Bag):
    def __init__(self):
        super().__init__(angular_damping=0.01, max_angular_velocity=100.0, default_dof_drive_mode=DOF_MODE_NONE, fix_base_link=False)


class PlaneParams(_Bag):
    def __init__(self):
        super().__init__(normal=Vec3(0, 0, 1), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


class TriangleMeshParams(_Bag):
    def __init__(self):
        super().__init__(nb_vertices=0, nb_triangles=0, transform=Transform(), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


class Asset:
    """gym.load_asset result: the articulation parsed from the MJCF file (emloco_b200.mjcf), shared by every actor made from it."""

    def __init__(self, model, options):
        self.model, self.options, self.force_sensors = model, options, []


class PendingSim:
    """What create_sim returns when the env count is not known yet: the creation calls are recorded and `prepare_sim` builds
    the device-side sim from them.  The handle stays valid afterwards (every Gym method resolves it)."""

    def __init__(self, device, params, cfg):
        self.device, self.params, self.cfg = device, params, cfg
        self.envs, self.height, self.real, self.friction = [], None, None, None


class _Env:
    def __init__(self, sim, index):
        self.sim, self.index, self.actors = sim, index, []


class _Actor:
    def __init__(self, asset, pose, name, group, filt):
        self.asset, self.pose, self.name, self.group, self.filter = asset, pose, name, group, filt
        self.dof_props, self.shape_props, self.force_sensors = None, None, False


def _real(sim):
    if isinstance(sim, PendingSim):
        if sim.real is None:
            raise _lib.EmlocoError("gym.prepare_sim(sim) has not been called yet (base_task.py:128)")
        return sim.real
    return sim


class Gym:
    def create_sim(self, compute_device_id=0, graphics_device_id=-1, engine=None, sim_params: SimParams = None, num_envs=None, **cfg):
        p = sim_params or SimParams()
        if num_envs is None:                                          # the reference's flow: envs are created one by one
            return PendingSim(compute_device_id, p, cfg)
        return EmlocoSim(num_envs, device=compute_device_id, sim_dt=p.dt, substeps=p.substeps, gravity_z=p.gravity_z,
                         contact_offset=p.physx.contact_offset, **cfg)

    # ---- ground (base_task / humanoid.py:_create_ground_plane, humanoid_pedestrain_terrain.py:860-880) ----
    def add_ground(self, sim, plane_params):
        sim.height, sim.friction = None, plane_params.static_friction          # flat default height field

    def add_triangle_mesh(self, sim, vertices, triangles, tm_params, vertical_scale=0.005):
        """The terrain tri-mesh of Terrain.convert_heightfield_to_trimesh (..terrain.py:1300-1340 via :866-877) is a regular
        grid: vertex (i, j) at (i*hs, j*hs, height*vs), row-major in i.  The height field is recovered from it; the
        border shift of tm_params.transform is the task's own (it adds it back when it samples, ..terrain.py:1212-1218)."""
        import numpy as np
        v = np.asarray(vertices, np.float32).reshape(-1, 3)
        cols = int(np.argmax(v[:, 0] != v[0, 0])) or v.shape[0]
        if v.shape[0] % cols:
            raise _lib.EmlocoError("add_triangle_mesh: vertices are not a regular height-field grid")
        sim.height = np.rint(v[:, 2] / vertical_scale).astype(np.int16).reshape(v.shape[0] // cols, cols)
        sim.friction = tm_params.static_friction

    # ---- assets (humanoid.py:744-800) ----
    def load_asset(self, sim, asset_root, asset_file, options=None):
        import os
        from .mjcf import default_model, load_mjcf
        path = os.path.join(asset_root, asset_file)
        if os.path.exists(path):
            return Asset(load_mjcf(path), options or AssetOptions())
        if os.path.basename(asset_file).startswith("smpl_humanoid"):
            # the MJCF itself is not redistributed with this package: the table parsed from it is (assets/*.json)
            return Asset(default_model(), options or AssetOptions())
        raise _lib.EmlocoError(f"load_asset: {path} not found")

    def get_asset_actuator_properties(self, asset):
        eff = asset.model.extra.get("motor_effort", [500.0] * asset.model.num_dof)      # MJCF motor gear
        return [_Bag(motor_effort=float(e)) for e in eff]

    def find_asset_rigid_body_index(self, asset, name):
        return asset.model.names.index(name) if name in asset.model.names else -1

    def create_asset_force_sensor(self, asset, body_idx, local_pose, props=None):
        asset.force_sensors.append(body_idx)
        return len(asset.force_sensors) - 1

    def get_asset_dof_properties(self, asset):
        import numpy as np
        m = asset.model
        dt = np.dtype([("hasLimits", "?"), ("lower", "f4"), ("upper", "f4"), ("driveMode", "i4"), ("velocity", "f4"), ("effort", "f4"),
                       ("stiffness", "f4"), ("damping", "f4"), ("friction", "f4"), ("armature", "f4")])
        a = np.zeros(m.num_dof, dt)
        a["hasLimits"], a["lower"], a["upper"] = True, m.limit_lo, m.limit_hi
        a["driveMode"], a["velocity"], a["effort"] = asset.options.default_dof_drive_mode, 3.4e38, 500.0
        a["stiffness"], a["damping"], a["armature"] = m.kp, m.kd, m.armature
        return a

    # ---- envs and actors (humanoid.py:802-946) ----
    def create_env(self, sim, lower, upper, num_per_row):
        e = _Env(sim, len(sim.envs))
        sim.envs.append(e)
        return e

    def create_actor(self, env, asset, pose, name="", group=-1, filter=-1, segmentation_id=0):
        if env.actors:
            raise _lib.EmlocoError("one humanoid actor per env (headless PACER task); markers / objects are not simulated")
        env.actors.append(_Actor(asset, pose, name, group, filter))
        return 0

    def enable_actor_dof_force_sensors(self, env, actor):
        env.actors[actor].force_sensors = True                       # the DOF-force tensor is always produced

    def get_actor_rigid_body_properties(self, env, actor):
        m = env.actors[actor].asset.model
        return [_Bag(mass=float(x), com=Vec3(*c)) for x, c in zip(m.mass, m.com)]

    def get_actor_dof_properties(self, env, actor):
        a = env.actors[actor]
        return a.dof_props if a.dof_props is not None else self.get_asset_dof_properties(a.asset)

    def set_actor_dof_properties(self, env, actor, props):
        env.actors[actor].dof_props = props.copy()
        return True

    def get_actor_rigid_shape_properties(self, env, actor):
        a = env.actors[actor]
        if a.shape_props is None:
            a.shape_props = [_Bag(filter=0, friction=1.0, rolling_friction=0.0, torsion_friction=0.0, restitution=0.0, compliance=0.0,
                                  thickness=0.0) for _ in range(a.asset.model.num_bodies)]
        return a.shape_props

    def set_actor_rigid_shape_properties(self, env, actor, props):
        env.actors[actor].shape_props = list(props)                  # self-collision filters: recorded (no self-collision in the kernel)
        return True

    def set_rigid_body_color(self, *a, **k):
        return None                                                  # viewer only

    def get_actor_rigid_body_count(self, env, actor): return env.actors[actor].asset.model.num_bodies
    def get_actor_dof_count(self, env, actor): return env.actors[actor].asset.model.num_dof

    def prepare_sim(self, sim):                                       # base_task.py:128
        """Builds the device-side sim from the recorded creation calls.  All envs must share one articulation and one set
        of drive gains (per-env shape variation is SURVEY 8 row f3)."""
        if not isinstance(sim, PendingSim):
            return True
        import numpy as np
        from .model import build_model_arrays
        if not sim.envs or any(len(e.actors) != 1 for e in sim.envs):
            raise _lib.EmlocoError("prepare_sim: every env needs exactly one actor")
        a0 = sim.envs[0].actors[0]
        props0 = a0.dof_props if a0.dof_props is not None else self.get_asset_dof_properties(a0.asset)
        for e in sim.envs[1:]:
            a = e.actors[0]
            if a.asset.model is not a0.asset.model:
                raise _lib.EmlocoError("prepare_sim: per-env assets (shape variation) are not supported yet (SURVEY 8 f3)")
            if a.dof_props is not None and not (np.array_equal(a.dof_props["stiffness"], props0["stiffness"])
                                                and np.array_equal(a.dof_props["damping"], props0["damping"])):
                raise _lib.EmlocoError("prepare_sim: per-env drive gains are not supported yet (SURVEY 8 f3)")
        if int(props0["driveMode"][0]) != DOF_MODE_POS:
            raise _lib.EmlocoError("prepare_sim: only DOF_MODE_POS (pd_control) is implemented (humanoid.py:905-910)")
        arrs = build_model_arrays(a0.asset.model, scale_by_mass=False)
        per_joint = lambda x: np.concatenate([[0.0], np.asarray(x, np.float64)[0::3]])
        arrs.update(kp=np.asarray(props0["stiffness"], np.float64), kd=np.asarray(props0["damping"], np.float64),
                    kp_joint=per_joint(props0["stiffness"]), kd_joint=per_joint(props0["damping"]))
        p = sim.params
        cfg = dict(sim.cfg)
        if sim.friction is not None:
            cfg.setdefault("friction_mu", float(sim.friction))
        cfg.setdefault("max_ang_vel", float(a0.asset.options.max_angular_velocity))
        cfg.setdefault("angular_damping", float(a0.asset.options.angular_damping))
        real = EmlocoSim(len(sim.envs), device=sim.device, model_arrays=arrs, sim_dt=p.dt, substeps=p.substeps, gravity_z=p.gravity_z,
                         contact_offset=p.physx.contact_offset, **cfg)
        if sim.height is not None:
            real.set_height_field(sim.height)
        # start poses of create_actor (humanoid.py:849-865) -> root state, then forward kinematics
        root = torch.zeros(len(sim.envs), 13)
        for i, e in enumerate(sim.envs):
            t = e.actors[0].pose
            root[i, 0:7] = torch.tensor([t.p.x, t.p.y, t.p.z, t.r.x, t.r.y, t.r.z, t.r.w])
        real.root_state.copy_(root.to(real.root_state.device))
        real.reset_indexed(None)
        sim.real = real
        return True

    def get_frame_count(self, sim):
        return int(getattr(_real(sim), "_frames", 0))

    def _acquire(self, sim, name):
        sim = _real(sim)
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
        _real(sim).set_pd_targets(unwrap_tensor(targets))
        return True

    def set_dof_actuation_force_tensor(self, sim, forces):                                        # :1207 (DOF_MODE_EFFORT)
        raise NotImplementedError("torque control is not the configured drive mode (pd_control, humanoid.py:905-910)")

    def simulate(self, sim):                                                                       # base_task.py:795
        sim = _real(sim)
        sim.simulate()
        sim._frames = getattr(sim, "_frames", 0) + 1

    def fetch_results(self, sim, wait=True):                                                       # base_task.py:258
        if wait:
            torch.cuda.current_stream().synchronize()

    def _indexed(self, sim, ids, n):
        sim = _real(sim)
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
        if isinstance(sim, PendingSim) and sim.real is None:
            return sim.params
        sim = _real(sim)
        return SimParams(dt=sim.cfg.sim_dt, substeps=sim.cfg.substeps, gravity_z=sim.cfg.gravity_z)

    def get_asset_rigid_body_count(self, asset=None): return _lib.NB if asset is None else asset.model.num_bodies
    def get_asset_dof_count(self, asset=None): return _lib.ND if asset is None else asset.model.num_dof
    def get_asset_joint_count(self, asset=None): return _lib.ND if asset is None else asset.model.num_dof

    def find_actor_rigid_body_handle(self, sim_or_env, env=None, actor=None, name=None):          # humanoid.py:917-944
        """gym.find_actor_rigid_body_handle(env_ptr, actor_handle, body_name) - or (sim, name=...) on a built sim."""
        if isinstance(sim_or_env, _Env):
            names, name = sim_or_env.actors[env or 0].asset.model.names, (actor if name is None else name)
        else:
            names = _real(sim_or_env).model_arrays["names"]
        return names.index(name) if name in names else -1

    def destroy_sim(self, sim):
        _real(sim).close()


def acquire_gym():                                                                                 # base_task.py:59
    return Gym()
