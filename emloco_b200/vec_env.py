"""The vec-env surface rl_games sees (reference: `RLGPUEnv`, pacer/pacer/run.py:135-182; `VecTaskPython.step`,
env/tasks/vec_task.py:125-134; `VecTaskPythonWrapper`, env/tasks/vec_task_wrappers.py:27-72) over `EmlocoSim`.

`step(actions) -> (obs, rewards, dones, infos)`, `reset(env_ids=None) -> obs`, `get_number_of_agents()`,
`get_env_info()`, and the LocoVal getters `get_waypoint_traj / get_init_pose / get_init_vel`.  All tensors stay on the
sim device and alias sim memory (no copies), except when `rl_device` differs - then obs / rewards / dones are moved like
`.to(self.rl_device)` in the reference.  Registration names of run.py:185-197 ("RLGPU" / "rlgpu") are kept as constants.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .sim import EmlocoSim
from .synthetic import synthetic_env_state

VECENV_NAME, ENV_CONFIG_NAME = "RLGPU", "rlgpu"          # run.py:185-197
NUM_OBS, NUM_ACTIONS, NUM_AMP_OBS = 1422, 69, 3090


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency of the hot path): shape / low / high."""

    def __init__(self, low, high):
        self.low, self.high = np.asarray(low, np.float32), np.asarray(high, np.float32)
        self.shape = self.low.shape


class RLGPUEnv:
    def __init__(self, num_envs, device=0, rl_device=None, clip_observations=np.inf, seed=0, init_state=None, **cfg_over):
        self.sim = EmlocoSim(num_envs, device=device, **cfg_over)
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", device)
        self.rl_device = torch.device(rl_device) if rl_device is not None else self.device
        self.clip_obs = float(clip_observations)             # parse_task.py:45: np.inf by default
        st = init_state or synthetic_env_state(self.num_envs, seed=seed, root_height=self.sim.rest_height)
        self.init_root = torch.from_numpy(st["root"]).to(self.device)
        self.init_dof = torch.from_numpy(st["dof"]).to(self.device)
        self.sim.traj_verts.copy_(torch.from_numpy(st["verts"]).to(self.device))
        self._waypoints = torch.from_numpy(st["waypoints"]).to(self.device)
        self.observation_space = Box(np.full(NUM_OBS, -np.inf), np.full(NUM_OBS, np.inf))
        self.action_space = Box(-np.ones(NUM_ACTIONS), np.ones(NUM_ACTIONS))
        self.amp_observation_space = Box(np.full(NUM_AMP_OBS, -np.inf), np.full(NUM_AMP_OBS, np.inf))
        self.num_states = 0
        self.sim.reset.fill_(1)
        self.sim.reset_done(self.init_root, self.init_dof)
        rb = self.sim.rb_state.view(self.num_envs, 24, 13)
        self._init_pose = rb[:, :, 0:3].clone()
        self._init_vel = self.init_root[:, 7:9].clone()
        self.full_state = {"obs": self._obs()}

    def _obs(self):
        o = self.sim.obs
        if np.isfinite(self.clip_obs):
            o = torch.clamp(o, -self.clip_obs, self.clip_obs)
        return o.to(self.rl_device)

    # ---- run.py:148-160 / vec_task.py:125-134 ----
    def step(self, actions):
        a = actions.to(self.device, torch.float32)
        if not a.is_contiguous():
            raise _lib.EmlocoError("actions must be contiguous (gymtorch.unwrap_tensor rule, gymtorch.py:89-106)")
        self.sim.step(a)
        infos = {"amp_obs": self.sim.amp_obs.view(self.num_envs, NUM_AMP_OBS), "terminate": self.sim.terminate,
                 "reward_raw": self.sim.rew_raw, "flip_obs": self.sim.flip_obs}
        self.full_state["obs"] = self._obs()
        return self.full_state["obs"], self.sim.rew.to(self.rl_device), self.sim.reset.to(self.rl_device), infos

    # ---- run.py:162-168 / vec_task_wrappers.py:36-39 ----
    def reset(self, env_ids=None):
        if env_ids is None:
            self.sim.reset.fill_(1)
        elif len(env_ids) > 0:
            self.sim.reset.zero_()
            self.sim.reset[env_ids.to(self.device, torch.long)] = 1
        else:
            return self.full_state["obs"]
        self.sim.reset_done(self.init_root, self.init_dof)
        self.full_state["obs"] = self._obs()
        return self.full_state["obs"]

    def get_number_of_agents(self):
        return 1

    def get_env_info(self):
        return {"action_space": self.action_space, "observation_space": self.observation_space,
                "amp_observation_space": self.amp_observation_space}

    # ---- LocoVal getters, vec_task_wrappers.py:47-66 ----
    def get_waypoint_traj(self):
        w = self._waypoints.clone()
        return w - w[:, :1].clone()

    def get_init_pose(self):
        p = self._init_pose.clone()
        return p - p[:, :1].clone()

    def get_init_vel(self):
        return self._init_vel.clone()

    def raw_reward(self):
        return self.sim.rew_raw.to(self.rl_device)

    def close(self):
        self.sim.close()
