"""Host-side handle over the C ABI: the `gym.*` tensor seam of the reference
(isaacgym/python/isaacgym/gymtorch.py:61-106 `wrap_tensor`; pacer/pacer/env/tasks/humanoid.py:137-216 tensor
views; base_task.py:245-265 step driver), one process per GPU.

PyTorch is used for device memory, streams and views only - every computation on the path is a
hand-written kernel in libemloco_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .model import build_model_arrays

_DT = {0: (torch.float32, "<f4"), 1: (torch.int64, "<i8"), 2: (torch.int16, "<i2")}


class _DevBuf:
    """Minimal __cuda_array_interface__ holder so torch can alias sim-owned device memory (zero copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


def wrap_tensor(ptr, shape, dtype_code, device, owner=None):
    """gymtorch.wrap_tensor equivalent: zero-copy torch view of a device buffer."""
    tdt, typestr = _DT[dtype_code]
    t = torch.as_tensor(_DevBuf(ptr, shape, typestr, owner), device=torch.device("cuda", device))
    assert t.data_ptr() == int(ptr), "wrap_tensor must alias, not copy"
    return t


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class EmlocoSim:
    """Owns one `emloco_sim*`.  Tensor attributes alias sim memory with stable addresses."""

    def __init__(self, num_envs, device=0, model_arrays=None, **cfg_over):
        if not torch.cuda.is_available():
            raise _lib.EmlocoError("emloco_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.model_arrays = build_model_arrays() if model_arrays is None else model_arrays
        self.cfg = _lib.default_cfg(num_envs=int(num_envs), device=int(device), **cfg_over)
        self._model = _lib.make_model(self.model_arrays)
        h = C.c_void_p()
        torch.cuda.set_device(device)
        torch.cuda.init()
        _lib.check(self.lib.emloco_create(C.byref(self.cfg), C.byref(self._model), C.byref(h)), "emloco_create")
        self._h = h
        self._traj_on, self._traj_keep = 0, None
        self.num_envs = int(num_envs)
        self.device = int(device)
        from .model import rest_root_height
        self.rest_height = float(rest_root_height(self.model_arrays))
        self._tensors = {}
        for name in _lib.T_IDS:
            self._tensors[name] = self._acquire(name)

    # -- gym.acquire_*_tensor ---------------------------------------------------------------
    def _acquire(self, name):
        p = C.c_void_p(); shape = (C.c_int64 * 4)(); nd = C.c_int32(); dt = C.c_int32()
        _lib.check(self.lib.emloco_tensor(self._h, _lib.T_IDS[name], C.byref(p), shape, C.byref(nd), C.byref(dt)),
                   f"emloco_tensor({name})")
        return wrap_tensor(p.value, [shape[i] for i in range(nd.value)], dt.value, self.device, owner=self)

    def tensor(self, name):
        return self._tensors[name]

    def __getattr__(self, name):
        t = self.__dict__.get("_tensors", {})
        if name in t:
            return t[name]
        raise AttributeError(name)

    # -- per-env body models (has_shape_variation, humanoid.py:597-739,905-910) --------------
    def set_env_models(self, models, shape_of_env=None, betas=None):
        """models: list of model-array dicts (model.build_model_arrays / scaled_model_arrays ...), one per distinct shape;
        shape_of_env: int array [N] (default env i -> shape i % len(models), the reference's assignment, humanoid.py:606).
        betas: optional [num_shapes, 17] shape parameters copied into the `betas` observation tensor.  None restores the
        shared model."""
        if models is None:
            _lib.check(self.lib.emloco_set_env_models(self._h, None), "emloco_set_env_models")
            self._env_models = None
            return
        from .model import pack_env_model
        packed = np.stack([pack_env_model(m) for m in models])
        idx = np.arange(self.num_envs) % len(models) if shape_of_env is None else np.asarray(shape_of_env, np.int64)
        per_env = np.ascontiguousarray(packed[idx], dtype=np.float32)
        _lib.check(self.lib.emloco_set_env_models(self._h, per_env.ctypes.data_as(C.c_void_p)), "emloco_set_env_models")
        self._env_models = (models, idx)
        if betas is not None:
            self.betas.copy_(torch.as_tensor(np.asarray(betas, np.float32)[idx]).to(self.betas.device))

    # -- terrain -----------------------------------------------------------------------------
    def set_height_field(self, samples):
        a = np.ascontiguousarray(samples, dtype=np.int16)
        _lib.check(self.lib.emloco_set_height_field(self._h, a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]),
                   "emloco_set_height_field")
        self._tensors["height"] = self._acquire("height")

    # -- gym.set_dof_position_target_tensor / simulate / set_*_state_tensor_indexed ----------
    def set_pd_targets(self, targets):
        assert targets.is_cuda and targets.dtype == torch.float32 and targets.is_contiguous()
        _lib.check(self.lib.emloco_set_pd_targets(self._h, _ptr(targets), _stream()), "emloco_set_pd_targets")

    def simulate(self):
        _lib.check(self.lib.emloco_simulate(self._h, _stream()), "emloco_simulate")

    def reset_indexed(self, env_ids=None):
        if env_ids is None:
            _lib.check(self.lib.emloco_reset_indexed(self._h, None, 0, _stream()), "emloco_reset_indexed")
            return
        ids = env_ids.to(device=self.root_state.device, dtype=torch.int32).contiguous()
        if ids.numel() == 0:
            return
        _lib.check(self.lib.emloco_reset_indexed(self._h, _ptr(ids), ids.numel(), _stream()), "emloco_reset_indexed")

    def reset_done(self, init_root, init_dof):
        """env_reset(done_indices) on the device: envs with reset_buf set restart from init_root [N,13] / init_dof [N*69,2]."""
        for t, shp in ((init_root, (self.num_envs, 13)), (init_dof, (self.num_envs * _lib.ND, 2))):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shp):
                raise _lib.EmlocoError(f"reset_done: initial state must be contiguous float32 CUDA of shape {shp}")
        _lib.check(self.lib.emloco_reset_done(self._h, _ptr(init_root), _ptr(init_dof), _stream()), "emloco_reset_done")
        _lib.launch_count += int(self._traj_on == 1)  # the appended trajectory-reset stage

    def traj_cfg(self, flags=0, seed=0, pool=None, uniform=None, waypoint_traj=None, init_pose=None, init_vel=None,
                 inverted=None, origin_relative=True, **over):
        """emloco_traj_cfg with the pacer.yaml defaults (TrajGenerator args, humanoid_traj.py:113-119).  Tensors are CUDA,
        contiguous: pool [P,101,3] f32, uniform [N,>=405] f32, waypoint_traj [N,<=15,3], init_pose [N,24,3], init_vel [N,2]
        f32, inverted [N] uint8.  The sim keeps references to them."""
        c = _lib.TrajCfg(dtheta_max=2.0, speed_min=0.0005, speed_max=3.0, accel_max=2.0, sharp_turn_prob=0.02,
                         hybrid_init_prob=0.5, flags=int(flags), origin_relative=int(bool(origin_relative)), seed=int(seed))
        for k, v in over.items():
            if not hasattr(c, k):
                raise _lib.EmlocoError(f"unknown traj cfg field {k}")
            setattr(c, k, v)
        N = self.num_envs
        want = dict(pool=(pool, torch.float32, None), uniform=(uniform, torch.float32, None),
                    waypoint_traj=(waypoint_traj, torch.float32, None), init_pose=(init_pose, torch.float32, (N, _lib.NB, 3)),
                    init_vel=(init_vel, torch.float32, (N, 2)), inverted=(inverted, torch.uint8, (N,)))
        keep = []
        for k, (t, dt, shp) in want.items():
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == dt and t.is_contiguous() and (shp is None or tuple(t.shape) == shp)):
                raise _lib.EmlocoError(f"traj cfg: {k} must be contiguous {dt} CUDA" + (f" of shape {shp}" if shp else ""))
            setattr(c, k, t.data_ptr()); keep.append(t)
        if pool is not None:
            if pool.dim() != 3 or tuple(pool.shape[1:]) != (101, 3):
                raise _lib.EmlocoError("traj cfg: pool must be [P,101,3]")
            c.pool_count = pool.shape[0]
        if waypoint_traj is not None:
            if waypoint_traj.dim() != 3 or waypoint_traj.shape[0] != N or waypoint_traj.shape[2] != 3 or not 1 <= waypoint_traj.shape[1] <= 15:
                raise _lib.EmlocoError("traj cfg: waypoint_traj must be [N, 1..15, 3]")
            c.num_waypoints = waypoint_traj.shape[1]
        if uniform is not None:
            if uniform.dim() != 2 or uniform.shape[0] != N:
                raise _lib.EmlocoError("traj cfg: uniform must be [N, >=405]")
            c.ld_uniform = uniform.shape[1]
        c._keep = keep
        return c

    def traj_reset(self, cfg=None):
        """TrajGenerator.reset + _reset_task outputs for the envs whose reset_buf is set right now.  cfg None: the deferred
        stage stored by set_traj_reset(cfg with TRAJ_DEFERRED), which also clears reset/terminate."""
        _lib.check(self.lib.emloco_traj_reset(self._h, None if cfg is None else C.byref(cfg), _stream()), "emloco_traj_reset")

    def set_traj_reset(self, cfg=None):
        """reset_done regenerates the trajectories of the envs it resets (None: off)."""
        _lib.check(self.lib.emloco_set_traj_reset(self._h, None if cfg is None else C.byref(cfg)), "emloco_set_traj_reset")
        self._traj_keep = cfg
        self._traj_on = 0 if cfg is None else (2 if cfg.flags & _lib.TRAJ_DEFERRED else 1)

    def set_post_sinks(self, sinks=None):
        """Optional extra outputs of post_step / reset_done (emloco_post_sinks); None clears them."""
        _lib.check(self.lib.emloco_set_post_sinks(self._h, None if sinks is None else C.byref(sinks)), "emloco_set_post_sinks")

    def post_step(self, advance_progress=True):
        _lib.check(self.lib.emloco_post_step(self._h, int(bool(advance_progress)), _stream()), "emloco_post_step")

    def step(self, actions):
        """BaseTask.step: actions [N,69] float32 on this device."""
        if not (actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
                and tuple(actions.shape) == (self.num_envs, _lib.ND)):
            raise _lib.EmlocoError("actions must be a contiguous float32 CUDA tensor of shape [num_envs, 69]")
        _lib.check(self.lib.emloco_step(self._h, _ptr(actions), _stream()), "emloco_step")

    def physics_step(self, actions):
        """pre_physics_step + control_freq_inv x simulate only; follow with post_step()."""
        if not (actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
                and tuple(actions.shape) == (self.num_envs, _lib.ND)):
            raise _lib.EmlocoError("actions must be a contiguous float32 CUDA tensor of shape [num_envs, 69]")
        _lib.check(self.lib.emloco_physics_step(self._h, _ptr(actions), _stream()), "emloco_physics_step")

    def step_host(self, actions, obs=None, rew=None, reset=None, amp_obs=None):
        """Host-buffer step (numpy arrays); copies inside the call."""
        a = np.ascontiguousarray(actions, dtype=np.float32)
        assert a.shape == (self.num_envs, _lib.ND)

        def p(x, dt, shape):
            if x is None:
                return None
            assert x.dtype == dt and x.flags.c_contiguous and x.shape == shape
            return x.ctypes.data_as(C.c_void_p)
        N = self.num_envs
        _lib.check(self.lib.emloco_step_host(self._h, a.ctypes.data_as(C.c_void_p), p(obs, np.float32, (N, 1422)),
                                             p(rew, np.float32, (N,)), p(reset, np.int64, (N,)),
                                             p(amp_obs, np.float32, (N, 3090))), "emloco_step_host")

    def sync(self):
        _lib.check(self.lib.emloco_sync(self._h), "emloco_sync")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._tensors = {}
            self.lib.emloco_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- stateless kernels ---------------------------------------------------------------------
def gae(dones, values, rewards, next_values, gamma=0.99, tau=0.95):
    """discount_values (common_agent.py:573-587).  Inputs [T,N] or [T,N,1]; returns (adv, returns) same shape."""
    shp = values.shape
    T = shp[0]
    N = values.numel() // T
    f = lambda t: t.reshape(T, N).contiguous().float()
    d, v, r, nv = f(dones), f(values), f(rewards), f(next_values)
    adv = torch.empty_like(v); ret = torch.empty_like(v)
    _lib.check(_lib.load().emloco_gae(_ptr(d), _ptr(v), _ptr(r), _ptr(nv), _ptr(adv), _ptr(ret), T, N, gamma, tau, _stream()),
               "emloco_gae")
    return adv.reshape(shp), ret.reshape(shp)


def linear(x, weight, bias=None, relu=False, mean=None, var=None, eps=1e-5, out=None, tensor_cores=False):
    """y = act(norm(x) @ W^T + b); x may be a column slice of a wider row-major buffer (stride(0) = ld)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32 and x.is_cuda
    M, K = x.shape
    N = weight.shape[0]
    assert weight.shape == (N, K) and weight.is_contiguous() and weight.dtype == torch.float32
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(1) == 1
    mean32 = None if mean is None else mean.float().contiguous()
    var32 = None if var is None else var.float().contiguous()
    b = None if bias is None else bias.contiguous()
    _lib.check(_lib.load().emloco_linear(_ptr(x), x.stride(0), _ptr(weight), _ptr(b), _ptr(out), out.stride(0), M, N, K,
                                         _ptr(mean32), _ptr(var32), eps, int(relu), int(tensor_cores), _stream()),
               "emloco_linear")
    return out
