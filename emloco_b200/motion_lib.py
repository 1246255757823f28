"""Device-side motion library: the reset / demo half of SURVEY 8 row f2.

Reference: `MotionLibSMPL` (pacer/pacer/utils/motion_lib_smpl.py) - `sample_motions` :390-395, `sample_time` :398-407,
`get_motion_state_smpl` :485-563 - and the two consumers in `HumanoidAMP` (env/tasks/humanoid_amp.py): `fetch_amp_obs_demo` /
`build_amp_obs_demo` :168-220 (the discriminator's real samples) and `_sample_ref_state` -> `_reset_ref_state_init` (the mocap
state an env restarts from).  The reference loads AMASS clips from joblib pickles of poselib `SkeletonMotion`s; those data are
not redistributable, so this class takes the flat per-frame arrays directly (`from_arrays`, the layout MotionLibSMPL builds at
:248-330) and `emloco_b200.synthetic.synthetic_motion_lib` generates walking-like clips in that layout.

Sampling (multinomial ids, uniform times) uses torch's device RNG; the state / observation arithmetic runs in
`emloco_motion_state` / `emloco_amp_obs_demo` (csrc/motion.cu).  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .sim import _ptr, _stream

AMP_STEPS, AMP_STEP_DIM = 15, 206


class MotionLibSMPL:
    FRAME_KEYS = ("gts", "grs", "lrs", "gvs", "gavs", "dvs")

    def __init__(self, arrays, device=0, seed=0):
        if not torch.cuda.is_available():
            raise _lib.EmlocoError("MotionLibSMPL (emloco_b200) needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", device)
        f = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev).contiguous()
        self.t = {k: f(arrays[k], torch.float32) for k in self.FRAME_KEYS}
        F = self.t["gts"].shape[0]
        for k, shp in (("gts", (F, 24, 3)), ("grs", (F, 24, 4)), ("lrs", (F, 24, 4)), ("gvs", (F, 24, 3)), ("gavs", (F, 24, 3)), ("dvs", (F, 23, 3))):
            if tuple(self.t[k].shape) != shp:
                raise ValueError(f"motion library array {k} must have shape {shp}, got {tuple(self.t[k].shape)}")
        self.motion_lengths = f(arrays["motion_lengths"], torch.float32)
        self.motion_dt = f(arrays["motion_dt"], torch.float32)
        self.motion_num_frames = f(arrays["motion_num_frames"], torch.int32)
        self.length_starts = f(arrays["length_starts"], torch.int32)
        self.motion_bodies = f(arrays["motion_bodies"], torch.float32)
        M = self.motion_lengths.numel()
        w = arrays.get("weights")
        self._prob = f(np.full(M, 1.0 / M) if w is None else np.asarray(w) / np.sum(w), torch.float32)     # _sampling_batch_prob
        self.num_motions, self.device = M, dev
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        L = _lib.MotionLib()
        for k in self.FRAME_KEYS:
            setattr(L, "d_" + k, self.t[k].data_ptr())
        L.d_length, L.d_dt, L.d_bodies = self.motion_lengths.data_ptr(), self.motion_dt.data_ptr(), self.motion_bodies.data_ptr()
        L.d_num_frames, L.d_start, L.num_motions = self.motion_num_frames.data_ptr(), self.length_starts.data_ptr(), M
        self._L = L

    @classmethod
    def from_arrays(cls, arrays, **kw):
        return cls(arrays, **kw)

    # ---- sampling (motion_lib_smpl.py:390-407) ----
    def sample_motions(self, n):
        if int(n) <= 0:
            return torch.zeros(0, dtype=torch.int32, device=self.device)
        return torch.multinomial(self._prob, num_samples=int(n), replacement=True, generator=self.gen).to(torch.int32)

    def sample_time(self, motion_ids, truncate_time=None):
        phase = torch.rand(motion_ids.shape, device=self.device, generator=self.gen)
        length = self.motion_lengths[motion_ids.long()]
        if truncate_time is not None:
            assert truncate_time >= 0.0
            length = length - truncate_time
        return phase * length

    # ---- get_motion_state_smpl (:485-563) ----
    def get_motion_state_smpl(self, motion_ids, motion_times, root_out=None, dof_out=None, full=False):
        """-> dict(root_state [n,13], dof_state [n,69,2], key_pos [n,4,3][, rb_state [n,24,13]]).  root_out / dof_out: write
        straight into caller buffers (e.g. the init_root / init_dof rows a Rollout resets from)."""
        ids = motion_ids.to(self.device, torch.int32).contiguous()
        times = motion_times.to(self.device, torch.float32).contiguous()
        n = ids.numel()
        f = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
        root = f(n, 13) if root_out is None else root_out
        dof = f(n, 69, 2) if dof_out is None else dof_out
        assert root.is_contiguous() and dof.is_contiguous() and root.numel() == n * 13 and dof.numel() == n * 138
        key, rb = f(n, 4, 3), (f(n, 24, 13) if full else None)
        if n:
            _lib.check(_lib.load().emloco_motion_state(C.byref(self._L), _ptr(ids), _ptr(times), n, _ptr(root), _ptr(dof), _ptr(key), _ptr(rb),
                                                       _stream()), "emloco_motion_state")
        out = dict(root_state=root, dof_state=dof, key_pos=key)
        if full:
            out["rb_state"] = rb
        return out

    # ---- HumanoidAMP.fetch_amp_obs_demo (humanoid_amp.py:168-220) ----
    def fetch_amp_obs_demo(self, num_samples, dt=2.0 / 60.0, num_steps=AMP_STEPS, motion_ids=None, motion_times0=None):
        ids = self.sample_motions(num_samples) if motion_ids is None else motion_ids.to(self.device, torch.int32).contiguous()
        t0 = self.sample_time(ids) if motion_times0 is None else motion_times0.to(self.device, torch.float32).contiguous()
        out = torch.empty(ids.numel(), num_steps * AMP_STEP_DIM, device=self.device, dtype=torch.float32)
        if ids.numel() == 0:
            return out
        _lib.check(_lib.load().emloco_amp_obs_demo(C.byref(self._L), _ptr(ids), _ptr(t0), ids.numel(), num_steps, float(dt), _ptr(out), _stream()),
                   "emloco_amp_obs_demo")
        return out

    # ---- _sample_ref_state -> _reset_ref_state_init (humanoid_amp.py:295-317,407-470): the state envs restart from ----
    def sample_reset_state(self, init_root, init_dof, keep_xy=True):
        """Draws one (motion, time) per env and writes its root / DOF state into the buffers `emloco_reset_done` restarts envs
        from (init_root [N,13], init_dof [N*69,2]).  keep_xy: the env keeps its own start position on the terrain patch (the
        reference adds the env's terrain offset to the clip's root position); height, orientation and velocities come from
        the clip.  The SMPL-mesh ground-height correction of humanoid_amp.py:321-379 needs the licensed body model: not done."""
        N = init_root.shape[0]
        ids = self.sample_motions(N)
        times = self.sample_time(ids)
        xy = init_root[:, 0:2].clone() if keep_xy else None
        self.get_motion_state_smpl(ids, times, root_out=init_root, dof_out=init_dof.view(N, 69, 2))
        if keep_xy:
            init_root[:, 0:2] = xy
        return ids, times
