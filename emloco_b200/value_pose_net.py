"""LocoVal drop-in: same surface as the reference ``ValuePoseNet``
(pacer/pacer/learning/value_pose_net.py:10-159): constructor flags, ``forward``,
``calc_embodied_motion_loss``, state-dict keys ``_network.fc{1,2,3}.{weight,bias}``, ``.eval()/.to()``.
Callers: social-transmotion/train_jta.py:198-204,288-308, evaluate_jta.py:298-302,574-584,
pacer/pacer/learning/amp_value_players.py:128-137,365-373, amp_continuous_value.py:123-145.

The forward pass and the gradient w.r.t. the predicted trajectory run in the fused CUDA kernels of
csrc/locoval.cu through the C ABI.  Like the reference, ``forward`` rotates and zeroes the caller's
``init_pose`` IN PLACE (value_pose_net.py:97,141-144); pass ``mutate_pose=False`` to opt out.
When the LocoVal weights themselves require grad (the fine-tuning step of
amp_continuous_value.py:123-145, one small batch per rollout step) the weight gradients come from
PyTorch autograd, as the survey scopes it ("backward stays PyTorch autograd").
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

F_POSE, F_VEL, F_HIDE_TOE, F_HIDE_SPINE, F_NORMALIZE, F_WRITEBACK = 1, 2, 4, 8, 16, 32


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _LocoValFn(torch.autograd.Function):
    """value = LocoVal(traj, pose, vel).  With F_WRITEBACK the kernel rotates / zeroes `pose` in place like the reference
    (value_pose_net.py:97,141-144) and the tensor is returned as a second, dirty output: like in the reference it then carries
    autograd history, so a later call that reuses it (multi-modal loop, train_jta.py:294-296) back-propagates through it."""

    @staticmethod
    def forward(ctx, traj, pose, vel, wpack, flags, T):
        B = traj.shape[0]
        stride = traj.shape[-1]
        value = torch.empty(B, 1, device=traj.device, dtype=torch.float32)
        dirty = pose is not None and bool(flags & F_WRITEBACK)
        ctx.flags, ctx.T, ctx.dirty = flags, T, dirty
        if dirty:
            ctx.mark_dirty(pose)
        if B == 0:                                  # empty batch: nothing to launch (data_ptr() of an empty tensor is NULL)
            ctx.save_for_backward(traj, None, vel, wpack)
            return (value, pose) if dirty else value
        need_grad = any(ctx.needs_input_grad[:2])
        pose_saved = None
        if pose is not None and need_grad:
            pose_saved = pose.detach().clone() if dirty else pose.detach()      # the pose as it came in
        _lib.check(_lib.load().emloco_locoval_forward(_p(traj), stride, T, _p(pose), _p(vel), _p(wpack), _p(value), B,
                                                      flags, _stream()), "emloco_locoval_forward")
        ctx.save_for_backward(traj, pose_saved, vel, wpack)
        return (value, pose) if dirty else value

    @staticmethod
    def backward(ctx, gvalue, gpose_out=None):
        traj, pose, vel, wpack = ctx.saved_tensors
        g = gvalue.contiguous().float()
        gtraj = torch.empty_like(traj)
        want_pose = ctx.needs_input_grad[1] and pose is not None
        if traj.shape[0] == 0:
            return gtraj, (torch.zeros_like(pose) if want_pose else None), None, None, None, None
        gpo = gpose_out.contiguous().float() if (gpose_out is not None and pose is not None) else None
        gpi = torch.empty_like(pose) if want_pose else None
        _lib.check(_lib.load().emloco_locoval_backward_pose(_p(traj), traj.shape[-1], ctx.T, _p(pose), _p(vel), _p(wpack), _p(g),
                                                            _p(gtraj), _p(gpo), _p(gpi), traj.shape[0], ctx.flags & ~F_WRITEBACK,
                                                            _stream()), "emloco_locoval_backward_pose")
        return gtraj, gpi, None, None, None, None


class ValuePoseNet(nn.Module):
    def __init__(self, use_pose, use_vel, hide_toe=True, hide_spine=True, normalize=True, vru=False, mutate_pose=True,
                 **kwargs):
        super().__init__(**kwargs)
        self.use_pose, self.use_vel = bool(use_pose), bool(use_vel)
        self.hide_toe, self.hide_spine, self.normalize, self.use_vru = hide_toe, hide_spine, normalize, vru
        self.mutate_pose = mutate_pose
        self.traj_size = 13 * 2 if not vru else 5 * 2                     # value_pose_net.py:37
        self.pose_size, self.vel_size = 24 * 3, 2
        n_in = self.traj_size + (self.pose_size if use_pose else 0) + (self.vel_size if use_vel else 0)
        fc1_out = int(n_in / 2) - 1                                       # :52
        fc2_out = int(fc1_out / 2)                                        # :53
        self._network = nn.Sequential()
        self._network.add_module("fc1", nn.Linear(n_in, fc1_out))
        self._network.add_module("relu1", nn.ReLU())
        self._network.add_module("fc2", nn.Linear(fc1_out, fc2_out))
        self._network.add_module("relu2", nn.ReLU())
        self._network.add_module("fc3", nn.Linear(fc2_out, 1))
        self._network.add_module("sigmoid", nn.Sigmoid())
        for m in self._network:                                           # :62-66
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)
        self.criterion = nn.MSELoss()
        self._pack = None
        self._pack_key = None

    # ---- weights packed in state-dict order for the kernel (cached until a parameter changes) ----
    def _weights(self):
        ft = getattr(self, "_ft", None)
        if ft is not None and self._network.fc1.weight.data_ptr() == ft["flat"].data_ptr():
            return ft["flat"]                    # fine-tuning: the parameters ARE views of the packed buffer (enable_finetune)
        ps = [self._network.fc1.weight, self._network.fc1.bias, self._network.fc2.weight, self._network.fc2.bias,
              self._network.fc3.weight, self._network.fc3.bias]
        key = tuple((p.data_ptr(), p._version, p.device) for p in ps)
        if key != self._pack_key:
            self._pack = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
            self._pack_key = key
        return self._pack

    def _flags(self):
        return ((F_POSE if self.use_pose else 0) | (F_VEL if self.use_vel else 0) | (F_HIDE_TOE if self.hide_toe else 0)
                | (F_HIDE_SPINE if self.hide_spine else 0) | (F_NORMALIZE if self.normalize else 0)
                | (F_WRITEBACK if self.mutate_pose else 0))

    def _weights_need_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self._network.parameters()) and self.training

    def forward(self, waypoint_traj, init_pose=None, init_vel=None):
        if self.use_pose:
            assert init_pose is not None, "init_pose should be included"   # :114
        if self.use_vel:
            assert init_vel is not None, "init_vel should be included"     # :126
        if not waypoint_traj.is_cuda:
            raise _lib.EmlocoError("ValuePoseNet (emloco_b200) runs on CUDA tensors only; there is no CPU fallback")
        if self._weights_need_grad():
            return self._forward_autograd(waypoint_traj, init_pose, init_vel)
        T = self.traj_size // 2
        traj = waypoint_traj
        if traj.dtype != torch.float32 or not traj.is_contiguous():
            traj = traj.float().contiguous()
        assert traj.dim() == 3 and traj.shape[1] == T and traj.shape[2] >= 2, "waypoint_traj must be [B, T, >=2]"
        pose = vel = None
        if self.use_pose:
            pose = init_pose
            if pose.dtype != torch.float32 or not pose.is_contiguous():
                if self.mutate_pose:
                    raise _lib.EmlocoError("init_pose must be contiguous float32 to be rotated in place like the reference does")
                pose = pose.float().contiguous()
            assert pose.shape[-2:] == (24, 3) and pose.shape[0] == traj.shape[0]
        if self.use_vel:
            vel = init_vel[:, :2].float().contiguous()
        out = _LocoValFn.apply(traj, pose, vel, self._weights(), self._flags(), T)
        return out[0] if isinstance(out, tuple) else out

    # ---- fine-tuning inside the rollout (amp_continuous_value.py:122-146, optimiser common_agent.py:94-96) ----
    def enable_finetune(self):
        """Re-homes the six parameters as views of ONE packed fp32 buffer (the layout the kernels read) so that the fused
        AdamW step updates them in place, and allocates the optimiser state."""
        ps = [self._network.fc1.weight, self._network.fc1.bias, self._network.fc2.weight, self._network.fc2.bias,
              self._network.fc3.weight, self._network.fc3.bias]
        if not ps[0].is_cuda:
            raise _lib.EmlocoError("ValuePoseNet.enable_finetune needs the module on a CUDA device; there is no CPU fallback")
        flat = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
        off = 0
        for p in ps:
            n = p.numel()
            p.data = flat[off:off + n].view_as(p)
            off += n
        self._pack, self._pack_key = flat, tuple((p.data_ptr(), p._version, p.device) for p in ps)
        ft = getattr(self, "_ft", None)
        if ft is None or ft["m"].device != flat.device:
            z = lambda n: torch.zeros(n, device=flat.device, dtype=torch.float32)
            ft = dict(m=z(flat.numel()), v=z(flat.numel()), step=z(1), stats=z(4), ws=None)
        ft["flat"] = flat
        self._ft = ft
        return self

    def finetune_step(self, waypoint_traj, init_pose, init_vel, game_combined_rewards, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                      weight_decay=1e-4, min_cum_rewards=-10.0, max_cum_rewards=100.0):
        """One `_do_finetune` block: MSE(sum) of the scores of the envs with game_combined_rewards != 0 against their
        normalised rewards, backward, AdamW step, game_combined_rewards zeroed for those envs - two launches, no host sync.
        Inputs are CUDA float32 contiguous: traj [N,T,>=2], pose [N,24,3], vel [N,2], game_combined_rewards [N]."""
        ft = getattr(self, "_ft", None)
        if ft is None or self._network.fc1.weight.data_ptr() != ft["flat"].data_ptr():
            self.enable_finetune()
            ft = self._ft
        N, T = waypoint_traj.shape[0], self.traj_size // 2
        for name, t, shp in (("waypoint_traj", waypoint_traj, None), ("init_pose", init_pose, (N, 24, 3) if self.use_pose else None),
                             ("init_vel", init_vel, (N, 2) if self.use_vel else None), ("game_combined_rewards", game_combined_rewards, (N,))):
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and (shp is None or tuple(t.shape) == shp)):
                raise _lib.EmlocoError(f"finetune_step: {name} must be contiguous float32 CUDA" + (f" of shape {shp}" if shp else ""))
        assert waypoint_traj.dim() == 3 and waypoint_traj.shape[1] == T and waypoint_traj.shape[2] >= 2
        need = _lib.load().emloco_locoval_train_workspace_bytes(N)
        if ft["ws"] is None or ft["ws"].numel() < need:
            ft["ws"] = torch.empty(need, device=ft["flat"].device, dtype=torch.uint8)
        _lib.check(_lib.load().emloco_locoval_train_step(
            _p(waypoint_traj), waypoint_traj.shape[-1], T, _p(init_pose) if self.use_pose else None,
            _p(init_vel) if self.use_vel else None, _p(game_combined_rewards), _p(ft["flat"]), _p(ft["m"]), _p(ft["v"]), _p(ft["step"]),
            _p(ft["stats"]), _p(ft["ws"]), N, lr, betas[0], betas[1], eps, weight_decay, min_cum_rewards, max_cum_rewards,
            self._flags() & ~F_WRITEBACK, _stream()), "emloco_locoval_train_step")

    def finetune_stats(self, reset=True):
        """(vnet_loss, mean vnet_pred, mean vnet_gt, count) accumulated since the last reset (common_agent.py:204-207,241-243)."""
        s = self._ft["stats"].tolist()
        if reset:
            self._ft["stats"].zero_()
        n = max(s[3], 1.0)
        return s[0] / n, s[1] / n, s[2] / n, int(s[3])

    def _forward_autograd(self, traj, pose, vel):
        """Weight-gradient path (LocoVal fine-tuning only): plain torch ops, same math as the kernel."""
        xy = traj[..., :2]
        if self.normalize:
            x, y = xy[:, 1, 0], xy[:, 1, 1]
            near = x.abs() < 1e-10
            x = x * (~near) + near * 1e-10
            ang = torch.atan2(y, x)
            c, s = torch.cos(ang), torch.sin(ang)
            R = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
            xy = torch.bmm(xy, R)
            if pose is not None:
                pose[..., :2] = torch.bmm(pose[:, :, :2].clone(), R)
            if vel is not None:
                vel = torch.bmm(vel[:, :2].clone().unsqueeze(1), R)[:, 0]
        feats = [xy.reshape(-1, self.traj_size)]
        if self.use_pose:
            if self.hide_toe:
                pose[:, [4, 8]] = 0
            if self.hide_spine:
                pose[:, [9, 10, 11]] = 0
            feats.append(pose.reshape(-1, self.pose_size))
        if self.use_vel:
            feats.append(vel.reshape(-1, self.vel_size))
        return self._network(torch.cat(feats, dim=-1))

    def calc_embodied_motion_loss(self, pred_traj, init_pose=None, init_vel=None):
        pred_value = self.forward(pred_traj, init_pose, init_vel)
        loss = self.criterion(pred_value, torch.ones_like(pred_value))     # :157
        return pred_value, loss


def score_host(traj, pose, vel, state_dict, device=0, use_pose=True, use_vel=True, hide_toe=True, hide_spine=True,
               normalize=True):
    """Batched LocoVal scoring straight from host (numpy) buffers through the C ABI - replaces the
    batch-of-1 triple loop of social-transmotion/evaluate_jta.py:214-357."""
    import numpy as np
    traj = np.ascontiguousarray(traj, np.float32)
    B, T, stride = traj.shape
    keys = ["_network.fc1.weight", "_network.fc1.bias", "_network.fc2.weight", "_network.fc2.bias",
            "_network.fc3.weight", "_network.fc3.bias"]
    w = np.concatenate([np.asarray(state_dict[k], np.float32).reshape(-1) for k in keys])
    flags = ((F_POSE if use_pose else 0) | (F_VEL if use_vel else 0) | (F_HIDE_TOE if hide_toe else 0)
             | (F_HIDE_SPINE if hide_spine else 0) | (F_NORMALIZE if normalize else 0))
    out = np.empty(B, np.float32)
    pp = np.ascontiguousarray(pose, np.float32) if use_pose else None
    vv = np.ascontiguousarray(vel, np.float32) if use_vel else None
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.load().emloco_locoval_forward_host(vp(traj), stride, T, vp(pp), vp(vv), vp(w), vp(out), B, flags, device),
               "emloco_locoval_forward_host")
    return out
