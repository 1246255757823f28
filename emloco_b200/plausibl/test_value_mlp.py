"""Drop-in for ``plausibl.test_value_mlp.MLP`` (reference plausibl/test_value_mlp.py:24-113): the trajectory-only value
network ``Linear(24,12) -> ReLU -> Linear(12,6) -> ReLU -> Linear(6,1)`` (no sigmoid) on a heading-local 12 x 2 trajectory.

Same attributes (``_value_mlp``, ``_value_logits``), same construction-time initialisation (default nn.Linear weights, zero
biases, uniform(-1, 1) logits layer :44-50), same ``load_weights(path)`` (two non-strict ``load_state_dict`` calls on one file,
:93-104) and ``forward(trajs)``.  The arithmetic runs in ``plausibl_mlp_kernel`` (csrc/locoval.cu) through
``emloco_plausibl_mlp_forward``; there is no CPU path.  Like the reference class this is not an ``nn.Module``.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from .. import _lib


class MLP:
    def __init__(self, **kwargs):
        self._build_value_mlp()
        self._pack = self._pack_key = None

    def _build_value_mlp(self):
        units, k = (12, 6), 24
        layers = []
        for u in units:
            layers += [nn.Linear(k, u), nn.ReLU()]
            k = u
        self._value_mlp = nn.Sequential(*layers)
        self._value_logits = nn.Linear(k, 1)
        for m in self._value_mlp:
            if isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)
        nn.init.uniform_(self._value_logits.weight, -1.0, 1.0)
        nn.init.zeros_(self._value_logits.bias)

    def load_weights(self, path):
        sd = torch.load(path, map_location="cpu")
        self._value_mlp.load_state_dict(sd, strict=False)
        self._value_logits.load_state_dict(sd, strict=False)

    def to(self, device):
        self._value_mlp.to(device); self._value_logits.to(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def _weights(self):
        ps = [self._value_mlp[0].weight, self._value_mlp[0].bias, self._value_mlp[2].weight, self._value_mlp[2].bias,
              self._value_logits.weight, self._value_logits.bias]
        key = tuple((p.data_ptr(), p._version, p.device) for p in ps)
        if key != self._pack_key:
            self._pack = torch.cat([p.detach().reshape(-1).float() for p in ps]).contiguous()
            self._pack_key = key
        return self._pack

    def forward(self, trajs):
        """trajs [B, 24] float32 CUDA -> values [B, 1]."""
        if not (torch.is_tensor(trajs) and trajs.is_cuda):
            raise _lib.EmlocoError("plausibl MLP (emloco_b200) runs on CUDA tensors only; there is no CPU fallback")
        w = self._weights()
        if w.device != trajs.device:
            raise _lib.EmlocoError("plausibl MLP: parameters and input are on different devices (call .to(device))")
        x = trajs.reshape(-1, 24)
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        out = torch.empty(x.shape[0], 1, device=x.device, dtype=torch.float32)
        if x.shape[0]:
            p = lambda t: C.c_void_p(t.data_ptr())
            _lib.check(_lib.load().emloco_plausibl_mlp_forward(p(x), p(w), p(out), x.shape[0],
                                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "emloco_plausibl_mlp_forward")
        return out.reshape(*trajs.shape[:-1], 1)

    __call__ = forward
