"""The PPO / AMP update step on one GPU: `AMPValueAgent.calc_gradients`
(pacer/pacer/learning/amp_continuous_value.py:276-428) - model forward in training mode, actor / critic / task-value / bound
losses (learning/common_agent.py:594-602,657-683, amp_continuous_value.py:430-444), discriminator loss with logit
regularisation, gradient penalty and weight decay (learning/amp_continuous.py:536-616), backward, gradient average over ranks
(Horovod `optimizer.synchronize()` :386-394 -> one NCCL all-reduce of the flat gradient), nn.utils.clip_grad_norm_(50) and
the torch.optim.Adam step (common_agent.py:84-87) - with the running-statistics updates of the input normalisers
(utils/running_mean_std.py:86-96) that `set_train()` switches on.

No autograd: every derivative is written out (oracle/update_oracle.py is the CPU restatement of the same decomposition, pinned
to the reference's own loss code under torch autograd by tests/golden/update_step.npz).  Every dense product - forward, dgrad
(dY W), wgrad (dY^T X) and the six products of the gradient penalty's double backward - is one launch of the tcgen05 bf16x3
GEMM (emloco_linear_bf16x3: y = A W^T with both operands K-major); `emloco_xform` produces the operands (bf16 hi/lo splits in
row-major and transposed orientation), applies ReLU masks and reduces bias gradients.  Parameters, gradients and Adam moments
live in flat fp32 buffers (`FlatParams`): the gradient average is ONE collective over one buffer, clip-norm + Adam two
launches, and the actor / critic first layers are adjacent so that they run as one stacked GEMM in all three passes.

PyTorch provides device memory, the uniform draws of the dropout masks and `torch.distributed`; there is no torch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .policy import ACTIONS, AMP_OBS, OBS, SELF_OBS, TASK_OBS, TRAJ_OBS, AMPSeptValueNetwork, RunningMeanStd, _Split, _pad64, linear_bf16x3
from .sim import _ptr, _stream

# data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml:84-122
DEFAULT_CFG = dict(e_clip=0.2, actor_coef=1.0, critic_coef=5.0, tv_coef=5.0, bounds_loss_coef=10.0, entropy_coef=0.0, disc_coef=5.0,
                   disc_logit_reg=0.01, disc_grad_penalty=5.0, disc_weight_decay=0.0001, dropout_rate=0.3, lr=2e-5, grad_norm=50.0,
                   amp_dropout=True)


def xform(x=None, rowvec=None, rowscale=None, mean=None, var=None, eps=1e-5, gate=None, scale=1.0, y32=None, yT32=None, split=None,
          splitT=None, colsum=None, sumsq=None, M=None, K=None, drop_u=None, drop_rate=0.3):
    """emloco_xform: y = scale * rowscale[m] * src * (gate > 0); see include/emloco.h.  split / splitT: `_Split` (row-major [M,K] /
    transposed [K,M])."""
    if x is not None:
        assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32
        M, K = x.shape
    assert M is not None and K is not None
    if gate is not None:
        assert gate.shape == (M, K) and gate.stride(1) == 1
    if rowscale is not None:
        assert rowscale.numel() == M and rowscale.is_contiguous()
    if drop_u is not None:
        assert drop_u.shape == (M, 19) and drop_u.is_contiguous()
    if split is not None:
        assert split.rows == M and split.K == K
    if splitT is not None:
        assert splitT.rows == K and splitT.K == M
    if y32 is not None:
        assert y32.shape == (M, K) and y32.stride(1) == 1
    if yT32 is not None:
        assert yT32.shape == (K, M) and yT32.stride(1) == 1
    _lib.check(_lib.load().emloco_xform(
        _ptr(x), 0 if x is None else x.stride(0), _ptr(rowvec), _ptr(rowscale), 1, _ptr(mean), _ptr(var), eps, _ptr(gate),
        0 if gate is None else gate.stride(0), _ptr(drop_u), float(drop_rate), float(scale), _ptr(y32), 0 if y32 is None else y32.stride(0), _ptr(yT32),
        0 if yT32 is None else yT32.stride(0), None if split is None else _ptr(split.hi), None if split is None else _ptr(split.lo),
        0 if split is None else split.ld, None if splitT is None else _ptr(splitT.hi), None if splitT is None else _ptr(splitT.lo),
        0 if splitT is None else splitT.ld, _ptr(colsum), _ptr(sumsq), M, K, _stream()), "emloco_xform")


class FlatParams:
    """Parameters of `AMPSeptValueNetwork` re-homed as views of ONE flat fp32 buffer, with a gradient buffer and Adam moments of
    the same layout.  Order: the big GEMM-written weights first (actor / critic first layers adjacent = one stacked [4096,624]
    matrix), then every tensor whose gradient is accumulated by atomics (biases, single-row heads) in one tail block that is
    cleared with one memset per step."""
    GEMM = ("_disc_mlp.0.weight", "_disc_mlp.2.weight",                       # bucket 0: complete when the discriminator's backward is
            "actor_mlp.0.weight", "critic_mlp.0.weight", "actor_mlp.2.weight", "critic_mlp.2.weight", "mu.weight", "_task_mlp.0.weight",
            "_task_mlp.2.weight", "_task_value_mlp.0.weight", "_task_value_mlp.2.weight")
    TAIL = ("actor_mlp.0.bias", "critic_mlp.0.bias", "actor_mlp.2.bias", "critic_mlp.2.bias", "mu.bias", "value.weight", "value.bias",
            "_task_mlp.0.bias", "_task_mlp.2.bias", "_disc_mlp.0.bias", "_disc_mlp.2.bias", "_disc_logits.weight", "_disc_logits.bias",
            "_task_value_mlp.0.bias", "_task_value_mlp.2.bias", "_value_logits.weight", "_value_logits.bias")

    def __init__(self, net: AMPSeptValueNetwork, symmetric=False):
        named = dict(net.named_parameters())
        assert set(self.GEMM) | set(self.TAIL) | {"sigma"} == set(named), sorted(set(named) ^ (set(self.GEMM) | set(self.TAIL) | {"sigma"}))
        dev = named["mu.weight"].device
        if dev.type != "cuda":
            raise _lib.EmlocoError("FlatParams needs the network on a CUDA device; there is no CPU fallback")
        self.off, n = {}, 0
        for k in self.GEMM + self.TAIL:
            if k == self.TAIL[0]:
                self.tail_start = n
            self.off[k] = n
            n += (named[k].numel() + 3) // 4 * 4                       # 16-byte aligned starts
        self.n = n
        self.bucket0 = self.off["actor_mlp.0.weight"]                      # [0, bucket0): discriminator weights
        z = lambda: torch.zeros(n, device=dev, dtype=torch.float32)
        if symmetric:
            # parameters and gradients in symmetric memory: every rank maps every rank's buffers (and one multicast address for
            # all of them) - the fused optimiser step of `NvlsShardedAdam` reads / writes them through the NVSwitch
            import torch.distributed._symmetric_memory as symm
            self.p, self.g = (symm.empty(n, dtype=torch.float32, device=dev).zero_() for _ in range(2))
            self.m = self.v = None                                         # moments are sharded (NvlsShardedAdam)
        else:
            self.p, self.g, self.m, self.v = z(), z(), z(), z()
        self.state = torch.zeros(2, device=dev, dtype=torch.float32)      # Adam step count, sum of squares of the gradient
        self.partials = torch.zeros(148 * 8, device=dev, dtype=torch.float32)
        for k in self.GEMM + self.TAIL:
            q = named[k]
            view = self.p[self.off[k]:self.off[k] + q.numel()].view_as(q)
            view.copy_(q.detach())
            q.data = view
        self.named = named
        self.net = net

    def param(self, k):
        return self.named[k]

    def grad(self, k, rows=None):
        q = self.named[k]
        return self.g[self.off[k]:self.off[k] + q.numel()].view_as(q)

    def stacked(self, a, b, what):
        """Two adjacent tensors as one view (actor + critic first layer)."""
        buf = self.p if what == "p" else self.g
        qa, qb = self.named[a], self.named[b]
        assert self.off[b] == self.off[a] + qa.numel(), "stacked tensors must be adjacent"
        shape = (qa.shape[0] + qb.shape[0],) + tuple(qa.shape[1:])
        return buf[self.off[a]:self.off[a] + qa.numel() + qb.numel()].view(shape)

    def zero_tail(self):
        self.g[self.tail_start:].zero_()


class NvlsShardedAdam:
    """The data-parallel optimiser step fused with its collective over NVLink / NVSwitch multicast memory
    (`emloco_dp_reduce_shard` + `emloco_dp_adam_shard`, csrc/update.cu): rank r reduces ITS 1/W slice of the flat gradient inside
    the switch (`multimem.ld_reduce`), the slices' sums of squares are exchanged through a W-slot multicast buffer, the rank
    clips and runs Adam on its slice with its shard of the moments, and broadcasts the new parameters to every rank
    (`multimem.st`).  Three stream-ordered symmetric-memory barriers order the ranks; no NCCL call, no host synchronisation.
    Every rank ends up with bit-identical parameters (each element is reduced and updated exactly once, by its owner)."""

    def __init__(self, flat: "FlatParams", group=None):
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        self.flat, self.world, self.rank = flat, dist.get_world_size(group), dist.get_rank(group)
        dev = flat.p.device
        self.hg, self.hp = symm.rendezvous(flat.g, group), symm.rendezvous(flat.p, group)
        self.x = symm.empty(4 * self.world, dtype=torch.float32, device=dev).zero_()
        self.hx = symm.rendezvous(self.x, group)
        if not (self.hg.multicast_ptr and self.hp.multicast_ptr and self.hx.multicast_ptr):
            raise _lib.EmlocoError("NvlsShardedAdam needs NVSwitch multicast (NVLS) support; use reducer='nccl'")
        per = -(-flat.n // (4 * self.world)) * 4
        self.lo = min(self.rank * per, flat.n)
        self.count = max(0, min(flat.n, self.lo + per) - self.lo)
        z = lambda: torch.zeros(max(self.count, 4), device=dev, dtype=torch.float32)
        self.shard, self.m, self.v = z(), z(), z()
        self.hg.barrier(channel=0)

    def step(self, lr, beta1, beta2, eps, max_norm):
        FP, lib = self.flat, _lib.load()
        p = lambda a: C.c_void_p(int(a))
        self.hg.barrier(channel=0)                                        # every rank's backward pass has been issued and finished
        _lib.check(lib.emloco_dp_reduce_shard(p(self.hg.multicast_ptr), _ptr(self.shard), self.lo, self.count, _ptr(FP.partials),
                                              p(self.hx.multicast_ptr), self.rank, _stream()), "emloco_dp_reduce_shard")
        self.hg.barrier(channel=1)                                        # all slices reduced (gradients may be overwritten), all slots published
        _lib.check(lib.emloco_dp_adam_shard(p(self.hp.multicast_ptr), _ptr(FP.p), _ptr(self.shard), _ptr(self.m), _ptr(self.v), self.lo, self.count,
                                            _ptr(self.x), self.world, _ptr(FP.state), lr, beta1, beta2, eps, max_norm, 1.0 / self.world, _stream()),
                   "emloco_dp_adam_shard")
        self.hg.barrier(channel=2)                                        # every rank's parameter buffer holds all slices


class PPOUpdate:
    """One optimiser step per `step(batch)` for fixed minibatch sizes B (policy rows) and Ba (AMP rows per source).
    reducer: how the ranks' gradients meet - "nccl" (bucketed all-reduce, then norm + Adam on every rank), "nvls" (the fused
    NVSwitch-multicast step of `NvlsShardedAdam`) or "auto" (nvls on more than one rank when multicast is available)."""

    def __init__(self, net: AMPSeptValueNetwork, obs_norm: RunningMeanStd, amp_norm: RunningMeanStd, B, Ba, cfg=None, world=None,
                 overlap_allreduce=True, reducer="auto"):
        self.cfg = dict(DEFAULT_CFG, **(cfg or {}))
        self.net, self.obs_norm, self.amp_norm, self.B, self.Ba = net, obs_norm, amp_norm, int(B), int(Ba)
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else int(world)
        assert reducer in ("auto", "nccl", "nvls")
        use_nvls = self.world > 1 and reducer in ("auto", "nvls") and dist.is_initialized() and dist.get_backend() == "nccl"
        self.nvls = None
        if use_nvls:
            try:
                self.flat = FlatParams(net, symmetric=True)
                self.nvls = NvlsShardedAdam(self.flat)
            except Exception:
                if reducer == "nvls":
                    raise
                self.flat, self.nvls = None, None
        if self.nvls is None:
            self.flat = FlatParams(net)
        self.reducer_name = "nvls" if self.nvls is not None else ("nccl" if self.world > 1 else "none")
        dev = self.flat.p.device
        self.dev = dev
        self.overlap = bool(overlap_allreduce)
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        S = lambda rows, k: _Split(rows, k, dev)
        B, Ba = self.B, self.Ba
        n = net
        self.h = h = n.actor_mlp[0].out_features
        a2, t1, t2 = n.actor_mlp[2].out_features, n._task_mlp[0].out_features, n._task_mlp[2].out_features
        d1, d2 = n._disc_mlp[0].out_features, n._disc_mlp[2].out_features
        v1, v2 = n._task_value_mlp[0].out_features, n._task_value_mlp[2].out_features
        ain = SELF_OBS + t2
        self.dims = dict(a2=a2, t1=t1, t2=t2, d1=d1, d2=d2, v1=v1, v2=v2, ain=ain)
        # ---- weights as GEMM operands: row-major split (forward, and the "W itself" products of the gradient penalty) and
        # transposed split (dgrad), refreshed from the flat parameters after every optimiser step ----
        self.W, self.WT = {}, {}
        for k in ("_task_mlp.0", "_task_mlp.2", "ac0", "actor_mlp.2", "critic_mlp.2", "mu", "_disc_mlp.0", "_disc_mlp.2", "_task_value_mlp.0",
                  "_task_value_mlp.2", "_value_logits"):
            w = self._weight(k)
            self.W[k] = S(w.shape[0], w.shape[1])
            self.WT[k] = S(w.shape[1], w.shape[0])
        # ---- policy-side activations ----
        self.s_tin, self.tinT = S(B, TASK_OBS), S(TASK_OBS, B)
        self.s_ain, self.ainT = S(B, ain), S(ain, B)
        self.t1_32, self.s_t1, self.t1T = f(B, t1), S(B, t1), S(t1, B)
        self.t2_32 = f(B, t2)
        self.ac1_32, self.s_ac1, self.ac1T = f(B, 2 * h), S(B, 2 * h), S(2 * h, B)
        self.a2_32, self.s_a2, self.a2T = f(B, a2), S(B, a2), S(a2, B)
        self.c2_32 = f(B, a2)
        self.hp_c = f(B, (a2 + 63) // 64)
        self.mu32, self.value, self.tv = f(B, ACTIONS), f(B, 1), f(B, 1)
        self.v1_32, self.s_v1, self.v1T = f(B, v1), S(B, v1), S(v1, B)
        self.v2_32, self.s_v2 = f(B, v2), S(B, v2)
        # gradients flowing back
        self.dmu32, self.s_dmu, self.dmuT = f(B, ACTIONS), S(B, ACTIONS), S(ACTIONS, B)
        self.dvalue, self.dtv = f(B), f(B)
        self.da2_32, self.s_da2, self.da2T = f(B, a2), S(B, a2), S(a2, B)
        self.s_dc2, self.dc2T = S(B, a2), S(a2, B)
        self.dac1_32, self.s_dac1, self.dac1T = f(B, 2 * h), S(B, 2 * h), S(2 * h, B)
        self.dain_32 = f(B, ain)
        self.s_dt2, self.dt2T = S(B, t2), S(t2, B)
        self.dt1_32, self.dt1T = f(B, t1), S(t1, B)
        self.s_dv2, self.dv2T = S(B, v2), S(v2, B)
        self.dv1_32, self.dv1T = f(B, v1), S(v1, B)
        # ---- discriminator side: rows = [agent | replay | demo] (3 Ba); the wgrad operands carry Ba extra columns so that the
        # gradient-penalty term is accumulated by the same GEMM (contraction over 3 Ba + Ba columns) ----
        R3, R4 = 3 * Ba, 4 * Ba
        self.u = f(R3, 19)                                    # per-joint uniform draws of the dropout gate (no mask tensor)
        self.s_xa, self.xaT = S(R3, AMP_OBS), S(AMP_OBS, R4)
        self.h1_32, self.s_h1, self.h1T = f(R3, d1), S(R3, d1), S(d1, R4)
        self.h2_32 = f(R3, d2)
        self.hp_d = f(R3, (d2 + 63) // 64)
        self.logit, self.dlogit = f(R3, 1), f(R3)
        self.s_dh2, self.dh2T = S(R3, d2), S(d2, R4)
        self.dh1_32, self.dh1T = f(R3, d1), S(d1, R4)
        # gradient penalty (demo rows only)
        self.s_u2 = S(Ba, d2)
        self.v1g_32, self.s_u1 = f(Ba, d1), S(Ba, d1)
        self.gx_32, self.s_e0 = f(Ba, AMP_OBS), S(Ba, AMP_OBS)
        self.du1_32, self.s_dv1g = f(Ba, d1), S(Ba, d1)
        self.du2_32 = f(Ba, d2)
        self.stats = torch.zeros(16, device=dev)            # [0:7] PPO sums, [8:12] disc sums, [12] sum g^2 of the penalty
        self.rms_scratch = torch.zeros(2 * AMP_OBS, device=dev, dtype=torch.float64)
        self._ws, self._e0_scale = None, 1.0
        self.split_k = True                 # weight gradients with few output tiles: split the contraction (the minibatch) over CTAs
        from .dist import BucketedAllReduce
        self.reducer = BucketedAllReduce(self.flat.g, self.flat.bucket0, overlap=self.overlap, world=self.world)
        self.refresh_weights()

    # ---- parameters -------------------------------------------------------------------------------------------------------
    def _weight(self, k):
        if k == "ac0":
            return self.flat.stacked("actor_mlp.0.weight", "critic_mlp.0.weight", "p")
        return self.flat.param(k + ".weight")

    def _bias(self, k):
        if k == "ac0":
            return self.flat.stacked("actor_mlp.0.bias", "critic_mlp.0.bias", "p")
        return self.flat.param(k + ".bias")

    def refresh_weights(self):
        """bf16 hi/lo splits of every weight in both orientations, from the flat fp32 parameters (one launch per weight)."""
        for k in self.W:
            xform(x=self._weight(k).detach(), split=self.W[k], splitT=self.WT[k])

    def adopt_into(self, nets):
        """Makes a `RolloutNets` read THIS object's weight splits and stacked first layer (same memory, refreshed by every
        `step`): the rollout that follows an update needs no re-split, and its CUDA graphs keep valid pointers."""
        m = {"t0": "_task_mlp.0", "t2": "_task_mlp.2", "ac1": "ac0", "a2": "actor_mlp.2", "mu": "mu", "c2": "critic_mlp.2",
             "d0": "_disc_mlp.0", "d2": "_disc_mlp.2"}
        for name, k in m.items():
            nets.w16.adopt(name, self.W[k])
        h = self.h
        nets.w16.adopt("c0", self.W["ac0"].rows_view(h, h))
        nets._stacked = ["adopted", self._weight("ac0").detach(), self._bias("ac0").detach()]

    # ---- one optimiser step -----------------------------------------------------------------------------------------------
    def step(self, batch, dropout_u=None):
        """batch: dict of CUDA float32 tensors - obs [B,1422], actions [B,69], old_logp_actions [B], advantages [B], returns [B,1],
        old_values [B,1] (unused: clip_value False), mu / sigma [B,69] (kl only), amp_obs / amp_obs_replay / amp_obs_demo [Ba,3090].
        dropout_u: optional [19, Ba, 3] uniform draws (tests replay the reference's); drawn on the device otherwise.
        Returns nothing; `info()` reads the loss terms of the last step back."""
        self.forward_backward(batch, dropout_u)
        self.reduce_and_apply()

    def forward_backward(self, batch, dropout_u=None):
        lib, cfg, n, B, Ba = _lib.load(), self.cfg, self.net, self.B, self.Ba
        h, D = self.h, self.dims
        FP, W, WT = self.flat, self.W, self.WT
        lin = linear_bf16x3
        b_ = lambda k: self._bias(k).detach()
        obs = batch["obs"]
        assert obs.shape == (B, OBS) and obs.is_contiguous()
        self.stats.zero_()
        FP.zero_tail()
        lib_ = lib
        _lib.check(lib_.emloco_adam_begin(_ptr(FP.state), _stream()), "emloco_adam_begin")

        # ---- input normalisation in training mode: normalise with the current statistics, then absorb the batch ----
        om, ov = self.obs_norm.f32()
        xform(x=obs[:, :SELF_OBS], mean=om[:SELF_OBS], var=ov[:SELF_OBS], split=self.s_ain.cols(0, SELF_OBS), splitT=self.ainT.rows_view(SELF_OBS))
        xform(x=obs[:, SELF_OBS:], mean=om[SELF_OBS:], var=ov[SELF_OBS:], split=self.s_tin, splitT=self.tinT)
        self._rms_update(self.obs_norm, obs)
        am, av = self.amp_norm.f32()
        R3 = 3 * Ba
        if cfg["amp_dropout"]:
            if dropout_u is None:
                self.u.uniform_()
            else:                                   # reference layout [19, Ba, 3] -> rows (source, sample), columns joints
                self.u.copy_(dropout_u.permute(2, 1, 0).reshape(R3, 19))
        for i, k in enumerate(("amp_obs", "amp_obs_replay", "amp_obs_demo")):     # three calls, each sees the previous update
            a = batch[k]
            assert a.shape == (Ba, AMP_OBS) and a.is_contiguous()
            rows = slice(i * Ba, (i + 1) * Ba)
            xform(x=a, mean=am, var=av, drop_u=self.u[rows] if cfg["amp_dropout"] else None, drop_rate=cfg["dropout_rate"],
                  split=self.s_xa.rows_view(Ba, i * Ba), splitT=self.xaT.cols(i * Ba, (i + 1) * Ba))
            self._rms_update(self.amp_norm, a)

        # ---- forward, policy side ----
        lin(self.s_tin, W["_task_mlp.0"], b_("_task_mlp.0"), True, y32=self.t1_32, y16=self.s_t1)
        xform(x=self.t1_32, splitT=self.t1T)
        lin(self.s_t1, W["_task_mlp.2"], b_("_task_mlp.2"), True, y32=self.t2_32, y16=self.s_ain.cols(SELF_OBS, D["ain"]))
        xform(x=self.t2_32, splitT=self.ainT.rows_view(D["t2"], SELF_OBS))
        lin(self.s_ain, W["ac0"], b_("ac0"), True, y32=self.ac1_32, y16=self.s_ac1)
        xform(x=self.ac1_32, splitT=self.ac1T)
        lin(self.s_ac1.cols(0, h), W["actor_mlp.2"], b_("actor_mlp.2"), True, y32=self.a2_32, y16=self.s_a2)
        xform(x=self.a2_32, splitT=self.a2T)
        lin(self.s_a2, W["mu"], b_("mu"), False, y32=self.mu32)
        lin(self.s_ac1.cols(h, 2 * h), W["critic_mlp.2"], b_("critic_mlp.2"), True, y32=self.c2_32, head=(n.value, self.value, self.hp_c))
        lin(self.s_tin.cols(0, TRAJ_OBS), W["_task_value_mlp.0"], b_("_task_value_mlp.0"), True, y32=self.v1_32)
        xform(x=self.v1_32, split=self.s_v1, splitT=self.v1T)
        lin(self.s_v1, W["_task_value_mlp.2"], b_("_task_value_mlp.2"), True, y32=self.v2_32)
        xform(x=self.v2_32, split=self.s_v2)
        lin(self.s_v2, W["_value_logits"], b_("_value_logits"), False, y32=self.tv)
        # ---- forward, discriminator ----
        lin(self.s_xa, W["_disc_mlp.0"], b_("_disc_mlp.0"), True, y32=self.h1_32, y16=self.s_h1)
        xform(x=self.h1_32, splitT=self.h1T.cols(0, R3))
        lin(self.s_h1, W["_disc_mlp.2"], b_("_disc_mlp.2"), True, y32=self.h2_32, head=(n._disc_logits, self.logit, self.hp_d))

        # ---- loss heads ----
        _lib.check(lib_.emloco_ppo_heads(
            _ptr(self.mu32), ACTIONS, _ptr(n.sigma), _ptr(batch["actions"]), _ptr(batch["old_logp_actions"]), _ptr(batch["advantages"]),
            _ptr(self.value), _ptr(self.tv), _ptr(batch["returns"]), _ptr(batch.get("mu")), _ptr(batch.get("sigma")), _ptr(self.dmu32), ACTIONS,
            _ptr(self.dvalue), _ptr(self.dtv), _ptr(self.stats), B, ACTIONS, cfg["e_clip"], cfg["actor_coef"], cfg["critic_coef"], cfg["tv_coef"],
            cfg["bounds_loss_coef"], _stream()), "emloco_ppo_heads")
        _lib.check(lib_.emloco_disc_heads(_ptr(self.logit), _ptr(self.dlogit), _ptr(self.stats[8:]), 2 * Ba, Ba, cfg["disc_coef"], _stream()),
                   "emloco_disc_heads")

        # ---- backward, discriminator first (its gradients form the first all-reduce bucket) ----
        G = FP.grad
        w3 = n._disc_logits.weight.detach().reshape(-1)
        # prediction loss
        xform(x=self.h2_32, rowscale=self.dlogit, colsum=G("_disc_logits.weight").reshape(-1))                       # d w3
        xform(x=self.dlogit.view(R3, 1), colsum=G("_disc_logits.bias"))
        xform(rowvec=w3, rowscale=self.dlogit, gate=self.h2_32, M=R3, K=D["d2"], split=self.s_dh2, splitT=self.dh2T.cols(0, R3),
              colsum=G("_disc_mlp.2.bias"))
        lin(self.s_dh2, WT["_disc_mlp.2"], None, False, y32=self.dh1_32)
        xform(x=self.dh1_32, gate=self.h1_32, splitT=self.dh1T.cols(0, R3), colsum=G("_disc_mlp.0.bias"))
        # gradient penalty on the demo rows: g = mask * (((w3 * m2) W2 * m1) W1); masks are constants of the second backward
        demo = slice(2 * Ba, R3)
        h1d, h2d = self.h1_32[demo], self.h2_32[demo]
        cgp = cfg["disc_coef"] * cfg["disc_grad_penalty"]
        xform(rowvec=w3, gate=h2d, M=Ba, K=D["d2"], split=self.s_u2, splitT=self.dh2T.cols(R3, R3 + Ba))             # u2 (and u2^T: wgrad of W2)
        lin(self.s_u2, WT["_disc_mlp.2"], None, False, y32=self.v1g_32)                                               # u2 W2
        xform(x=self.v1g_32, gate=h1d, split=self.s_u1, splitT=self.dh1T.cols(R3, R3 + Ba))                           # u1 (and u1^T: wgrad of W1)
        lin(self.s_u1, WT["_disc_mlp.0"], None, False, y32=self.gx_32)                                                # u1 W1
        # e0 = d penalty / d gx = (2 cgp / Ba) * mask * gx; its sum of squares gives the penalty itself (info() undoes the scale)
        self._e0_scale = 2.0 * cgp / Ba
        xform(x=self.gx_32, drop_u=self.u[demo] if cfg["amp_dropout"] else None, drop_rate=cfg["dropout_rate"], scale=self._e0_scale,
              split=self.s_e0, splitT=self.xaT.cols(R3, R3 + Ba), sumsq=self.stats[12:13])
        lin(self.s_e0, W["_disc_mlp.0"], None, False, y32=self.du1_32)                                                # e0 W1^T
        xform(x=self.du1_32, gate=h1d, split=self.s_dv1g, splitT=self.h1T.cols(R3, R3 + Ba))                          # d / d (u2 W2)
        lin(self.s_dv1g, W["_disc_mlp.2"], None, False, y32=self.du2_32)                                              # dv1g W2^T
        xform(x=self.du2_32, gate=h2d, colsum=G("_disc_logits.weight").reshape(-1))
        # weight gradients: prediction-loss and penalty terms in ONE product each (contraction over 3 Ba + Ba columns)
        lin(self.dh1T, self.xaT, None, False, y32=G("_disc_mlp.0.weight"))
        self._wgrad(self.dh2T, self.h1T, G("_disc_mlp.2.weight"))
        # regularisers: logit_reg * |w3|^2 + weight_decay * (|W1|^2 + |W2|^2 + |w3|^2), times disc_coef
        dc = cfg["disc_coef"]
        self._axpy(G("_disc_logits.weight"), n._disc_logits.weight, 2.0 * dc * (cfg["disc_logit_reg"] + cfg["disc_weight_decay"]))
        self._axpy(G("_disc_mlp.0.weight"), n._disc_mlp[0].weight, 2.0 * dc * cfg["disc_weight_decay"])
        self._axpy(G("_disc_mlp.2.weight"), n._disc_mlp[2].weight, 2.0 * dc * cfg["disc_weight_decay"])
        # the discriminator's weight gradients (33 % of all parameters) are final: their all-reduce runs under the actor /
        # critic backward that follows
        if self.nvls is None:
            self.reducer.start_first()

        # ---- backward, actor / critic / task trunk ----
        xform(x=self.dmu32, split=self.s_dmu, splitT=self.dmuT, colsum=G("mu.bias"))
        lin(self.s_dmu, WT["mu"], None, False, y32=self.da2_32)
        self._wgrad(self.dmuT, self.a2T, G("mu.weight"))
        xform(x=self.da2_32, gate=self.a2_32, split=self.s_da2, splitT=self.da2T, colsum=G("actor_mlp.2.bias"))
        lin(self.s_da2, WT["actor_mlp.2"], None, False, y32=self.dac1_32[:, :h])
        lin(self.da2T, self.ac1T.rows_view(h, 0), None, False, y32=G("actor_mlp.2.weight"))
        wv = n.value.weight.detach().reshape(-1)
        xform(x=self.c2_32, rowscale=self.dvalue, colsum=G("value.weight").reshape(-1))
        xform(x=self.dvalue.view(B, 1), colsum=G("value.bias"))
        xform(rowvec=wv, rowscale=self.dvalue, gate=self.c2_32, M=B, K=D["a2"], split=self.s_dc2, splitT=self.dc2T, colsum=G("critic_mlp.2.bias"))
        lin(self.s_dc2, WT["critic_mlp.2"], None, False, y32=self.dac1_32[:, h:])
        lin(self.dc2T, self.ac1T.rows_view(h, h), None, False, y32=G("critic_mlp.2.weight"))
        xform(x=self.dac1_32, gate=self.ac1_32, split=self.s_dac1, splitT=self.dac1T, colsum=FP.stacked("actor_mlp.0.bias", "critic_mlp.0.bias", "g"))
        lin(self.s_dac1, WT["ac0"], None, False, y32=self.dain_32)
        lin(self.dac1T, self.ainT, None, False, y32=FP.stacked("actor_mlp.0.weight", "critic_mlp.0.weight", "g"))
        xform(x=self.dain_32[:, SELF_OBS:], gate=self.t2_32, split=self.s_dt2, splitT=self.dt2T, colsum=G("_task_mlp.2.bias"))
        lin(self.s_dt2, WT["_task_mlp.2"], None, False, y32=self.dt1_32)
        self._wgrad(self.dt2T, self.t1T, G("_task_mlp.2.weight"))
        xform(x=self.dt1_32, gate=self.t1_32, splitT=self.dt1T, colsum=G("_task_mlp.0.bias"))
        self._wgrad(self.dt1T, self.tinT, G("_task_mlp.0.weight"))
        # task-value MLP
        wl = n._value_logits.weight.detach().reshape(-1)
        xform(x=self.v2_32, rowscale=self.dtv, colsum=G("_value_logits.weight").reshape(-1))
        xform(x=self.dtv.view(B, 1), colsum=G("_value_logits.bias"))
        xform(rowvec=wl, rowscale=self.dtv, gate=self.v2_32, M=B, K=D["v2"], split=self.s_dv2, splitT=self.dv2T, colsum=G("_task_value_mlp.2.bias"))
        lin(self.s_dv2, WT["_task_value_mlp.2"], None, False, y32=self.dv1_32)
        self._wgrad(self.dv2T, self.v1T, G("_task_value_mlp.2.weight"))
        xform(x=self.dv1_32, gate=self.v1_32, splitT=self.dv1T, colsum=G("_task_value_mlp.0.bias"))
        self._wgrad(self.dv1T, self.tinT.rows_view(TRAJ_OBS, 0), G("_task_value_mlp.0.weight"))

    def reduce_and_apply(self):
        """Gradient average over ranks (one all-reduce of the flat buffer), clip-norm, Adam, fresh operand splits."""
        FP, cfg, lib = self.flat, self.cfg, _lib.load()
        if self.nvls is not None:                                      # collective and optimiser in one pass over NVLink multicast memory
            self.nvls.step(cfg["lr"], 0.9, 0.999, 1e-8, cfg["grad_norm"])
            self.refresh_weights()
            return
        self.reducer.finish()                                          # summed; the 1 / world factor is folded into the Adam kernel
        _lib.check(lib.emloco_grad_sumsq(_ptr(FP.g), FP.n, _ptr(FP.state), _ptr(FP.partials), _stream()), "emloco_grad_sumsq")
        _lib.check(lib.emloco_adam_clip(_ptr(FP.p), _ptr(FP.g), _ptr(FP.m), _ptr(FP.v), FP.n, _ptr(FP.state), cfg["lr"], 0.9, 0.999, 1e-8,
                                        cfg["grad_norm"], 1.0 / self.world, _stream()), "emloco_adam_clip")
        self.refresh_weights()

    def _wgrad(self, aT, xT, out):
        """out[N_out, K_in] = aT [N_out, rows] . xT [K_in, rows]^T.  When the output has too few 128 x 128 tiles to fill the 148
        SMs the contraction (the minibatch rows) is split over up to 15 CTAs per tile and the partial matrices are added in
        split order (deterministic)."""
        M, N = out.shape
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        splits = min(15, 128 // tiles) if (self.split_k and tiles <= 48) else 1
        while splits > 1:                   # every split must own at least one k-block, for both k-block sizes of the kernel
            if all((splits - 1) * (-(-(-(-aT.K // bk)) // splits)) < -(-aT.K // bk) for bk in (64, 32)):
                break
            splits -= 1
        if splits <= 1:
            linear_bf16x3(aT, xT, None, False, y32=out)
            return
        need = splits * M * N
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=self.dev, dtype=torch.float32)
        parts = self._ws[:need].view(splits, M, N)
        linear_bf16x3(aT, xT, None, False, y32=parts[0], splits=splits)
        _lib.check(_lib.load().emloco_sum_parts(_ptr(parts), splits, M * N, _ptr(out), M * N, 0, _stream()), "emloco_sum_parts")

    def dropout_mask(self):
        """The dropout gates of the last step as a [3 Ba, 3090] 0/1 tensor (tests; the step itself never materialises it)."""
        m = torch.empty(3 * self.Ba, AMP_OBS, device=self.dev, dtype=torch.float32)
        _lib.check(_lib.load().emloco_amp_dropout_mask(_ptr(self.u), _ptr(m), 3 * self.Ba, self.cfg["dropout_rate"], _stream()), "emloco_amp_dropout_mask")
        return m

    def _axpy(self, y, x, a):
        _lib.check(_lib.load().emloco_axpy(_ptr(y), _ptr(x.detach()), float(a), y.numel(), _stream()), "emloco_axpy")

    def _rms_update(self, norm: RunningMeanStd, x):
        norm.f32()
        c = norm._f32
        K = x.shape[1]
        _lib.check(_lib.load().emloco_rms_update(_ptr(x), x.stride(0), x.shape[0], K, _ptr(self.rms_scratch), _ptr(norm.running_mean),
                                                 _ptr(norm.running_var), _ptr(norm.count), _ptr(c[1]), _ptr(c[2]), _ptr(c[3]), norm.epsilon,
                                                 _stream()), "emloco_rms_update")

    def info(self):
        """Loss terms of the last step (one device->host read): the `train_result` entries of calc_gradients (:408-423)."""
        s = self.stats.tolist()
        total_sq, cfg, B, Ba = float(self.flat.state[1].item()), self.cfg, self.B, self.Ba
        n = self.net
        w3 = n._disc_logits.weight.detach()
        logit_reg = float((w3 ** 2).sum().item())
        a_loss, c_loss, tv_loss, b_loss = s[0] / B, s[1] / B, s[2] / B, s[3] / B
        pred = 0.5 * (s[8] / (2 * Ba) + s[9] / Ba)
        gp = s[12] / (self._e0_scale ** 2) / Ba
        return dict(a_loss=a_loss, c_loss=c_loss, tv_loss=tv_loss, b_loss=b_loss, a_clip_frac=s[4] / B, kl=s[5] / B, entropy=s[6] / B,
                    disc_pred_loss=pred, disc_grad_penalty=gp, disc_logit_loss=logit_reg, disc_agent_acc=s[10] / (2 * Ba), disc_demo_acc=s[11] / Ba,
                    total_norm=(total_sq ** 0.5) / self.world)
