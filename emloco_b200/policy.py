"""Host-side mirror of the reference's rollout networks, evaluated by the kernels of libemloco_b200.so.

Reference: ``AMPSeptValueBuilder.Network`` (pacer/pacer/learning/amp_network_sept_value_builder.py:19-90 ->
amp_network_sept_builder.py:22-127 -> amp_network_builder.py:16-121) wrapped by ``ModelAMPContinuousSeptValue``
(amp_sept_value_models.py:22-30) with the input normaliser of utils/running_mean_std.py:60-98, as configured by
data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml:12-70.  Parameter names follow the reference's
``a2c_network.*`` state-dict keys so an rl_games checkpoint's ``model`` dict loads with ``load_state_dict``.

PyTorch holds the parameters and the workspace; every multiply-add, normalisation, activation and the Gaussian
head runs in this repo's CUDA kernels through the C ABI (``emloco_linear``, ``emloco_normalize``,
``emloco_sample_actions``).  There is no torch fallback: without the library or a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from .sim import _ptr, _stream, linear

SELF_OBS, TASK_OBS, TRAJ_OBS, OBS, AMP_OBS, ACTIONS = 368, 1054, 30, 1422, 3090, 69


def _mlp(in_size, units):
    layers, k = [], in_size
    for u in units:
        layers += [nn.Linear(k, u), nn.ReLU()]
        k = u
    return nn.Sequential(*layers)


class RunningMeanStd(nn.Module):
    """State holder with the reference's buffer names (running_mean/running_var/count, float64)."""

    def __init__(self, size, epsilon=1e-5):
        super().__init__()
        self.epsilon = epsilon
        self.register_buffer("running_mean", torch.zeros(size, dtype=torch.float64))
        self.register_buffer("running_var", torch.ones(size, dtype=torch.float64))
        self.register_buffer("count", torch.ones((), dtype=torch.float64))
        self._f32 = None

    def f32(self):
        """(mean, var) as float32 - `current_var.float()` of running_mean_std.py:78-84.  The three fp32 copies (mean, var,
        1/sqrt(var+eps)) live in buffers of fixed address that are REFRESHED IN PLACE when the statistics change - by an
        in-place update (`_version`) or by attribute re-assignment, which is how the reference updates them
        (running_mean_std.py:93-96: fresh tensors, `_version` 0 again, hence `data_ptr` in the key).  Kernels captured in a
        CUDA graph therefore read current statistics after `refresh()` / `Rollout.sync_weights()`."""
        m, v = self.running_mean, self.running_var
        key = (m.data_ptr(), m._version, v.data_ptr(), v._version)
        c = self._f32
        if c is None or c[1].device != m.device or c[1].shape != m.shape:
            var = v.float().contiguous()
            c = self._f32 = [key, m.float().contiguous(), var, (1.0 / torch.sqrt(var + self.epsilon)).contiguous()]
        elif c[0] != key:
            c[1].copy_(m); c[2].copy_(v)
            torch.rsqrt(c[2] + self.epsilon, out=c[3])
            c[0] = key
        return c[1], c[2]

    refresh = f32

    def inv_std(self):
        """1/sqrt(var + eps) as float32, for the kernels that multiply instead of dividing (post-step operand sinks)."""
        self.f32()
        return self._f32[3]

    def update(self, x):
        """Training-mode moment update of RunningMeanStd.forward (utils/running_mean_std.py:86-96 ->
        _update_mean_var_count_from_moments :33-43) in float64, written IN PLACE.  x [B, size] (any float dtype, same device)."""
        x = x.reshape(-1, self.running_mean.numel() if self.running_mean.dim() else 1)
        b = x.shape[0]
        mean, var = x.mean(0).reshape(self.running_mean.shape), x.var(0).reshape(self.running_var.shape)
        delta = mean - self.running_mean                       # type promotion as in the reference: batch moments stay in the
        tot = self.count + b                                   # input's dtype until they meet the float64 buffers
        new_mean = self.running_mean + delta * b / tot
        m2 = self.running_var * self.count + var * b + delta ** 2 * self.count * b / tot
        self.running_mean.copy_(new_mean); self.running_var.copy_(m2 / tot); self.count.copy_(tot)


class AMPSeptValueNetwork(nn.Module):
    """The `a2c_network` of the reference with the default cfg: shared task MLP 1054->512->256, actor and critic
    624->2048->1024 (+mu 69 / value 1), task-value 30->15->6->1, discriminator 3090->1024->512->1, fixed logstd -2.9."""

    def __init__(self, mlp_units=(2048, 1024), task_units=(512, 256), value_units=(15, 6), disc_units=(1024, 512),
                 sigma_init=-2.9):
        super().__init__()
        ain = SELF_OBS + task_units[-1]
        self.actor_mlp = _mlp(ain, mlp_units)
        self.critic_mlp = _mlp(ain, mlp_units)
        self.mu = nn.Linear(mlp_units[-1], ACTIONS)
        self.sigma = nn.Parameter(torch.full((ACTIONS,), float(sigma_init)), requires_grad=False)
        self.value = nn.Linear(mlp_units[-1], 1)
        self._task_mlp = _mlp(TASK_OBS, task_units)
        self._task_value_mlp = _mlp(TRAJ_OBS, value_units)
        self._value_logits = nn.Linear(value_units[-1], 1)
        self._disc_mlp = _mlp(AMP_OBS, disc_units)
        self._disc_logits = nn.Linear(disc_units[-1], 1)
        for m in list(self._disc_mlp) + list(self._task_value_mlp):
            if isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)                                     # amp_network_builder.py:107-111
        nn.init.uniform_(self._disc_logits.weight, -1.0, 1.0)              # :113-114
        nn.init.zeros_(self._disc_logits.bias)
        nn.init.uniform_(self._value_logits.weight, -1.0, 1.0)             # amp_network_sept_value_builder.py:86-87
        nn.init.zeros_(self._value_logits.bias)


class Fork:
    """Runs independent launch groups on side streams and joins them: parallel branches of the CUDA graph under capture,
    plain stream concurrency otherwise.  Small layers (heads, task-value MLP, LocoVal) occupy a fraction of the 148 SMs;
    branches let them overlap instead of serialising."""

    def __init__(self, device, n=3, priority=0):
        # priority -1: kernels of the side branches are dispatched ahead of the main branch's when both are pending (the
        # priority of the capturing stream is recorded into the graph's kernel nodes)
        self.streams = [torch.cuda.Stream(device=device, priority=priority) for _ in range(n)]

    def run(self, *fns):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        side = self.streams[:len(fns) - 1]
        for st, fn in zip(side, fns[1:]):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                fn()
        fns[0]()
        for st in side:
            cur.wait_stream(st)


class RolloutNets:
    """Evaluates the networks for a fixed row count with preallocated workspaces (CUDA-graph friendly)."""

    def __init__(self, net: AMPSeptValueNetwork, obs_norm: RunningMeanStd, amp_norm: RunningMeanStd, rows: int,
                 tensor_cores: bool = False, concurrent: bool = False, amp_slots: int = 1, chain: bool = False):
        self.net, self.obs_norm, self.amp_norm, self.M, self.tc = net, obs_norm, amp_norm, int(rows), bool(tensor_cores)
        # chain: the dense layers of a network pass run as ONE persistent launch (emloco_linear_chain) instead of one per layer
        self.chain = bool(chain) and self.tc
        self._chain_ws = {}
        dev = net.mu.weight.device
        if dev.type != "cuda":
            raise _lib.EmlocoError("RolloutNets needs its parameters on a CUDA device; there is no CPU fallback")
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        M = self.M
        a1, a2 = net.actor_mlp[0].out_features, net.actor_mlp[2].out_features
        t1, t2 = net._task_mlp[0].out_features, net._task_mlp[2].out_features
        d1, d2 = net._disc_mlp[0].out_features, net._disc_mlp[2].out_features
        v1, v2 = net._task_value_mlp[0].out_features, net._task_value_mlp[2].out_features
        self.t1, self.ain = f(M, t1), f(M, SELF_OBS + t2)
        self.ac1 = f(M, 2 * a1)                 # actor and critic first layers stacked: one GEMM over the shared input
        self.a2, self.c2 = f(M, a2), f(M, a2)
        self.mu, self.value, self.next_value, self.task_value = f(M, ACTIONS), f(M, 1), f(M, 1), f(M, 1)
        self.v1, self.v2 = f(M, v1), f(M, v2)
        self.d1, self.d2, self.logit = f(M, d1), f(M, d2), f(M, 1)
        self.actions, self.neglogp = f(M, ACTIONS), f(M)
        self._stacked = None
        self.fork = Fork(dev, priority=-1) if concurrent else None
        if self.tc:
            S = lambda k: _Split(M, k, dev)
            self.s_tin, self.s_t1, self.s_ain, self.s_ac1 = S(TASK_OBS), S(t1), S(SELF_OBS + t2), S(2 * a1)
            self.s_a2, self.s_d1 = S(a2), S(d1)      # (the last hidden layers of critic and discriminator are consumed in the epilogue: no buffer)
            # partial sums of the fused value / logit heads (one per 64 columns of the hidden layer that feeds them)
            self.hp_c, self.hp_d = f(M, (a2 + 63) // 64), f(M, (d2 + 63) // 64)
            # the mu layer (a2 -> 69) is 32 tiles of 16 k-blocks; split-K x4 (mu_splits = 4: 128 CTAs, partials added by the action
            # sampler) makes the layer itself faster but its CTAs then crowd out the critic branch's second wave: the policy pass
            # got 13 us SLOWER (202 -> 216 us), so it stays off
            self.mu_splits = 1
            self.mu_parts = f(self.mu_splits, M, ACTIONS)
            # discriminator operands of every step of the horizon are kept (slot n = rows [n*M, (n+1)*M)), so the post-horizon
            # discriminator pass of play_steps (:157) runs as ONE M*T-row GEMM chain without re-reading the fp32 AMP rows
            self.amp_slots = int(amp_slots)
            self.s_amp = _Split(M * self.amp_slots, AMP_OBS, dev)
            self._all = None
            self._cc = None
            self.w16 = _TcWeights()
            self._q = None          # second operand / activation set of the critic(next obs) pass (merged launches), made on demand
            self.value2 = f(M, 1)   # policy-pass value of odd steps (merged launches: step n-1's value is still needed during step n)

    def _w_ac1(self):
        """Actor and critic first layers stacked into one [2*h, 624] weight (buffers of fixed address, refreshed in place)."""
        n = self.net
        ps = (n.actor_mlp[0].weight, n.actor_mlp[0].bias, n.critic_mlp[0].weight, n.critic_mlp[0].bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        c = self._stacked
        if c is not None and c[0] == "adopted":       # views of a flat parameter buffer (update.PPOUpdate.adopt_into): always current
            return c[1], c[2]
        if c is None or c[1].device != ps[0].device:
            c = self._stacked = [key, torch.cat([ps[0].detach(), ps[2].detach()]).contiguous(),
                                 torch.cat([ps[1].detach(), ps[3].detach()]).contiguous()]
        elif c[0] != key:
            h = ps[0].shape[0]
            c[1][:h].copy_(ps[0].detach()); c[1][h:].copy_(ps[2].detach())
            c[2][:h].copy_(ps[1].detach()); c[2][h:].copy_(ps[3].detach())
            c[0] = key
        return c[1], c[2]

    def sync_weights(self):
        """Brings every derived copy the kernels read up to date with the parameters and normaliser statistics, IN PLACE (same
        device addresses): fp32 normaliser copies, the stacked actor / critic first layer, the bf16 hi / lo weight splits.
        Python does not run when a CUDA graph replays, so `Rollout` calls this eagerly before every horizon; afterwards the
        captured kernels read the new values through the old pointers.  -> True when a buffer had to be re-allocated (the
        caller must then drop its graphs)."""
        n = self.net
        self.obs_norm.f32(); self.amp_norm.f32()
        w, _ = self._w_ac1()
        if not self.tc:
            return False
        gen = self.w16.generation
        W = self.w16.get
        W("t0", n._task_mlp[0].weight); W("t2", n._task_mlp[2].weight); W("ac1", w); W("a2", n.actor_mlp[2].weight)
        W("mu", n.mu.weight); W("c2", n.critic_mlp[2].weight); W("c0", n.critic_mlp[0].weight)
        W("d0", n._disc_mlp[0].weight); W("d2", n._disc_mlp[2].weight)
        return self.w16.generation != gen

    def pointer_fingerprint(self):
        """Addresses of everything a captured graph reads through a baked-in pointer (parameters used directly by the kernels,
        derived copies).  A change means the graphs are stale and must be re-captured."""
        n = self.net
        direct = [n.sigma, n.value.weight, n.value.bias, n._disc_logits.weight, n._disc_logits.bias, n.mu.bias,
                  n._value_logits.weight, n._value_logits.bias] + [p for m in (n._task_value_mlp, n._task_mlp, n.actor_mlp, n.critic_mlp,
                                                                                 n._disc_mlp) for p in m.parameters()]
        fp = [p.data_ptr() for p in direct]
        fp += [t.data_ptr() for t in (self.obs_norm._f32 or [None])[1:]] + [t.data_ptr() for t in (self.amp_norm._f32 or [None])[1:]]
        if self._stacked is not None:
            fp += [self._stacked[1].data_ptr(), self._stacked[2].data_ptr()]
        if self.tc:
            fp.append(self.w16.generation)
        return tuple(fp)

    def next_obs_set(self):
        """Second operand / activation set: emloco_post_step writes the normalised next observation into BOTH sets; resets patch only
        the first, so `critic(next obs)` of step n can run after the reset of step n+1 - in the same launch as its policy pass."""
        if self._q is None:
            n, M, dev = self.net, self.M, self.mu.device
            S = lambda k: _Split(M, k, dev)
            t1, t2 = n._task_mlp[0].out_features, n._task_mlp[2].out_features
            self._q = dict(tin=S(TASK_OBS), t1=S(t1), ain=S(SELF_OBS + t2), c1=S(n.critic_mlp[0].out_features),
                           hp=torch.empty(M, (n.critic_mlp[2].out_features + 63) // 64, device=dev))
        return self._q

    def post_sinks(self, obs_copy=None, amp_copy=None, slot=0, flip_copy=None, rows_only=False, second=False):
        """emloco_post_sinks pointing at this object's operand buffers (tensor-core path) plus the given experience rows.
        rows_only: sim.flip_obs / sim.amp_obs are not refreshed, the experience rows are the only copies."""
        k = _lib.PostSinks()
        k.rows_only = int(bool(rows_only))
        k.obs_copy = None if obs_copy is None else obs_copy.data_ptr()
        k.amp_copy = None if amp_copy is None else amp_copy.data_ptr()
        k.flip_copy = None if flip_copy is None else flip_copy.data_ptr()
        if self.tc:
            om, _ = self.obs_norm.f32(); am, _ = self.amp_norm.f32()
            k.obs_mean, k.obs_inv_std = om.data_ptr(), self.obs_norm.inv_std().data_ptr()
            k.self_hi, k.self_lo, k.ld_self = self.s_ain.hi.data_ptr(), self.s_ain.lo.data_ptr(), self.s_ain.ld
            k.task_hi, k.task_lo, k.ld_task = self.s_tin.hi.data_ptr(), self.s_tin.lo.data_ptr(), self.s_tin.ld
            k.amp_mean, k.amp_inv_std = am.data_ptr(), self.amp_norm.inv_std().data_ptr()
            sa = self.s_amp.rows_view(self.M, slot * self.M)
            k.amp_hi, k.amp_lo, k.ld_amp = sa.hi.data_ptr(), sa.lo.data_ptr(), sa.ld
            if second:
                q = self.next_obs_set()
                assert q["ain"].ld == self.s_ain.ld and q["tin"].ld == self.s_tin.ld
                k.self_hi2, k.self_lo2, k.task_hi2, k.task_lo2 = (q["ain"].hi.data_ptr(), q["ain"].lo.data_ptr(), q["tin"].hi.data_ptr(),
                                                                  q["tin"].lo.data_ptr())
        return k

    def _lin(self, x, layer, relu, out, mean=None, var=None):
        return linear(x, layer.weight.detach(), layer.bias.detach(), relu=relu, mean=mean, var=var, out=out,
                      eps=self.obs_norm.epsilon)

    def _trunk(self, obs, operands_ready=False):
        """normalise -> task MLP -> [norm(self obs) | task_out] (amp_network_sept_builder.py:69-76,82-96).
        operands_ready: the post-step kernel already wrote the normalised bf16 operands (emloco_set_post_sinks)."""
        n = self.net
        mean, var = self.obs_norm.f32()
        if self.tc:
            W, eps = self.w16.get, self.obs_norm.epsilon
            if not operands_ready:
                split_bf16(obs[:, :SELF_OBS], self.s_ain.cols(0, SELF_OBS), mean[:SELF_OBS], var[:SELF_OBS], eps)
                split_bf16(obs[:, SELF_OBS:], self.s_tin, mean[SELF_OBS:], var[SELF_OBS:], eps)
            linear_bf16x3(self.s_tin, W("t0", n._task_mlp[0].weight), n._task_mlp[0].bias.detach(), True, y16=self.s_t1)
            linear_bf16x3(self.s_t1, W("t2", n._task_mlp[2].weight), n._task_mlp[2].bias.detach(), True,
                          y16=self.s_ain.cols(SELF_OBS, self.s_ain.K))
            return mean, var
        normalize(obs[:, :SELF_OBS], mean[:SELF_OBS], var[:SELF_OBS], self.obs_norm.epsilon, out=self.ain[:, :SELF_OBS])
        self._lin(obs[:, SELF_OBS:], n._task_mlp[0], True, self.t1, mean[SELF_OBS:], var[SELF_OBS:])
        self._lin(self.t1, n._task_mlp[2], True, self.ain[:, SELF_OBS:])
        return mean, var

    def action_values(self, obs, noise, mu_out=None, task_value_out=None, actions_out=None, neglogp_out=None,
                      operands_ready=False, value_out=None):
        """get_action_values (rl_games A2CBase; called at amp_continuous_value.py:53): mu, sigma(logstd), value (still in
        normalised units), task value, sampled action, neglogp.  obs [M,1422], noise [M,69] (standard normal).
        The *_out tensors let the heads write directly into the caller's experience rows."""
        n = self.net
        mu_out = self.mu if mu_out is None else mu_out
        task_value_out = self.task_value if task_value_out is None else task_value_out
        actions_out = self.actions if actions_out is None else actions_out
        neglogp_out = self.neglogp if neglogp_out is None else neglogp_out
        value_out = self.value if value_out is None else value_out
        assert value_out is self.value or self.chain
        mean, var = self.obs_norm.f32()
        w, b = self._w_ac1()
        h = n.actor_mlp[0].out_features
        W = self.w16.get if self.tc else None

        def trunk_ac1():
            self._trunk(obs, operands_ready)
            if self.tc:
                linear_bf16x3(self.s_ain, W("ac1", w), b, True, y16=self.s_ac1)
            else:
                linear(self.ain, w, b, relu=True, out=self.ac1)

        def actor():
            if self.tc:
                linear_bf16x3(self.s_ac1.cols(0, h), W("a2", n.actor_mlp[2].weight), n.actor_mlp[2].bias.detach(), True, y16=self.s_a2)
                if self.mu_splits > 1:
                    linear_bf16x3(self.s_a2, W("mu", n.mu.weight), n.mu.bias.detach(), False, y32=self.mu_parts[0], splits=self.mu_splits)
                    sample_actions_parts(self.mu_parts, mu_out, n.sigma, noise, actions_out, neglogp_out)
                    return
                linear_bf16x3(self.s_a2, W("mu", n.mu.weight), n.mu.bias.detach(), False, y32=mu_out)
            else:
                self._lin(self.ac1[:, :h], n.actor_mlp[2], True, self.a2)
                self._lin(self.a2, n.mu, False, mu_out)
            sample_actions(mu_out, n.sigma, noise, actions_out, neglogp_out)

        def critic():
            if self.tc:
                linear_bf16x3(self.s_ac1.cols(h, 2 * h), W("c2", n.critic_mlp[2].weight), n.critic_mlp[2].bias.detach(), True,
                              head=(n.value, self.value, self.hp_c))            # value layer fused into the epilogue
            else:
                self._lin(self.ac1[:, h:], n.critic_mlp[2], True, self.c2)
                self._lin(self.c2, n.value, False, self.value)

        def task_value():   # eval_task_value (amp_network_sept_value_builder.py:31-46): the 30 normalised trajectory features
            self._lin(obs[:, SELF_OBS:SELF_OBS + TRAJ_OBS], n._task_value_mlp[0], True, self.v1,
                      mean[SELF_OBS:SELF_OBS + TRAJ_OBS], var[SELF_OBS:SELF_OBS + TRAJ_OBS])
            self._lin(self.v1, n._task_value_mlp[2], True, self.v2)
            self._lin(self.v2, n._value_logits, False, task_value_out)

        def chain_main():
            # task MLP -> stacked actor / critic first layer -> actor / critic second layers -> mu, value: one launch
            if not operands_ready:
                eps = self.obs_norm.epsilon
                split_bf16(obs[:, :SELF_OBS], self.s_ain.cols(0, SELF_OBS), mean[:SELF_OBS], var[:SELF_OBS], eps)
                split_bf16(obs[:, SELF_OBS:], self.s_tin, mean[SELF_OBS:], var[SELF_OBS:], eps)
            tm = (self.M + 127) // 128
            L = self._policy_layers(mu_out, value_out)
            linear_chain(L, self.policy_order(tm, L), self._ws("policy", L))
            sample_actions(mu_out, n.sigma, noise, actions_out, neglogp_out)

        if self.chain:
            if self.fork is not None:
                self.fork.run(chain_main, task_value)
            else:
                chain_main(); task_value()
        elif self.fork is not None:
            trunk_ac1()
            # the actor branch (a2 -> mu -> sample) is the longer one and the physics step waits for it: it runs on a
            # high-priority side stream so that a2's tiles are dispatched before c2's and mu / sample overlap c2's second wave
            self.fork.run(critic, actor, task_value)
        else:
            trunk_ac1(); actor(); critic(); task_value()
        return dict(mus=mu_out, sigmas=n.sigma, values=value_out, task_values=task_value_out, actions=actions_out,
                    neglogpacs=neglogp_out)

    def critic(self, obs, operands_ready=False):
        """_eval_critic (common_agent.py:647-655) before value un-normalisation: [M,1]."""
        n = self.net
        self._trunk(obs, operands_ready)
        h = n.critic_mlp[0].out_features
        if self.tc:
            W = self.w16.get
            linear_bf16x3(self.s_ain, W("c0", n.critic_mlp[0].weight), n.critic_mlp[0].bias.detach(), True, y16=self.s_ac1.cols(h, 2 * h))
            linear_bf16x3(self.s_ac1.cols(h, 2 * h), W("c2", n.critic_mlp[2].weight), n.critic_mlp[2].bias.detach(), True,
                          head=(n.value, self.next_value, self.hp_c))
            return self.next_value
        self._lin(self.ain, n.critic_mlp[0], True, self.ac1[:, h:])
        self._lin(self.ac1[:, h:], n.critic_mlp[2], True, self.c2)
        self._lin(self.c2, n.value, False, self.next_value)
        return self.next_value

    def _ws(self, name, layers):
        """Zero-initialised counter workspace of a chain (the kernel leaves it zeroed); one per pass."""
        ws = self._chain_ws.get(name)
        need = chain_workspace_ints(layers)
        if ws is None or ws.numel() < need:
            ws = self._chain_ws[name] = torch.zeros(need, device=self.mu.device, dtype=torch.int32)
        return ws

    @staticmethod
    def policy_order(tm, L):
        """Ticket order of the policy pass: task MLP, stacked first layer, actor second layer, most of the critic's second
        layer, the mu tiles (their rows are complete by then), the rest of the critic."""
        t = [tiles_of(l) for l in L]
        c_first = min(t[4], 5 * tm)
        return [(0, 0, t[0]), (1, 0, t[1]), (2, 0, t[2]), (3, 0, t[3]), (4, 0, c_first), (5, 0, t[5])] + \
               ([(4, c_first, t[4] - c_first)] if t[4] > c_first else [])

    @staticmethod
    def post_order(tm, L, sms=148):
        """Ticket order of the critic(next obs) + discriminator pass: the discriminator's first layer (independent of
        everything, the longest tiles) fills the SMs while the dependent critic chain works through its short layers; the
        pass ends on the short tiles of the discriminator's second layer."""
        t = [tiles_of(l) for l in L]
        a = max(0, min(t[4], sms - t[0]))
        b = min(t[4] - a, 84 * tm // 32)
        c = t[4] - a - b
        o = [(0, 0, t[0])]
        if a: o.append((4, 0, a))
        o.append((1, 0, t[1]))
        if b: o.append((4, a, b))
        o.append((2, 0, t[2]))
        if c: o.append((4, a + b, c))
        o += [(3, 0, t[3]), (5, 0, t[5])]
        return o

    def _policy_layers(self, mu_out, value_out, base=0):
        n, W = self.net, self.w16.get
        w, b = self._w_ac1()
        h = n.actor_mlp[0].out_features
        return [chain_layer(self.s_tin, W("t0", n._task_mlp[0].weight), n._task_mlp[0].bias.detach(), True, y16=self.s_t1),
                chain_layer(self.s_t1, W("t2", n._task_mlp[2].weight), n._task_mlp[2].bias.detach(), True, dep=base,
                            y16=self.s_ain.cols(SELF_OBS, self.s_ain.K)),
                chain_layer(self.s_ain, W("ac1", w), b, True, dep=base + 1, y16=self.s_ac1),
                chain_layer(self.s_ac1.cols(0, h), W("a2", n.actor_mlp[2].weight), n.actor_mlp[2].bias.detach(), True, dep=base + 2, y16=self.s_a2),
                chain_layer(self.s_ac1.cols(h, 2 * h), W("c2", n.critic_mlp[2].weight), n.critic_mlp[2].bias.detach(), True, dep=base + 2,
                            head=(n.value, value_out, self.hp_c)),
                chain_layer(self.s_a2, W("mu", n.mu.weight), n.mu.bias.detach(), False, dep=base + 3, y32=mu_out)]

    def _next_obs_layers(self, slot, logit_out, base=0, second=False):
        """critic(next obs) + discriminator; second: operands / activations of the second set (next_obs_set)."""
        n, W, M = self.net, self.w16.get, self.M
        h = n.critic_mlp[0].out_features
        if second:
            q = self.next_obs_set()
            tin, t1, ain, c1, hp = q["tin"], q["t1"], q["ain"], q["c1"], q["hp"]
        else:
            tin, t1, ain, c1, hp = self.s_tin, self.s_t1, self.s_ain, self.s_ac1.cols(h, 2 * h), self.hp_c
        s_amp = self.s_amp.rows_view(M, slot * M)
        return [chain_layer(tin, W("t0", n._task_mlp[0].weight), n._task_mlp[0].bias.detach(), True, y16=t1),
                chain_layer(t1, W("t2", n._task_mlp[2].weight), n._task_mlp[2].bias.detach(), True, dep=base, y16=ain.cols(SELF_OBS, ain.K)),
                chain_layer(ain, W("c0", n.critic_mlp[0].weight), n.critic_mlp[0].bias.detach(), True, dep=base + 1, y16=c1),
                chain_layer(c1, W("c2", n.critic_mlp[2].weight), n.critic_mlp[2].bias.detach(), True, dep=base + 2,
                            head=(n.value, self.next_value, hp)),
                chain_layer(s_amp, W("d0", n._disc_mlp[0].weight), n._disc_mlp[0].bias.detach(), True, y16=self.s_d1),
                chain_layer(self.s_d1, W("d2", n._disc_mlp[2].weight), n._disc_mlp[2].bias.detach(), True, dep=base + 4,
                            head=(n._disc_logits, logit_out, self.hp_d))]

    @staticmethod
    def merged_order(tm, L):
        """Ticket order of the merged launch, layers 0-5 = policy pass (t0 t2 ac1 a2 c2 mu), 6-11 = next-observation pass (t0 t2 c0 c2
        d0 d2).  Both task MLPs, then the discriminator's long independent tiles (they fill the SMs while the dependent chains work through
        their first layers), the wide first layers, the second layers, and the short tiles (discriminator second layer, mu) last so that
        the launch ends evenly.  Measured on one box: 289 us against 292 (discriminator tiles split around the task MLPs), 291
        (discriminator first) and 306 (policy pass, then next-observation pass)."""
        t = [tiles_of(l) for l in L]
        return [(i, 0, t[i]) for i in (0, 6, 1, 7, 10, 2, 3, 8, 4, 9, 11, 5)]

    def merged_pass(self, noise, prev_slot, mu_out, value_out, actions_out, neglogp_out, task_value_out, obs, logit_out=None):
        """get_action_values of step n AND `_eval_critic(next obs)` + `_eval_disc` of step n-1 in ONE launch (12 layers): the two
        passes are independent given the two operand sets (next_obs_set), so the dependency bubbles of one are filled by the
        other's tiles and there is one tail instead of two.  Outputs as action_values / critic_disc."""
        n = self.net
        assert self.chain
        out = self.logit if logit_out is None else logit_out
        mean, var = self.obs_norm.f32()

        def chain_main():
            L = self._policy_layers(mu_out, value_out) + self._next_obs_layers(prev_slot, out, base=6, second=True)
            linear_chain(L, self.merged_order((self.M + 127) // 128, L), self._ws("merged", L))
            sample_actions(mu_out, n.sigma, noise, actions_out, neglogp_out)

        def task_value():
            self._lin(obs[:, SELF_OBS:SELF_OBS + TRAJ_OBS], n._task_value_mlp[0], True, self.v1,
                      mean[SELF_OBS:SELF_OBS + TRAJ_OBS], var[SELF_OBS:SELF_OBS + TRAJ_OBS])
            self._lin(self.v1, n._task_value_mlp[2], True, self.v2)
            self._lin(self.v2, n._value_logits, False, task_value_out)

        if self.fork is not None:
            self.fork.run(chain_main, task_value)
        else:
            chain_main(); task_value()
        return dict(mus=mu_out, sigmas=n.sigma, values=value_out, task_values=task_value_out, actions=actions_out,
                    neglogpacs=neglogp_out), self.next_value, out

    def critic_disc(self, obs, amp_obs, slot=0, operands_ready=False, logit_out=None, second=False):
        """`_eval_critic(next obs)` (common_agent.py:647-655, before un-normalisation) and `_eval_disc`
        (amp_continuous.py:666-668) as ONE launch: -> (next_value [M,1], disc logit [M,1])."""
        n, M = self.net, self.M
        assert self.chain and amp_obs.shape[0] == M
        out = self.logit if logit_out is None else logit_out
        if not operands_ready:
            assert not second
            s_amp = self.s_amp.rows_view(M, slot * M)
            mean, var = self.obs_norm.f32(); am, av = self.amp_norm.f32()
            split_bf16(obs[:, :SELF_OBS], self.s_ain.cols(0, SELF_OBS), mean[:SELF_OBS], var[:SELF_OBS], self.obs_norm.epsilon)
            split_bf16(obs[:, SELF_OBS:], self.s_tin, mean[SELF_OBS:], var[SELF_OBS:], self.obs_norm.epsilon)
            split_bf16(amp_obs, s_amp, am, av, self.amp_norm.epsilon)
        L = self._next_obs_layers(slot, out, second=second)
        linear_chain(L, self.post_order((M + 127) // 128, L), self._ws("post", L))
        return self.next_value, out

    def critic_timeouts(self, reset, terminate):
        """The critic on the terminal observation of the envs that were reset WITHOUT terminating (episode time-out) - the
        only envs whose `_eval_critic(next obs)` is not the next step's own critic output (emloco_timeout_gather).  Uses the
        operands the post-step sinks left in s_ain / s_tin.  -> (values [M,1] (rows 0..count-1 valid), idx int32 [M], count int32 [1])."""
        n, W, M = self.net, self.w16.get, self.M
        if self._cc is None:
            dev = reset.device
            S = lambda k: _Split(M, k, dev)
            h = n.critic_mlp[0].out_features
            self._cc = dict(ain=S(SELF_OBS + n._task_mlp[2].out_features), tin=S(TASK_OBS), t1=S(n._task_mlp[0].out_features), c1=S(h),
                            hp=torch.zeros(M, (n.critic_mlp[2].out_features + 63) // 64, device=dev), val=torch.zeros(M, 1, device=dev),
                            idx=torch.zeros(M, dtype=torch.int32, device=dev), count=torch.zeros(1, dtype=torch.int32, device=dev))
        c = self._cc
        _lib.check(_lib.load().emloco_timeout_gather(
            _ptr(reset), _ptr(terminate), M, _ptr(self.s_ain.hi), _ptr(self.s_ain.lo), self.s_ain.ld, _ptr(self.s_tin.hi),
            _ptr(self.s_tin.lo), self.s_tin.ld, _ptr(c["ain"].hi), _ptr(c["ain"].lo), c["ain"].ld, _ptr(c["tin"].hi), _ptr(c["tin"].lo),
            c["tin"].ld, _ptr(c["idx"]), _ptr(c["count"]), _stream()), "emloco_timeout_gather")
        r = c["count"]
        linear_bf16x3(c["tin"], W("t0", n._task_mlp[0].weight), n._task_mlp[0].bias.detach(), True, y16=c["t1"], rows=r)
        linear_bf16x3(c["t1"], W("t2", n._task_mlp[2].weight), n._task_mlp[2].bias.detach(), True,
                      y16=c["ain"].cols(SELF_OBS, c["ain"].K), rows=r)
        linear_bf16x3(c["ain"], W("c0", n.critic_mlp[0].weight), n.critic_mlp[0].bias.detach(), True, y16=c["c1"], rows=r)
        linear_bf16x3(c["c1"], W("c2", n.critic_mlp[2].weight), n.critic_mlp[2].bias.detach(), True, rows=r,
                      head=(n.value, c["val"], c["hp"]))
        return c["val"], c["idx"], c["count"]

    def disc_logits_all(self, out):
        """The discriminator over the operands of ALL slots at once (rows = M * amp_slots): out [M*amp_slots, 1]."""
        n, W, rows = self.net, self.w16.get, self.M * self.amp_slots
        if self._all is None:
            self._all = (_Split(rows, n._disc_mlp[0].out_features, out.device),
                         torch.empty(rows, (n._disc_mlp[2].out_features + 63) // 64, device=out.device))      # d1, head partials
        d1, d2 = self._all
        linear_bf16x3(self.s_amp, W("d0", n._disc_mlp[0].weight), n._disc_mlp[0].bias.detach(), True, y16=d1)
        linear_bf16x3(d1, W("d2", n._disc_mlp[2].weight), n._disc_mlp[2].bias.detach(), True, head=(n._disc_logits, out, d2))
        return out

    def disc_logits(self, amp_obs, out=None, operands_ready=False, slot=0):
        """_eval_disc (amp_continuous.py:666-668): normalise -> 3090->1024->512->1.  amp_obs [M',3090], M' <= M.
        slot: which block of the stored discriminator operands this call's rows occupy."""
        n = self.net
        m = amp_obs.shape[0]
        mean, var = self.amp_norm.f32()
        out = self.logit[:m] if out is None else out
        if self.tc:
            W = self.w16.get
            s_amp = self.s_amp.rows_view(m, slot * self.M)
            if not operands_ready:
                split_bf16(amp_obs, s_amp, mean, var, self.amp_norm.epsilon)
            linear_bf16x3(s_amp, W("d0", n._disc_mlp[0].weight), n._disc_mlp[0].bias.detach(), True,
                          y16=self.s_d1.rows_view(m))
            linear_bf16x3(self.s_d1.rows_view(m), W("d2", n._disc_mlp[2].weight), n._disc_mlp[2].bias.detach(), True,
                          head=(n._disc_logits, out, self.hp_d))
            return out
        linear(amp_obs, n._disc_mlp[0].weight.detach(), n._disc_mlp[0].bias.detach(), relu=True, mean=mean, var=var,
               eps=self.amp_norm.epsilon, out=self.d1[:m])
        self._lin(self.d1[:m], n._disc_mlp[2], True, self.d2[:m])
        self._lin(self.d2[:m], n._disc_logits, False, out)
        return out


def _pad64(k):
    return (k + 63) // 64 * 64


class _Split:
    """An fp32 [rows, K] matrix carried as two bf16 terms (x = hi + lo), row pitch padded to 64 elements for the TMA boxes."""

    def __init__(self, rows, K, dev):
        self.rows, self.K, self.ld = rows, K, _pad64(K)
        self.hi = torch.zeros(rows, self.ld, device=dev, dtype=torch.bfloat16)
        self.lo = torch.zeros(rows, self.ld, device=dev, dtype=torch.bfloat16)

    def cols(self, a, b):
        v = _Split.__new__(_Split)
        v.rows, v.K, v.ld, v.hi, v.lo = self.rows, b - a, self.ld, self.hi[:, a:b], self.lo[:, a:b]
        return v

    def rows_view(self, m, start=0):
        v = _Split.__new__(_Split)
        v.rows, v.K, v.ld, v.hi, v.lo = m, self.K, self.ld, self.hi[start:start + m], self.lo[start:start + m]
        return v


def split_bf16(x, dst: _Split, mean=None, var=None, eps=1e-5):
    M, K = x.shape
    assert K == dst.K and M == dst.rows and x.stride(1) == 1
    _lib.check(_lib.load().emloco_split_bf16(_ptr(x), x.stride(0), M, K, _ptr(mean), _ptr(var), eps, _ptr(dst.hi), _ptr(dst.lo),
                                             dst.ld, _stream()), "emloco_split_bf16")
    return dst


def linear_bf16x3(a: _Split, w: _Split, bias, relu, y32=None, y16: _Split = None, tile=0, rows=None, head=None, splits=0):
    """y = act(a w^T + bias) on the tcgen05 path; fp32 output and/or split output for the next layer.
    tile: 0 = library picks the output-tile width, 128 / 256 = forced.
    rows: optional int32 device scalar - only that many leading rows are valid (compacted row sets).
    splits: split-K count (> 1): y32 must be the first of `splits` consecutive [M, N] matrices, which receive partial sums.
    head: optional (layer nn.Linear(N, 1), out [M,1], part [M, ceil(N/64)]) - the single-output layer that follows, fused
          into the epilogue (emloco_linear_bf16x3_head); y32 / y16 may then be omitted."""
    M, K, N = a.rows, a.K, w.rows
    assert w.K == K
    _lib.mac_count += M * N * K
    if y32 is not None:
        assert y32.shape == (M, N) and y32.stride(1) == 1
    args = (_ptr(a.hi), _ptr(a.lo), a.ld, _ptr(w.hi), _ptr(w.lo), w.ld, _ptr(bias), M, N, K, int(bool(relu)) | (int(tile) << 8) | (int(splits) << 20),
            _ptr(y32), 0 if y32 is None else y32.stride(0), None if y16 is None else _ptr(y16.hi),
            None if y16 is None else _ptr(y16.lo), 0 if y16 is None else y16.ld)
    if head is not None:
        layer, out, part = head
        hw = layer.weight.detach()
        assert hw.shape == (1, N) and hw.is_contiguous() and out.shape == (M, 1) and out.is_contiguous()
        assert part.is_contiguous() and part.shape[0] >= M and part.shape[1] == (N + 63) // 64
        _lib.check(_lib.load().emloco_linear_bf16x3_head(_ptr(rows), *args, _ptr(hw), _ptr(layer.bias.detach()), _ptr(part), _ptr(out),
                                                         _stream()), "emloco_linear_bf16x3_head")
        return
    args = args + (_stream(),)
    if rows is None:
        _lib.check(_lib.load().emloco_linear_bf16x3(*args), "emloco_linear_bf16x3")
    else:
        _lib.check(_lib.load().emloco_linear_bf16x3_rows(_ptr(rows), *args), "emloco_linear_bf16x3_rows")


def chain_layer(a: _Split, w: _Split, bias, relu, dep=-1, y32=None, y16: _Split = None, head=None):
    """One layer of an emloco_linear_chain launch (arguments as linear_bf16x3; dep = index of the layer that writes `a`)."""
    M, K, N = a.rows, a.K, w.rows
    assert w.K == K
    L = _lib.ChainLayer()
    L.a_hi, L.a_lo, L.lda, L.w_hi, L.w_lo, L.ldw = _ptr(a.hi), _ptr(a.lo), a.ld, _ptr(w.hi), _ptr(w.lo), w.ld
    L.d_bias, L.M, L.N, L.K, L.relu, L.dep = _ptr(bias), M, N, K, int(bool(relu)), dep
    if y32 is not None:
        assert y32.shape == (M, N) and y32.stride(1) == 1
        L.d_y32, L.ldy = _ptr(y32), y32.stride(0)
    if y16 is not None:
        assert y16.rows == M and y16.K == N
        L.y_hi, L.y_lo, L.ldy16 = _ptr(y16.hi), _ptr(y16.lo), y16.ld
    if head is not None:
        layer, out, part = head
        hw = layer.weight.detach()
        assert hw.shape == (1, N) and hw.is_contiguous() and out.shape == (M, 1) and out.is_contiguous()
        assert part.is_contiguous() and part.shape[0] >= M and part.shape[1] == (N + 63) // 64
        L.d_head_w, L.d_head_bias, L.d_head_part, L.d_head_out = _ptr(hw), _ptr(layer.bias.detach()), _ptr(part), _ptr(out)
    return L


def tiles_of(layer):
    return ((layer.M + 127) // 128) * ((layer.N + 127) // 128)


def chain_workspace_ints(layers):
    arr = (_lib.ChainLayer * len(layers))(*layers)
    return int(_lib.load().emloco_linear_chain_workspace_ints(arr, len(layers)))


def linear_chain(layers, order=None, ws=None):
    """All `layers` (chain_layer) in one persistent tcgen05 launch.  order: [(layer, first tile, tile count)] or None (layer by
    layer); ws: int32 workspace, zero before its first use (torch.zeros), at least chain_workspace_ints(layers) long."""
    arr = (_lib.ChainLayer * len(layers))(*layers)
    if ws is None:
        ws = torch.zeros(chain_workspace_ints(layers), device=torch.device("cuda", torch.cuda.current_device()), dtype=torch.int32)
    o, ns = None, 0
    if order is not None:
        flat = [int(v) for seg in order for v in seg]
        o, ns = (C.c_int32 * len(flat))(*flat), len(order)
    for L in layers:
        _lib.mac_count += L.M * L.N * L.K
    _lib.check(_lib.load().emloco_linear_chain(arr, len(layers), o, ns, _ptr(ws), ws.numel(), _stream()), "emloco_linear_chain")


class _TcWeights:
    """bf16 hi/lo copies of the nn.Linear weights.  When a parameter changes (optimizer step / load_state_dict) the split is
    redone INTO THE SAME BUFFERS, so pointers captured in a CUDA graph stay valid; `generation` counts re-allocations (shape or
    device change) - the only event that invalidates a graph."""

    def __init__(self):
        self._c = {}
        self.generation = 0

    def adopt(self, name, split):
        """Use a split owned (and kept current) by someone else - the update step refreshes its operand splits after every
        optimiser step (update.PPOUpdate.adopt_into); no version checks, no copies."""
        if name not in self._c or self._c[name][1] is not split:
            self.generation += 1
        self._c[name] = ["adopted", split]

    def get(self, name, weight):
        ent = self._c.get(name)
        if ent is not None and ent[0] == "adopted":
            return ent[1]
        key = (weight.data_ptr(), weight._version)
        if ent is not None and ent[0] == key:
            return ent[1]
        w = weight.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        if ent is not None and (ent[1].rows, ent[1].K) == tuple(w.shape) and ent[1].hi.device == w.device:
            split_bf16(w, ent[1])
            ent[0] = key
            return ent[1]
        sp = _Split(w.shape[0], w.shape[1], w.device)
        split_bf16(w, sp)
        self._c[name] = [key, sp]
        self.generation += 1
        return sp


# ---- thin wrappers over the stateless C entry points -------------------------------------------------------
def normalize(x, mean, var, eps=1e-5, out=None):
    """RunningMeanStd.forward, eval branch (utils/running_mean_std.py:82-84): clamp((x-mean)/sqrt(var+eps), +-5)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32 and x.is_cuda
    M, K = x.shape
    if out is None:
        out = torch.empty(M, K, device=x.device, dtype=torch.float32)
    assert out.shape == (M, K) and out.stride(1) == 1
    _lib.check(_lib.load().emloco_normalize(_ptr(x), x.stride(0), _ptr(out), out.stride(0), M, K, _ptr(mean), _ptr(var), eps,
                                            _stream()), "emloco_normalize")
    return out


def sample_actions_parts(mu_parts, mu_out, logstd, noise, actions, neglogp):
    """sample_actions on the split-K partial sums of the mu layer: mu_parts [S, M, A] contiguous, their sum goes to mu_out."""
    S, M, A = mu_parts.shape
    assert mu_parts.is_contiguous() and mu_out.shape == (M, A) and mu_out.stride(1) == 1 and noise.is_contiguous()
    _lib.check(_lib.load().emloco_sample_actions_parts(_ptr(mu_parts), A, S, M * A, _ptr(mu_out), mu_out.stride(0), _ptr(logstd),
                                                       _ptr(noise), _ptr(actions), _ptr(neglogp), M, A, _stream()),
               "emloco_sample_actions_parts")
    return actions, neglogp


def sample_actions(mu, logstd, noise, actions=None, neglogp=None):
    M, A = mu.shape
    assert mu.stride(1) == 1 and noise.is_contiguous() and noise.shape == (M, A)
    if actions is None:
        actions = torch.empty(M, A, device=mu.device, dtype=torch.float32)
    if neglogp is None:
        neglogp = torch.empty(M, device=mu.device, dtype=torch.float32)
    _lib.check(_lib.load().emloco_sample_actions(_ptr(mu), mu.stride(0), _ptr(logstd), _ptr(noise), _ptr(actions),
                                                 _ptr(neglogp), M, A, _stream()), "emloco_sample_actions")
    return actions, neglogp


def disc_reward(logit, task_rew=None, scale=2.0, w_task=0.5, w_disc=0.5):
    """(_calc_disc_rewards, _combine_rewards): returns (disc_r, combined or None), same shape as logit."""
    lg = logit.contiguous()
    disc = torch.empty_like(lg)
    comb = None
    tr = None
    if task_rew is not None:
        tr = task_rew.contiguous()
        assert tr.numel() == lg.numel()
        comb = torch.empty_like(lg)
    _lib.check(_lib.load().emloco_disc_reward(_ptr(lg), _ptr(tr), _ptr(disc), _ptr(comb), lg.numel(), scale, w_task, w_disc,
                                              _stream()), "emloco_disc_reward")
    return disc, comb


LOG_2PI_HALF = 0.5 * math.log(2.0 * math.pi)
