"""The rollout hot loop on one GPU: `AMPValueAgent.play_steps` of the reference
(pacer/pacer/learning/amp_continuous_value.py:34-178) over `EmlocoSim` + `RolloutNets`.

Per control step, for all N envs of this rank (every item is a kernel of libemloco_b200.so):
  env_reset(done)            emloco_reset_done            :45
  get_action_values          emloco_linear x10, emloco_normalize, emloco_sample_actions   :53
  env_step                   emloco_step (physics + fused post-step)                        :61
  _eval_critic(next obs)     emloco_linear x5, emloco_normalize                             :85
  _calc_amp_rewards          emloco_linear x3                                               :93
  rewards/values/bookkeeping emloco_rollout_record                                          :63-118
and once per horizon: discriminator over the stored [T,N,3090] AMP observations, reward combine, GAE
(emloco_linear x3 per chunk, emloco_disc_reward, emloco_gae)  :150-163.

Experience rows are written in place ([T,N,...] tensors, as rl_games' ExperienceBuffer lays them out); the caller
reads them after `play_steps` exactly as `batch_dict` of the reference.  Envs are rank-local: no collective here.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .policy import (ACTIONS, AMP_OBS, OBS, AMPSeptValueNetwork, Fork, RolloutNets, RunningMeanStd)
from .sim import EmlocoSim, _ptr, _stream, gae
from .synthetic import synthetic_env_state


class Rollout:
    def __init__(self, num_envs, device=0, horizon=32, seed=0, tensor_cores=False, gamma=0.99, tau=0.95,
                 task_reward_w=0.5, disc_reward_w=0.5, disc_reward_scale=2.0, inversion_penalty_scale=0.3,
                 step_to_pred=144, normalize_value=True, net=None, obs_norm=None, amp_norm=None, value_norm=None,
                 recompute_disc=True, valuenet=None, fuse_sinks=True, concurrent=True, reuse_values=False, traj_flags=None,
                 traj_pool=None, traj_deferred=None, finetune=False, value_lr=1e-3,
                 min_cum_rewards=-10.0, max_cum_rewards=100.0, sim_cfg=None, traj_cfg=None, rows_only=True, chain=None, merged=None):
        self.N, self.T, self.device = int(num_envs), int(horizon), int(device)
        self.gamma, self.tau = gamma, tau
        self.task_reward_w, self.disc_reward_w, self.disc_reward_scale = task_reward_w, disc_reward_w, disc_reward_scale
        self.recompute_disc = recompute_disc
        dev = torch.device("cuda", device)
        torch.cuda.set_device(dev)
        self.sim = EmlocoSim(self.N, device=device, **(sim_cfg or {}))      # sim_cfg: emloco_cfg overrides (formats.kwargs_from_reference_cfg)
        if net is None:
            torch.manual_seed(seed)
            net = AMPSeptValueNetwork()
        self.net = net.to(dev)
        self.obs_norm = (obs_norm or RunningMeanStd(OBS)).to(dev)
        self.amp_norm = (amp_norm or RunningMeanStd(AMP_OBS)).to(dev)
        self.value_norm = (value_norm or RunningMeanStd(1)).to(dev)
        # chain (default with tensor cores + parallel branches): each of the two network passes of a step is ONE persistent
        # launch over all its dense layers (emloco_linear_chain); the value-reuse variant keeps the per-layer launches
        self.chain = (bool(tensor_cores) and bool(concurrent) and not reuse_values) if chain is None else (bool(chain) and bool(tensor_cores))
        self.nets = RolloutNets(self.net, self.obs_norm, self.amp_norm, self.N, tensor_cores=tensor_cores, concurrent=concurrent,
                                amp_slots=self.T if (tensor_cores and recompute_disc) else 1, chain=self.chain)
        self.concurrent = bool(concurrent)
        # with the chain every SM holds one of its CTAs (197 KB of shared memory), so the LocoVal kernel (107 KB) cannot run beside
        # it: it becomes a branch of the post-step launch instead (its inputs are fixed at reset, nothing in the step feeds it)
        self.locoval_with_post = self.chain and self.concurrent
        self._side = Fork(dev, 1) if concurrent else None      # outer branches (noise draw, trajectory reset) around the nets' own fork
        # value reuse needs the operand sinks (the compact critic reads the rows the post-step kernel wrote)
        self.reuse_values = bool(reuse_values) and bool(tensor_cores) and bool(fuse_sinks)
        self.gen = torch.Generator(device=dev).manual_seed(seed + 1)

        # synthetic initial state + JTA-shaped trajectories; rank-local seed like run.py:65
        st = synthetic_env_state(self.N, seed=seed, root_height=self.sim.rest_height)
        self.init_root = torch.from_numpy(st["root"]).to(dev)
        self.init_dof = torch.from_numpy(st["dof"]).to(dev)
        self.sim.traj_verts.copy_(torch.from_numpy(st["verts"]).to(dev))
        f = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)
        T, N = self.T, self.N
        # experience rows; `obses` has one spare row (T) that receives the observation following the last step of a horizon
        self.mb = dict(obses=f(T + 1, N, OBS), actions=f(T, N, ACTIONS), neglogpacs=f(T, N), values=f(T, N, 1), mus=f(T, N, ACTIONS),
                       task_values=f(T, N, 1), rewards=f(T, N, 1), next_values=f(T, N, 1), dones=f(T, N), amp_obs=f(T, N, AMP_OBS),
                       amp_rewards=f(T, N, 1), flip_obs=f(T, N, OBS))      # flip_obs: motion_sym_loss rows (:74-75)
        # fuse_sinks: the post-step / reset kernels write the experience rows and the normalised bf16 operands of the first
        # layers themselves (emloco_set_post_sinks) instead of separate copy / split launches
        self.fuse = bool(fuse_sinks)
        self.rows_only = bool(rows_only)
        self.sim.reset.fill_(1)
        if self.fuse:
            self.sim.set_post_sinks(self.nets.post_sinks(obs_copy=self.mb["obses"][T]))
        self.sim.reset_done(self.init_root, self.init_dof)
        if not self.fuse:
            self.mb["obses"][T].copy_(self.sim.obs)

        # LocoVal inputs captured at reset (humanoid_pedestrain_terrain.py:509-515, vec_task_wrappers.py:47-66)
        self.waypoint_traj = f(N, 13, 3)                                                # origin-relative, z = 0
        self.waypoint_traj[:, :, 0:2] = torch.from_numpy(st["waypoints"]).to(dev)
        rb = self.sim.rb_state.view(self.N, 24, 13)
        self.init_pose = (rb[:, :, 0:3] - rb[:, :1, 0:3]).contiguous()
        self.init_vel = self.init_root[:, 7:9].contiguous()
        # traj_flags (emloco TRAJ_* bits, None = off): every later env reset regenerates the env's trajectory and these
        # three LocoVal inputs on the device (emloco_set_traj_reset: TrajGenerator.reset + _reset_task)
        self.inverted, self._traj_deferred = None, False
        if traj_flags is not None:
            self.inverted = torch.zeros(N, device=dev, dtype=torch.uint8)
            self.traj_pool = None if traj_pool is None else torch.as_tensor(traj_pool, dtype=torch.float32).to(dev).contiguous()
            # with parallel branches the stage is deferred: it overlaps the policy pass (nothing before post_step reads it)
            self._traj_deferred = bool(concurrent) if traj_deferred is None else (bool(traj_deferred) and bool(concurrent))
            self.sim.set_traj_reset(self.sim.traj_cfg(flags=traj_flags | (_lib.TRAJ_DEFERRED if self._traj_deferred else 0), seed=seed, pool=self.traj_pool, waypoint_traj=self.waypoint_traj,
                                                      init_pose=self.init_pose, init_vel=self.init_vel, inverted=self.inverted, **(traj_cfg or {})))
        if valuenet is None:
            from .value_pose_net import ValuePoseNet
            valuenet = ValuePoseNet(True, True, mutate_pose=False)
        self.valuenet = valuenet.to(dev).eval()
        # finetune: the `_do_finetune` block of play_steps (:122-146) - LocoVal learns from the discounted returns of the
        # episodes that just finished, every control step, on the device (ValuePoseNet.finetune_step)
        self.finetune, self.value_lr = bool(finetune), float(value_lr)
        self.cum_reward_range = (float(min_cum_rewards), float(max_cum_rewards))          # common_agent.py:154-155
        if self.finetune:
            self.valuenet.enable_finetune()
        # merged (CUDA-graph steps only): `_eval_critic(next obs)` + `_calc_amp_rewards` + the bookkeeping of step n-1 run INSIDE
        # step n - the two network passes become ONE 12-layer emloco_linear_chain launch after the reset (RolloutNets.merged_pass),
        # the bookkeeping kernel (reading snapshots of flags / rewards) runs beside the next post-step; `finish` completes the last step.  Same numbers, bit for bit.
        # Needs what the bench configuration has: the chain, operand sinks, the deferred trajectory reset (it is what clears the
        # reset flags, after the snapshot the late bookkeeping reads) and no LocoVal fine-tuning inside the step.
        can_merge = bool(self.chain and self.concurrent and self.fuse and self._traj_deferred and not self.finetune and not self.reuse_values)
        self.merged = can_merge if merged is None else (bool(merged) and can_merge)
        self._pending = None                  # step whose critic / discriminator / bookkeeping is still outstanding
        if self.merged:
            self._snap = torch.zeros(2, N, device=dev, dtype=torch.int64)          # reset / terminate flags of the outstanding step
            self._snap_inv = torch.zeros(N, device=dev, dtype=torch.uint8)
            self._snap_rew = torch.zeros(N, device=dev, dtype=torch.float32)
            self._side3 = Fork(dev, 2)
        self._marks = None
        self._graphs = {}
        self._cur = {}
        self._warmed = False

        self.state = f(6, N)
        self.state[3].fill_(1.0)      # discount_coefs start at 1
        self.noise = f(N, ACTIONS)
        self.rcfg = _lib.RolloutCfg(inversion_penalty_scale, 1.0, 0.0, 1.0, disc_reward_scale, gamma, step_to_pred,
                                    int(bool(normalize_value)))
        # value_mean_std as the record kernels read it: {mean, sqrt(var + eps)} in a device buffer of fixed address
        self.value_stats = f(2)
        self.rcfg.d_value_stats = self.value_stats.data_ptr()
        self._fingerprint = None
        self.sync_weights()
        self.launches_per_step = 0

    def sync_weights(self):
        """Call after the parameters or any normaliser changed (optimiser step, load_state_dict, running-statistics update):
        refreshes, in place, every derived buffer the kernels read - bf16 weight splits, stacked first layer, fp32 normaliser
        copies and the value un-normalisation constants (`value_mean_std(value, True)`, common_agent.py:653-654).  `play_steps`
        and step 0 of the graphed paths call it themselves, so a captured CUDA graph never replays on stale numbers; if a
        parameter or buffer was re-allocated (its address changed) the graphs are dropped and captured again."""
        vn = self.value_norm
        self.value_stats[0:1].copy_(vn.running_mean.reshape(-1)[:1])
        self.value_stats[1:2].copy_(torch.sqrt(vn.running_var.reshape(-1)[:1].float() + vn.epsilon))
        realloc = self.nets.sync_weights()
        fp = self.nets.pointer_fingerprint() + (self.value_stats.data_ptr(), self.valuenet._weights().data_ptr())
        if realloc or fp != self._fingerprint:
            self._graphs.clear()
            self._fingerprint = fp

    MERGED_SEGMENTS = ("reset", "nets", "physics", "post_step+record")

    @property
    def SEGMENTS(self):
        if getattr(self, "_seg_merged", False):
            return self.MERGED_SEGMENTS
        if self.concurrent:
            return ("reset", "policy", "physics", "post_step", "critic+disc+locoval", "record")
        return ("reset", "policy", "physics", "post_step", "critic", "disc", "record")

    def enable_segment_timing(self, on=True):
        """Record a CUDA event on the launching stream at every segment boundary of step(); read with segment_ms()."""
        self._marks = [] if on else None

    def _mark(self):
        if self._marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append(e)

    def segment_ms(self):
        """Mean device time per segment over the steps recorded since enable_segment_timing(); call after a sync."""
        k = len(self.SEGMENTS) + 1
        ev = self._marks
        steps = len(ev) // k
        out = {s: 0.0 for s in self.SEGMENTS}
        for i in range(steps):
            for j, s in enumerate(self.SEGMENTS):
                out[s] += ev[i * k + j].elapsed_time(ev[i * k + j + 1])
        return {s: v / max(steps, 1) for s, v in out.items()}, steps

    # ---- one control step: everything inside the `for n in range(horizon_length)` body, as seven segments ----
    def _segment_fns(self, n, noise=None, host_obs=False, merged=False):
        if merged and self.merged and not host_obs:
            return self._segment_fns_merged(n, noise)
        sim, nets, mb, cur = self.sim, self.nets, self.mb, self._cur

        fuse = self.fuse                                # post-step side: experience rows + operands of critic(next obs) / disc
        fuse_in = self.fuse and not host_obs            # policy side: host-provided observations need their operands re-derived
        nxt = n + 1                                                                    # row T is the spare row
        slot = n if getattr(nets, "amp_slots", 1) > 1 else 0                           # block of the stored discriminator operands

        def draw_noise():
            cur["noise"] = self.noise.normal_(generator=self.gen) if noise is None else noise

        def seg_reset():                                                               # env_reset(done_indices), :45-46
            if self.concurrent:
                self._side.run(reset_main, draw_noise)
            else:
                reset_main(); draw_noise()

        def reset_main():
            if fuse_in:
                if n == 0:
                    mb["obses"][0].copy_(mb["obses"][self.T])                          # observation that followed the last horizon
                sim.set_post_sinks(nets.post_sinks(obs_copy=mb["obses"][n]))           # reset rows are patched in place
                sim.reset_done(self.init_root, self.init_dof)
            else:
                sim.set_post_sinks(None)
                sim.reset_done(self.init_root, self.init_dof)
                mb["obses"][n].copy_(sim.obs)

        def seg_policy():                                                              # get_action_values, :53
            if self._traj_deferred:
                self._side.run(policy_main, sim.traj_reset)                             # _reset_task overlaps the policy pass
            else:
                policy_main()

        def policy_main():
            # the heads write straight into row n of the experience tensors (experience_buffer.update_data, :47-56)
            cur["res"] = nets.action_values(sim.obs, cur["noise"], mu_out=mb["mus"][n], task_value_out=mb["task_values"][n],
                                            actions_out=mb["actions"][n], neglogp_out=mb["neglogpacs"][n], operands_ready=fuse_in)

        def seg_physics():                                                             # env_step: pre_physics + simulate
            sim.physics_step(cur["res"]["actions"])

        def seg_post():                                                                #           post_physics_step
            if self.locoval_with_post:
                self._side.run(post_main, seg_locoval)
            else:
                post_main()

        def post_main():
            if fuse:
                # rows_only: the mirrored observation and the AMP ring are written once, into the experience rows (the ring of
                # the next step is shifted out of row n); sim.flip_obs / sim.amp_obs stay stale while the rollout runs fused
                sim.set_post_sinks(nets.post_sinks(obs_copy=mb["obses"][nxt], amp_copy=mb["amp_obs"][n], slot=slot,
                                                   flip_copy=mb["flip_obs"][n], rows_only=self.rows_only, second=self.merged))
                sim.post_step(True)
            else:
                sim.post_step(True)
                mb["amp_obs"][n].copy_(sim.amp_obs.view(self.N, AMP_OBS))
                mb["flip_obs"][n].copy_(sim.flip_obs)

        # value reuse: critic(next obs) is NOT recomputed for envs that are not reset - the next step's policy pass evaluates
        # the critic on that very observation; only timed-out envs need their terminal observation evaluated now.  The last
        # step of the horizon has no next step inside the horizon and runs the full pass.
        reuse = self.reuse_values and fuse_in and n < self.T - 1

        def seg_critic():                                                              # _eval_critic(next obs), :85
            if reuse:
                cur["cval"], cur["cidx"], cur["ccount"] = nets.critic_timeouts(sim.reset, sim.terminate)
            else:
                cur["nv"] = nets.critic(sim.obs, operands_ready=fuse)

        def seg_disc():                                                                # _calc_amp_rewards, :93
            cur["logit"] = nets.disc_logits(mb["amp_obs"][n] if fuse else sim.amp_obs.view(self.N, AMP_OBS), operands_ready=fuse, slot=slot)

        inv = None if self.inverted is None else _ptr(self.inverted)      # task.inverted -> inversion penalty (:78-83)

        def seg_record():
            if reuse:
                # deferred next values: terminated -> 0, timed-out -> compact critic, the rest is completed by step n+1;
                # this step's critic output (policy pass) completes step n-1
                prev = n > 0
                _lib.check(_lib.load().emloco_rollout_record_deferred(
                    C.byref(self.rcfg), _ptr(sim.rew), _ptr(sim.reset), _ptr(sim.terminate), _ptr(cur["res"]["values"]),
                    _ptr(cur["logit"]), inv, _ptr(mb["values"][n]), _ptr(mb["rewards"][n]), _ptr(mb["dones"][n]),
                    _ptr(mb["next_values"][n]), _ptr(mb["amp_rewards"][n]), _ptr(self.state), self.N, _ptr(cur["cval"]),
                    _ptr(cur["cidx"]), _ptr(cur["ccount"]), _ptr(mb["dones"][n - 1]) if prev else None,
                    _ptr(mb["next_values"][n - 1]) if prev else None, _stream()), "emloco_rollout_record_deferred")
                return
            if self.reuse_values and fuse_in and n > 0:
                # last step of the horizon (full critic pass below): it still has to complete step n-1
                _lib.check(_lib.load().emloco_fill_next_values(
                    C.byref(self.rcfg), _ptr(cur["res"]["values"]), _ptr(mb["dones"][n - 1]), _ptr(mb["next_values"][n - 1]), self.N,
                    _stream()), "emloco_fill_next_values")
            seg_record_plain()

        def seg_record_plain():
            # values are un-normalised (get_action_values, normalize_value) together with next_values in the record kernel
            _lib.check(_lib.load().emloco_rollout_record(
                C.byref(self.rcfg), _ptr(sim.rew), _ptr(sim.reset), _ptr(sim.terminate), _ptr(cur["res"]["values"]),
                _ptr(cur["nv"]), _ptr(cur["logit"]), inv, _ptr(mb["values"][n]), _ptr(mb["rewards"][n]), _ptr(mb["dones"][n]),
                _ptr(mb["next_values"][n]), _ptr(mb["amp_rewards"][n]), _ptr(self.state), self.N, _stream()),
                "emloco_rollout_record")

        def seg_locoval():
            self.locoval_scores = self.valuenet(self.waypoint_traj, self.init_pose, self.init_vel)

        def seg_chain2():  # critic(next obs) and the discriminator: one launch over the six layers
            cur["nv"], cur["logit"] = nets.critic_disc(sim.obs, mb["amp_obs"][n] if fuse else sim.amp_obs.view(self.N, AMP_OBS), slot=slot,
                                                       operands_ready=fuse)

        def seg_nets2():   # critic(next obs), discriminator and LocoVal scoring are independent: three graph branches
            # the critic chain (5 dependent layers) is the longer branch: it goes on a high-priority side stream and the
            # discriminator's big GEMM fills the SMs its small layers leave idle (183 -> 172 us for the segment)
            if self.chain and not reuse:
                if self.locoval_with_post:
                    seg_chain2()
                else:
                    nets.fork.run(seg_chain2, seg_locoval)
                return
            nets.fork.run(seg_disc, seg_critic, seg_locoval)

        def seg_record_ft():
            seg_record()
            if self.finetune:                                                      # :122-146, after the bookkeeping of this step
                self.valuenet.finetune_step(self.waypoint_traj, self.init_pose, self.init_vel, self.state[4], lr=self.value_lr,
                                            min_cum_rewards=self.cum_reward_range[0], max_cum_rewards=self.cum_reward_range[1])

        if self.concurrent:
            return [seg_reset, seg_policy, seg_physics, seg_post, seg_nets2, seg_record_ft]
        return [seg_reset, seg_policy, seg_physics, seg_post, seg_critic, seg_disc, lambda: (seg_record_ft(), seg_locoval())]

    def _record_args(self, k, value_raw, nv, logit, reset, terminate, inverted, rew=None):
        mb = self.mb
        return (C.byref(self.rcfg), _ptr(self.sim.rew if rew is None else rew), _ptr(reset), _ptr(terminate), _ptr(value_raw), _ptr(nv), _ptr(logit),
                None if inverted is None else _ptr(inverted), _ptr(mb["values"][k]), _ptr(mb["rewards"][k]), _ptr(mb["dones"][k]),
                _ptr(mb["next_values"][k]), _ptr(mb["amp_rewards"][k]), _ptr(self.state), self.N, _stream())

    def _value_buf(self, k):
        return self.nets.value if k % 2 == 0 else self.nets.value2

    def _segment_fns_merged(self, n, noise=None):
        """Step n of the merged schedule (see __init__): reset -> ONE launch for get_action_values(n) and, when step n-1 is still
        outstanding, its critic(next obs) + discriminator -> physics(n) -> post-step(n) beside the bookkeeping of n-1."""
        sim, nets, mb, cur, T = self.sim, self.nets, self.mb, self._cur, self.T
        pend = self._pending
        assert pend is None or pend == n - 1, "merged steps must be consecutive (flush() in between otherwise)"

        def draw_noise():
            cur["noise"] = self.noise.normal_(generator=self.gen) if noise is None else noise

        def snapshot():          # the flags / inversion marks of step n-1, before the trajectory reset of step n clears / redraws them
            torch.stack((sim.reset, sim.terminate), out=self._snap)
            self._snap_rew.copy_(sim.rew)
            if self.inverted is not None:
                self._snap_inv.copy_(self.inverted)

        def reset_main():
            if n == 0:
                mb["obses"][0].copy_(mb["obses"][T])
            sim.set_post_sinks(nets.post_sinks(obs_copy=mb["obses"][n]))           # patches the first operand set only
            sim.reset_done(self.init_root, self.init_dof)

        def seg_reset():
            if pend is None:
                self._side3.run(reset_main, draw_noise)
            else:
                self._side3.run(reset_main, draw_noise, snapshot)

        def nets_main():
            kw = dict(mu_out=mb["mus"][n], task_value_out=mb["task_values"][n], actions_out=mb["actions"][n], neglogp_out=mb["neglogpacs"][n],
                      value_out=self._value_buf(n))
            if pend is None:
                cur["res"] = nets.action_values(sim.obs, cur["noise"], operands_ready=True, **kw)
            else:
                cur["res"], cur["nv"], cur["logit"] = nets.merged_pass(cur["noise"], pend, obs=sim.obs, **kw)

        def seg_nets():
            self._side.run(nets_main, sim.traj_reset)

        def record_prev():
            _lib.check(_lib.load().emloco_rollout_record(*self._record_args(
                pend, self._value_buf(pend), cur["nv"], cur["logit"], self._snap[0], self._snap[1],
                None if self.inverted is None else self._snap_inv, rew=self._snap_rew)), "emloco_rollout_record")

        def seg_physics():
            sim.physics_step(cur["res"]["actions"])

        def locoval():
            self.locoval_scores = self.valuenet(self.waypoint_traj, self.init_pose, self.init_vel)

        def post_main():
            sim.set_post_sinks(nets.post_sinks(obs_copy=mb["obses"][n + 1], amp_copy=mb["amp_obs"][n], slot=n, flip_copy=mb["flip_obs"][n],
                                               rows_only=self.rows_only, second=True))
            sim.post_step(True)

        def seg_post():      # the bookkeeping of step n-1 (it reads snapshots) hides under the memory-bound post-step launch
            if pend is None:
                self._side3.run(post_main, locoval)
            else:
                self._side3.run(post_main, locoval, record_prev)

        return [seg_reset, seg_nets, seg_physics, seg_post]

    def _flush_kernels(self):
        """critic(next obs) + discriminator + bookkeeping of the outstanding step (nothing has been reset since its post-step, so
        the flags are read in place)."""
        k = self._pending
        if k is None:
            return
        nv, logit = self.nets.critic_disc(None, self.mb["amp_obs"][k], slot=k, operands_ready=True, second=True)
        _lib.check(_lib.load().emloco_rollout_record(*self._record_args(k, self._value_buf(k), nv, logit, self.sim.reset, self.sim.terminate,
                                                                        self.inverted)), "emloco_rollout_record")

    def flush(self):
        """Completes the step the merged schedule left outstanding (its next values, AMP rewards, dones, bookkeeping rows)."""
        self._flush_kernels()
        self._pending = None

    def step(self, n, noise=None, host_obs=False, merged=False):
        """host_obs: the caller overwrote sim.obs (host-provided observations): operands are re-derived from it.
        merged: use the merged schedule (the step's critic / discriminator / bookkeeping are left to the next step or to flush())."""
        if n == 0 and not torch.cuda.is_current_stream_capturing():
            self.sync_weights()
        self._warmed = True
        merged = bool(merged) and self.merged and not host_obs
        self._seg_merged = merged                 # SEGMENTS / segment_ms() name the schedule that ran last
        if not merged and self._pending is not None:
            self.flush()
        for f in self._segment_fns(n, noise, host_obs, merged):
            self._mark()
            f()
        self._mark()
        self._pending = n if merged else None

    # ---- after the horizon: disc over the stored AMP obs, combine, GAE (:150-163) ----
    def finish(self):
        self._finish_warm = True
        self.flush()
        mb, T, N = self.mb, self.T, self.N
        if self.recompute_disc:
            if getattr(self.nets, "amp_slots", 1) == T:
                # the normalised operands of all T steps are still in place: one [T*N]-row GEMM chain
                self.nets.disc_logits_all(self._logits_TN().view(T * N, 1))
            else:
                for t in range(T):    # [T*N,3090] in N-row chunks through the same workspace
                    self.nets.disc_logits(mb["amp_obs"][t], out=self._logits_TN()[t])
            lg = self._logits_TN()
            _lib.check(_lib.load().emloco_disc_reward(_ptr(lg), _ptr(mb["rewards"]), _ptr(mb["amp_rewards"]), _ptr(self._comb()),
                                                      T * N, self.disc_reward_scale, self.task_reward_w, self.disc_reward_w,
                                                      _stream()), "emloco_disc_reward")
        else:
            # the per-step AMP rewards are the same numbers (same rows, same weights, eval-mode normaliser): combine only
            _lib.check(_lib.load().emloco_disc_reward(None, _ptr(mb["rewards"]), _ptr(mb["amp_rewards"]), _ptr(self._comb()),
                                                      T * N, self.disc_reward_scale, self.task_reward_w, self.disc_reward_w,
                                                      _stream()), "emloco_disc_reward")
        adv, ret = gae(mb["dones"], mb["values"], self._comb(), mb["next_values"], self.gamma, self.tau)
        self.mb_advs, self.mb_returns = adv, ret
        out = {k: (v[:T] if k == "obses" else v) for k, v in mb.items()}
        out.update(returns=ret, advantages=adv, task_rewards=mb["rewards"], rewards=self._comb())   # mb_rewards := combined (:160)
        self._finish_out = out
        return out

    def _logits_TN(self):
        if not hasattr(self, "_lg"):
            self._lg = torch.empty(self.T, self.N, 1, device=self.state.device)
        return self._lg

    def _comb(self):
        if not hasattr(self, "_cb"):
            self._cb = torch.empty(self.T, self.N, 1, device=self.state.device)
        return self._cb

    # ---- CUDA graphs: one captured graph per horizon slot (the experience-row pointers differ per slot) ----
    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        g.register_generator_state(self.gen)          # the policy-noise generator advances inside the graph
        l0 = _lib.launch_count
        marks, self._marks = self._marks, None        # timing events cannot be recorded into a capture
        with torch.cuda.graph(g):
            fn()
        self._marks = marks
        launches = _lib.launch_count - l0             # kernels of ours inside this graph (counted again on every replay)
        _lib.launch_count = l0
        return g, launches

    def _replay(self, key, fn):
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(fn)
        ent[0].replay()
        _lib.launch_count += ent[1]

    def step_graphed(self, n):
        """Same work as step(n) replayed from a CUDA graph (captured on first use; run a few eager steps first so that
        every lazy one-time initialisation - constant tables, weight splits, function attributes - has happened)."""
        if n == 0:
            self.sync_weights()
        if not self._warmed:          # the very first step runs eagerly: one-time initialisations must not land in a capture
            self.step(n)
            return
        if not self.merged:
            self._replay(n, lambda: self.step(n))
            return
        # merged schedule: the graph of slot n also holds the outstanding part of step n-1 (none at the start of a horizon)
        if self._pending is not None and self._pending != n - 1:
            self.flush()
        key = n if self._pending is not None or n == 0 else (n, "first")
        self._replay(key, lambda: self.step(n, merged=True))
        self._pending = n

    def step_graphed_host_noise(self, n, after_env_step=None):
        """step(n) on observations and policy noise the caller put into sim.obs / self.noise (no generator call in the graph).
        after_env_step: called between the env step (reset .. post_step) and the critic / discriminator / bookkeeping part, so
        that a vec-env style caller can start reading the step's observations back while the rest of the step runs."""
        if n == 0:
            self.sync_weights()
        if after_env_step is None:
            self._replay(("hn", n), lambda: self.step(n, noise=self.noise, host_obs=True))
            return
        fns = lambda: self._segment_fns(n, self.noise, True)
        k = 4                                                                           # reset, policy, physics, post_step
        self._replay(("hnA", n), lambda: [f() for f in fns()[:k]])
        after_env_step()
        self._replay(("hnB", n), lambda: [f() for f in fns()[k:]])

    def step_segments_graphed(self, n):
        """step(n) as seven per-segment graphs with a timing event between them: per-segment device time without host
        launch gaps inside a segment (used by bench.py for the roofline numbers)."""
        fns = None
        if n == 0:
            self.sync_weights()
        self._seg_merged = self.merged
        if self._pending is not None and self._pending != n - 1:
            self.flush()
        pend = self._pending is not None
        for i, name in enumerate(self.SEGMENTS):
            self._mark()
            if ("seg", n, i, pend) not in self._graphs and fns is None:
                fns = self._segment_fns(n, merged=self.merged)
            self._replay(("seg", n, i, pend), fns[i] if fns else None)
        self._mark()
        self._pending = n if self.merged else None

    def finish_graphed(self):
        if not getattr(self, "_finish_warm", False):      # first call eager: workspaces are allocated outside a capture
            return self.finish()
        self._replay("finish" if (self._pending is not None or not self.merged) else ("finish", "flushed"), self.finish)
        self._pending = None
        return self._finish_out

    def play_steps(self, graphed=False):
        self.sync_weights()
        for n in range(self.T):
            (self.step_graphed if graphed else self.step)(n)
        return self.finish_graphed() if graphed else self.finish()

    def close(self):
        self.sim.close()
