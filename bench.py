#!/usr/bin/env python
"""Benchmark of the EmLoco rollout hot path (BASELINE.json: env-steps/s at 4096 humanoid envs per B200).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs 4096] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one control step of `AMPValueAgent.play_steps` (pacer/pacer/learning/amp_continuous_value.py:44-118) for all
envs of every rank: device-side reset of done envs, actor/critic/task-value forward + action sampling, physics (4 sub-steps)
+ fused post-step, critic on the next obs, discriminator -> AMP reward, bookkeeping, LocoVal scoring; and once every 32
steps the post-horizon discriminator pass over the stored [32,N,3090] AMP observations + reward combine + GAE (:150-163).
Workload = BASELINE.json configs[1] (4096 envs on one B200; the other configs are parity-test sizes).  Envs are rank-local
(weak scaling, no data-path collective).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic bytes / flops per unit (restated in DESIGN.md "Measurement")
BYTES_PHYSICS = 3296          # per env-step: actions + reduced state in, reduced + rigid-body + contact + dof force out
BYTES_POST = 28264            # per env-step: state/contact/force/verts in, obs + flip obs + AMP ring shift + rewards out
# MACs per env actually executed per step: policy 7.493 M (task MLP evaluated once, not twice as the reference does),
# next-obs critic 4.047 M, discriminator 3.689 M (the post-horizon discriminator pass is outside the step segments)
FLOP_NETS_STEP = 2 * (7.493154e6 + 4.046848e6 + 3.688960e6)
# what the post-step launch moves with the rows_only sinks on (the benched configuration), per env-step: reads = rigid-body / dof /
# contact / dof-force state 2 364 + trajectory samples 360 + the 14 older AMP ring steps 11 536; writes = observation 5 688 +
# its experience row 5 688 + mirrored observation row 5 688 + AMP row 12 360 + bf16 hi/lo operands of the first layers
# (1422 + 3090) x 4 = 18 048
BYTES_POST_WITH_SINKS = 2364 + 360 + 11536 + 3 * 5688 + 12360 + 18048
BYTES_LOCOVAL = 404           # per score
HORIZON = 32
# DRAM traffic per launch from the committed `ncu --set full` captures (cold caches under ncu, so an upper bound of what a
# warm step moves): dram__bytes_read.sum + dram__bytes_write.sum.  physics / post_step: profiles/r01g_full.md (post_step with
# the rows_only sinks: 63 MB read + 137 MB written back by the end of the launch); nets: the same capture, summed over the
# 12 tcgen05 launches of one step; locoval: profiles/r01f_locoval.md, the 1 M-score launch
NCU_TRAFFIC = {"physics": 5.36e6, "post_step": 200.5e6, "nets": 363.7e6, "locoval": 427.1e6}
# with the layer chain: the two persistent launches of a step (profiles/r02r_full.md: policy pass 98.7 MB read + 82.6 MB written,
# critic + discriminator pass 154.4 + 57.4 MB)
NCU_TRAFFIC_NETS_CHAIN = 393.1e6
# merged schedule (profiles/r02s_full.md): the ONE 12-layer launch of a step moves 306.0 MB + 164.4 MB; the post-step launch also
# writes the second operand set (next-observation copy for the critic pass): 63.7 MB read + 160.4 MB written
NCU_TRAFFIC_NETS_MERGED, NCU_TRAFFIC_POST_MERGED = 470.4e6, 224.1e6
BYTES_POST_SECOND_SET = 1422 * 4      # bf16 hi + lo of the observation, second copy


_OUT_FD = None


def capture_stdout():
    """stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version banner, torch notices)
    is sent to stderr by pointing fd 1 at fd 2 for the duration of the run; emit() writes the line to the real stdout."""
    global _OUT_FD
    if _OUT_FD is None:
        sys.stdout.flush()
        _OUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _OUT_FD is None:
        os.write(1, line)
    else:
        os.write(_OUT_FD, line)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_env():
    from emloco_b200.dist import env_rank
    return env_rank()


# =====================================================================================================
# CPU arm: the oracle port on the host cores (the Isaac Gym CPU pipeline itself cannot run here)
# =====================================================================================================
# README training command: --real_path JTA+JRDB --adjust_root_vel --init_heading (emloco TRAJ_* bits 1|2|4); the pool
# stands in for the JTA/JRDB pickles (not redistributable)
TRAJ_FLAGS, TRAJ_POOL = 1 | 2 | 4, 2048


def cpu_rollout_rate(envs, steps, warmup, seed=0, budget_s=None):
    """env-steps/s of oracle/cpu_rollout.py on `envs` envs; stops early when `budget_s` is exceeded."""
    import numpy as np
    import torch
    from emloco_b200.model import build_model_arrays, rest_root_height
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.synthetic import synthetic_env_state
    from oracle.cpu_rollout import CpuRollout, weights_from_state_dict
    A = build_model_arrays()
    torch.manual_seed(seed)
    P, D = weights_from_state_dict(AMPSeptValueNetwork().state_dict())
    from emloco_b200.synthetic import synthetic_traj_pool
    R = CpuRollout(A, synthetic_env_state(envs, seed, rest_root_height(A)), P, D, traj_flags=TRAJ_FLAGS,
                   traj_pool=synthetic_traj_pool(TRAJ_POOL, seed), traj_seed=seed)
    rng = np.random.default_rng(seed)
    # the rest of what the GPU arm's step does: LocoVal scoring of every env each step, and once per horizon the
    # discriminator over the stored AMP observations + reward combine + GAE (amp_continuous_value.py:150-163)
    from emloco_b200.synthetic import synthetic_locoval_batch
    from oracle import oracle_np as O
    lv_traj, lv_pose, lv_vel = synthetic_locoval_batch(envs, seed)
    r2 = np.random.default_rng(seed + 1)
    LW = {"fc1": (r2.normal(0, 0.1, (49, 100)).astype(np.float32), np.zeros(49, np.float32)),
          "fc2": (r2.normal(0, 0.1, (24, 49)).astype(np.float32), np.zeros(24, np.float32)),
          "fc3": (r2.normal(0, 0.1, (1, 24)).astype(np.float32), np.zeros(1, np.float32))}
    hist = {k: [] for k in ("amp_obs", "rewards", "values", "next_values", "dones")}

    def full_step():
        o = R.step(rng.standard_normal((envs, 69)).astype(np.float32))
        O.locoval_forward(lv_traj, lv_pose.copy(), lv_vel, LW)
        for k in hist:
            hist[k].append(o[k])
        if len(hist["dones"]) == HORIZON:
            amp = np.concatenate(hist["amp_obs"], 0)
            amp_r, _ = O.disc_reward(amp, R.D, R.disc_scale)
            comb = (0.5 * np.stack(hist["rewards"]) + 0.5 * amp_r.reshape(HORIZON, envs))[..., None].astype(np.float32)
            O.discount_values(np.stack(hist["dones"]).astype(np.float32), np.stack(hist["values"])[..., None].astype(np.float32), comb,
                              np.stack(hist["next_values"])[..., None].astype(np.float32))
            for k in hist:
                hist[k].clear()
    for _ in range(warmup):
        full_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        full_step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return envs * done / dt, done, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """The CPU arm: the whole workload (all `--envs` envs, same step definition as the GPU arm) on the host cores of rank 0.
    main() has already lifted torchrun's OMP_NUM_THREADS=1 (set before numpy / torch / the OpenMP physics oracle load)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    cores = host_cores()
    from oracle import oracle_np as O
    O.set_linear_backend("torch", cores)            # dense layers through torch's CPU sgemm, as the reference's nn.Linear would
    envs = args.envs
    warm = min(args.warmup, 3)
    rate, done, dt = cpu_rollout_rate(envs, args.steps, warm, budget_s=240.0)
    sample = (f"all {envs} envs per step, {done} steps in {dt:.1f} s on {cores} threads (oracle port: fp64 C physics with OpenMP, torch CPU "
              f"sgemm nets, numpy post-step / LocoVal scoring / post-horizon pass)")
    emit({
        "impl": "reference", "metric": "env_steps_per_sec", "value": rate, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": done, "warmup": warm, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.envs} SMPL-humanoid envs per GPU, PACER AMP rollout step + LocoVal scoring (configs[1])",
                   "horizon": HORIZON},
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


# =====================================================================================================
# PPO / AMP update step (BASELINE configs[2]: training step with the NCCL gradient all-reduce)
# =====================================================================================================
def train_step_bench(args, R, D, rank, world, pk):
    """One "train step" = one minibatch of `AMPValueAgent.calc_gradients` (amp_continuous_value.py:276-428): forward in training
    mode, losses, backward, gradient all-reduce over the ranks (inside the timed window), clip-norm, Adam, operand refresh -
    emloco_b200.update.PPOUpdate.  Minibatches are contiguous slices of the experience the rollout just produced (zero copy);
    the demo buffer is synthetic (the AMASS clips are not redistributable).  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist
    from emloco_b200 import _lib
    from emloco_b200.update import PPOUpdate
    T, N = R.T, R.N
    B = Ba = args.minibatch
    up = PPOUpdate(R.net, R.obs_norm, R.amp_norm, B, Ba)
    up.adopt_into(R.nets)
    out = R.finish()
    fl = lambda t: t.reshape(T * N, *t.shape[2:])
    adv = fl(out["advantages"])[:, 0]
    adv = ((adv - adv.mean()) / (adv.std() + 1e-8)).contiguous()                       # _calc_advs (common_agent.py:685-696)
    ret = fl(out["returns"]); ret = ((ret - ret.mean()) / (ret.std() + 1e-5)).contiguous()
    obs, act, nlp, mus = fl(R.mb["obses"][:T]), fl(R.mb["actions"]), fl(R.mb["neglogpacs"].unsqueeze(-1))[:, 0], fl(R.mb["mus"])
    amp = fl(R.mb["amp_obs"])
    sig = torch.full_like(mus, float(torch.exp(R.net.sigma[0])))
    g = torch.Generator(device=obs.device).manual_seed(1234 + rank)
    demo = torch.randn(B, 3090, device=obs.device, generator=g)
    nmb = (T * N) // B
    batches = []
    for i in range(nmb):
        sl = slice(i * B, (i + 1) * B)
        rp = slice(((i + 1) % nmb) * B, ((i + 1) % nmb + 1) * B)                     # "replay" rows: another minibatch of the horizon
        batches.append(dict(obs=obs[sl], actions=act[sl], old_logp_actions=nlp[sl].contiguous(), advantages=adv[sl], returns=ret[sl],
                            mu=mus[sl], sigma=sig[sl], amp_obs=amp[sl][:Ba], amp_obs_replay=amp[rp][:Ba], amp_obs_demo=demo))
    K = args.train_steps

    def barrier():
        D.barrier(); torch.cuda.synchronize()
    for i in range(3):
        up.step(batches[i % nmb])
    barrier()
    m0, l0 = _lib.mac_count, _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        up.step(batches[i % nmb])
    e1.record()
    barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1), device="cuda") / K
    macs, launches = (_lib.mac_count - m0) / K, (_lib.launch_count - l0) / K
    info = up.info()
    reducer = up.reducer_name
    # the same step without the optimiser half / without the collective, and the collective alone (bus bandwidth)
    e0.record()
    for i in range(K):
        up.forward_backward(batches[i % nmb])
    e1.record()
    barrier()
    fb_ms = D.max_over_ranks(e0.elapsed_time(e1), device="cuda") / K
    ar = None
    if world > 1:
        flat = up.flat.g
        for _ in range(3):
            dist.all_reduce(flat)
        barrier()
        e0.record()
        for _ in range(20):
            dist.all_reduce(flat)
        e1.record()
        barrier()
        ar_ms = D.max_over_ranks(e0.elapsed_time(e1), device="cuda") / 20
        nbytes = flat.numel() * 4
        ar = {"bytes": nbytes, "nccl_allreduce_ms": ar_ms, "algbw_gbs": nbytes / (ar_ms * 1e-3) / 1e9, "busbw_gbs": nbytes / (ar_ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
              "exposed_ms": max(ms - fb_ms, 0.0), "note": "nccl_allreduce_ms / bus bandwidth: a stand-alone NCCL all-reduce of the flat gradient, for reference; exposed_ms = (step - forward/backward-only step) = collective + clip-norm + Adam + operand refresh with the reducer in use"}
        if reducer == "nvls":                             # the same step with the NCCL reducer (bucketed all-reduce, norm + Adam on every rank)
            up2 = PPOUpdate(R.net, R.obs_norm, R.amp_norm, B, Ba, reducer="nccl")
            for i in range(3):
                up2.step(batches[i % nmb])
            barrier()
            e0.record()
            for i in range(K):
                up2.step(batches[i % nmb])
            e1.record()
            barrier()
            ar["ms_per_minibatch_with_nccl_reducer"] = D.max_over_ranks(e0.elapsed_time(e1), device="cuda") / K
    tf = 2 * macs / (ms * 1e-3) / 1e12
    return {"metric": "train_samples_per_sec", "value": world * B / (ms * 1e-3), "unit": "samples/s", "ms_per_minibatch": ms,
            "forward_backward_ms": fb_ms, "reducer": reducer, "minibatch": B, "amp_minibatch": Ba, "minibatches_per_epoch": nmb, "steps": K,
            "gemm_macs_per_minibatch": macs, "launches_per_minibatch": launches,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"],
                         "note": "algorithmic FLOPs of the GEMMs (forward + dgrad + wgrad + gradient-penalty double backward) counted once; bf16x3 issues 3 MMAs per product"},
            "allreduce": ar, "parameters": up.flat.n, "losses": {k: info[k] for k in ("a_loss", "c_loss", "tv_loss", "b_loss", "disc_grad_penalty", "total_norm")}}


# =====================================================================================================
# GPU arm
# =====================================================================================================
def run_ours(args):
    import torch
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    from emloco_b200 import dist as D
    numa = D.bind_to_gpu_numa_node(local_rank)          # before any pinned allocation: host buffers on the GPU's NUMA node
    D.init("nccl")
    from emloco_b200 import _lib
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_locoval_batch
    from emloco_b200.value_pose_net import ValuePoseNet

    N, K, W = args.envs, args.steps, args.warmup
    from emloco_b200.synthetic import synthetic_traj_pool
    pool = synthetic_traj_pool(TRAJ_POOL, args.seed)
    R = Rollout(N, device=local_rank, seed=D.rank_seed(args.seed, rank), tensor_cores=args.tensor_cores, recompute_disc=not args.dedup_disc,
                concurrent=not args.serial, traj_flags=TRAJ_FLAGS, traj_pool=pool, chain=False if args.no_chain else None, merged=False if args.no_merge else None)
    chain_on = bool(R.chain)
    pk = peaks()

    def barrier():
        D.barrier()
        torch.cuda.synchronize()

    step_i = [0]

    graphs = not args.eager

    n_finish = [0]

    def one_step():
        n = step_i[0] % HORIZON
        (R.step_graphed if graphs else R.step)(n)
        if n == HORIZON - 1:
            (R.finish_graphed if graphs else R.finish)()
            n_finish[0] += 1
        step_i[0] += 1

    for n in range(3):                      # eager: every lazy one-time initialisation happens here
        R.step(n)
    R.finish()
    step_i[0] = 3
    for _ in range(max(W, HORIZON + 8 if graphs else 0)):   # with graphs: every slot's graph is captured during warm-up (merged
        # schedule: the slot the graphed steps start on is captured twice, without and with an outstanding previous step)
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The timed window always holds WHOLE horizons: K is rounded up to a multiple of 32 steps and the window starts at slot 0, so
    # it contains exactly one post-horizon pass (discriminator over the stored AMP observations + combine + GAE) per 32 steps -
    # nothing is charged or credited from a stand-alone measurement.  "steps" reports the request, "steps_timed" what was timed.
    K_req, K = K, -(-K // HORIZON) * HORIZON
    while step_i[0] % HORIZON:
        one_step()
    barrier()
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    n_finish[0] = 0
    e0.record()
    for _ in range(K):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    fin_in_window = n_finish[0]
    assert fin_in_window == K // HORIZON
    launches = _lib.launch_count - l0
    # the post-horizon pass on its own (reported, not charged)
    ef0, ef1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ef0.record()
    (R.finish_graphed if graphs else R.finish)()
    ef1.record()
    torch.cuda.synchronize()
    finish_ms = ef0.elapsed_time(ef1)
    clocks = sampler.stop() if rank == 0 else None
    # per-segment device times: the same step replayed as seven per-segment graphs (8 slots) with an event between them
    merged_on = bool(graphs and R.merged)
    if merged_on:
        # merged schedule: slots 1..8 of a horizon, each holding its policy pass AND the critic / discriminator / bookkeeping of the
        # slot before it (slot 0, untimed, only leaves its own outstanding)
        def seg_cycle():
            R.flush()
            R.step_graphed(0)
            for n in range(1, 9):
                R.step_segments_graphed(n)
        seg_cycle()
        torch.cuda.synchronize()
        R.enable_segment_timing(True)
        for _ in range(3):
            seg_cycle()
        torch.cuda.synchronize()
        seg, _ = R.segment_ms()
        R.enable_segment_timing(False)
        R.flush()
    else:
        seg_step = (lambda i: R.step_segments_graphed(i % 8)) if graphs else (lambda i: R.step(i % HORIZON))
        for i in range(8):
            seg_step(i)
        torch.cuda.synchronize()
        R.enable_segment_timing(True)
        for i in range(24):
            seg_step(i)
        torch.cuda.synchronize()
        seg, _ = R.segment_ms()
        R.enable_segment_timing(False)
    ms = D.max_over_ranks(ms, device="cuda")           # slowest rank
    value = world * N * K / (ms * 1e-3)

    # ---- end to end: the vec-env / agent boundary with HOST buffers (rl_device = cpu): every step copies the step's
    # inputs (obs for the nets, policy noise) from pinned host memory and reads the results back (next obs, rewards,
    # dones, actions, neglogp, values).  emloco_b200.host_pipeline.HostRolloutPipeline: the N envs run as `--e2e-groups`
    # independent groups whose copies and compute overlap; per group a step is issued only when the host holds the previous
    # step's results of that group (what came back is what the next step is fed) ----
    from emloco_b200.host_pipeline import HostRolloutPipeline
    pipe = HostRolloutPipeline(N, groups=args.e2e_groups, device=local_rank, seed=D.rank_seed(args.seed, rank) + 101, graphs=graphs,
                               tensor_cores=args.tensor_cores, recompute_disc=not args.dedup_disc, concurrent=not args.serial,
                               traj_flags=TRAJ_FLAGS, traj_pool=pool)
    pipe.warm()
    Ke = max(HORIZON, min(K, 2 * HORIZON))
    barrier()
    e0.record()
    pipe.run(Ke)
    for st_ in pipe.stream:
        torch.cuda.current_stream().wait_stream(st_)
    e1.record()
    barrier()
    e2e_value = world * N * Ke / (D.max_over_ranks(e0.elapsed_time(e1), device="cuda") * 1e-3)
    h2d, d2h = pipe.h2d_bytes_per_step, pipe.d2h_bytes_per_step
    pipe.close(); del pipe
    torch.cuda.empty_cache()

    # ---- variant (reported beside the headline, never as it): value reuse.  The headline evaluates the critic twice per
    # observation like the reference does; with reuse the second evaluation is taken from the next step's policy pass
    # (bit-identical experience, tests/test_gpu_rollout.py::test_value_reuse_is_bit_identical_to_the_second_critic_pass) ----
    variant = None
    if not args.no_variants and graphs and args.tensor_cores:
        R.close(); del R
        torch.cuda.empty_cache()
        R2 = Rollout(N, device=local_rank, seed=D.rank_seed(args.seed, rank), tensor_cores=True, recompute_disc=not args.dedup_disc,
                     concurrent=not args.serial, reuse_values=True, traj_flags=TRAJ_FLAGS, traj_pool=pool)
        for n in range(3):
            R2.step(n)
        R2.finish()
        c, nf = [3], [0]

        def step2():
            n = c[0] % HORIZON
            R2.step_graphed(n)
            if n == HORIZON - 1:
                R2.finish_graphed()
                nf[0] += 1
            c[0] += 1
        for _ in range(HORIZON + 3):
            step2()
        while c[0] % HORIZON:
            step2()
        barrier()
        nf[0] = 0
        e0.record()
        for _ in range(K):
            step2()
        e1.record()
        barrier()
        assert nf[0] == K // HORIZON
        ms2 = D.max_over_ranks(e0.elapsed_time(e1), device="cuda")
        variant = {"value_reuse": {"value": world * N * K / (ms2 * 1e-3), "unit": "env-steps/s", "ms_per_step": ms2 / K,
                                   "note": "critic(next obs) reused from the next step's policy pass; compact critic pass for timed-out envs"}}
        R = R2

    train = None
    if not args.no_train:
        train = train_step_bench(args, R, D, rank, world, pk)

    out = None
    if rank == 0:
        # ---- LocoVal scores/s: 1M synthetic 12-step futures (configs[3]), device-resident ----
        B = args.locoval_batch
        traj, pose, vel = (torch.from_numpy(a).cuda() for a in synthetic_locoval_batch(B, seed=args.seed))
        net = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
        for _ in range(3):
            net(traj, pose, vel)
        torch.cuda.synchronize()
        # one scoring call replayed back to back from a CUDA graph (device time of the call, no host launch gap inside the
        # window); no explicit L2 flush: the 424 MB of inputs exceed the 126 MB L2
        reps = 20
        lv_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(lv_graph):
            lv_scores = net(traj, pose, vel)
        lv_graph.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            lv_graph.replay()
        e1.record()
        torch.cuda.synchronize()
        lv_ms = e0.elapsed_time(e1) / reps
        lv_rate = B / (lv_ms * 1e-3)

        # ---- rooflines from the live segment timings ----
        if merged_on:
            nets_ms = seg["nets"]
            seg["post_step"] = seg["post_step+record"]
        else:
            nets_ms = seg["policy"] + (seg["critic+disc+locoval"] if "critic+disc+locoval" in seg else seg["critic"] + seg["disc"])
        kern = {
            "physics": {"bound": "hbm", "ms": seg["physics"], "achieved": N * BYTES_PHYSICS / (seg["physics"] * 1e-3) / 1e9,
                        "peak": pk["hbm"], "unit": "GB/s",
                        "note": "latency-bound, not HBM-bound: per sub-step the critical path is 16 dependent articulated-body "
                                "updates (arm chain up, spine, root solve, and down again) of ~2.4 k cycles each = 83 % of the "
                                "kernel; one 12-warp CTA per SM because 4096 envs are 28 per SM"},
            "post_step": {"bound": "hbm", "ms": seg["post_step"], "achieved": N * BYTES_POST / (seg["post_step"] * 1e-3) / 1e9,
                          "peak": pk["hbm"], "unit": "GB/s", "note": "the launch also writes the experience rows and the normalised bf16 operands of the first layers (sinks, not counted in the algorithmic bytes); achieved_incl_sinks counts them",
                          "bytes_incl_sinks": BYTES_POST_WITH_SINKS, "achieved_incl_sinks": N * BYTES_POST_WITH_SINKS / (seg["post_step"] * 1e-3) / 1e9,
                          "frac_incl_sinks": N * BYTES_POST_WITH_SINKS / (seg["post_step"] * 1e-3) / 1e9 / pk["hbm"]},
            "nets": {"bound": "tensor", "ms": nets_ms, "achieved": N * FLOP_NETS_STEP / (nets_ms * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s"},
            "locoval": {"bound": "hbm", "ms": lv_ms, "achieved": B * BYTES_LOCOVAL / (lv_ms * 1e-3) / 1e9, "peak": pk["hbm"],
                        "unit": "GB/s"},
        }
        for name, k in kern.items():
            k["frac"] = k["achieved"] / k["peak"]
            k["traffic"] = NCU_TRAFFIC[name] if (name != "locoval" or B == 1 << 20) else None
        if chain_on:
            kern["nets"]["traffic"] = NCU_TRAFFIC_NETS_MERGED if merged_on else NCU_TRAFFIC_NETS_CHAIN
        if merged_on:
            ps = kern["post_step"]
            ps["traffic"] = NCU_TRAFFIC_POST_MERGED
            ps["bytes_incl_sinks"] = BYTES_POST_WITH_SINKS + BYTES_POST_SECOND_SET
            ps["achieved_incl_sinks"] = N * ps["bytes_incl_sinks"] / (seg["post_step"] * 1e-3) / 1e9
            ps["frac_incl_sinks"] = ps["achieved_incl_sinks"] / pk["hbm"]
            ps["note"] += "; merged schedule: the segment also holds the LocoVal scoring launch and the bookkeeping kernel of the previous step on side branches"
        dom = max(("physics", "post_step", "nets"), key=lambda k: kern[k]["ms"])
        names = {"nets": "tc::linear_chain_kernel (ONE persistent tcgen05 launch per step: policy pass of the step + critic / discriminator pass of the step before, 12 layers)" if merged_on
                 else "tc::linear_chain_kernel (the 2 persistent tcgen05 launches of a step: policy pass, critic + discriminator pass)" if chain_on
                 else "tc::linear_bf16x3_kernel (the 12 tcgen05 dense-layer launches of a step)", "physics": "physics_soa_kernel",
                 "post_step": "post_step_kernel"}
        roof = dict(kern[dom]); roof.update(kernel=names[dom], traffic=kern[dom]["traffic"], peak_source=pk["src"],
                                            traffic_source=("profiles/r02s_full.md (ncu --set full; per launch)" if merged_on else "profiles/r02r_full.md (ncu --set full; per step for nets: both chain launches; per launch otherwise)") if chain_on
                                            else "profiles/r01g_full.md + gpurun r01g_prof.ncu-rep (ncu --set full; per step for nets, per launch otherwise)")
        if dom == "nets":
            roof["note"] = ("fp32 operands are carried as bf16 hi+lo and every k-step issues 3 MMAs (bf16x3, fp32-grade products): "
                            "frac counts algorithmic FLOPs once, so its ceiling is 1/3; MMA-issue rate = 3 x achieved")

        # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ----
        cores = host_cores()
        cpu = None
        if not args.no_cpu_baseline and world == 1:     # reported at N = 1 only
            from oracle import oracle_np as O
            O.set_linear_backend("torch", cores)        # dense layers through torch's CPU sgemm, like the --impl reference arm
            rate, done, dt = cpu_rollout_rate(1024, 1000, 2, seed=args.seed, budget_s=15.0)
            cpu = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
                   "sample": f"1024 of {N} envs per step, {done} steps in {dt:.1f} s (fp64 C physics oracle with OpenMP, torch CPU sgemm nets, numpy post-step / LocoVal scoring / post-horizon pass); the --impl reference arm runs all {N} envs"}

        out = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K_req, "steps_timed": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if not args.tensor_cores else "f32 (dense layers: bf16x3 split products on tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{N} SMPL-humanoid envs per GPU, PACER AMP rollout step + LocoVal scoring (configs[1])",
                       "horizon": HORIZON, "l2": "per-step working set (obs 23 MB + AMP obs 2x51 MB + experience rows + 45 MB weights) exceeds the 126 MB L2; experience rows rotate over 32 slots",
                       "post_horizon_disc_pass": "recomputed" if not args.dedup_disc else "reused per-step logits",
                       "env_reset": "on device: state reset + TrajGenerator.reset (--real_path pool of %d synthetic polylines, "
                                    "--adjust_root_vel, --init_heading) for the envs that finish, every step" % TRAJ_POOL,
                       "tensor_cores": bool(args.tensor_cores), "cuda_graphs": graphs, "parallel_branches": not args.serial,
                       "layer_chain": chain_on, "merged_passes": merged_on},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "groups": args.e2e_groups, "numa_node": numa,
                    "note": "HostRolloutPipeline: envs split into independent groups so host<->device copies of one group overlap the other's step; per group, step k+1 is issued after the host holds step k's results"},
            "roofline": roof, "kernels": kern, "segments_ms": seg,
            "post_horizon_ms": finish_ms,
            "post_horizon_share": {"passes_in_window": fin_in_window, "expected": K / HORIZON, "charged_ms": 0.0,
                                   "note": "the window holds whole 32-step horizons (steps rounded up to steps_timed)"},
            "locoval": {"metric": "locoval_scores_per_sec", "value": lv_rate, "unit": "scores/s", "batch": B, "ms": lv_ms},
            "cpu_baseline": cpu, "variants": variant, "train_step": train,
        }
        emit(out)
    R.close()
    D.finalize()


def main():
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tensor-cores", dest="tensor_cores", action="store_true", default=True)
    ap.add_argument("--fma", dest="tensor_cores", action="store_false", help="fp32 FMA dense layers instead of the tcgen05 bf16x3 path")
    ap.add_argument("--dedup-disc", action="store_true", default=False)
    ap.add_argument("--locoval-batch", type=int, default=1 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the value-reuse variant measurement")
    ap.add_argument("--serial", action="store_true", help="no parallel graph branches (critic / discriminator / LocoVal / heads)")
    ap.add_argument("--minibatch", type=int, default=16384, help="PPO / AMP minibatch (rows) of the train-step measurement")
    ap.add_argument("--train-steps", type=int, default=16)
    ap.add_argument("--no-train", action="store_true", help="skip the train-step (update) measurement")
    ap.add_argument("--e2e-groups", type=int, default=2, help="env groups of the end-to-end (host buffer) pipeline")
    ap.add_argument("--no-merge", action="store_true", help="two chain launches per step (policy pass; critic + discriminator pass) instead of one merged launch")
    ap.add_argument("--no-chain", action="store_true", help="one launch per dense layer instead of one persistent launch per network pass")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        # all host threads for the CPU arm: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the OpenMP
        # physics oracle, OpenBLAS and torch's intra-op pool single-threaded (numpy / torch are not imported yet at this point)
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[k] = str(host_cores())
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
