#!/usr/bin/env python
"""The reference's OWN GPU path on this B200, per segment, beside ours ("the bar to beat on the same box is the
PyTorch eager / TorchScript path on B200 and cuBLAS nn.Linear").  Needs the staged reference files:
    python scripts/stage_reference.py          (build container; writes baseline/_ref, git-ignored)
    gpurun -- 'EMLOCO_REFERENCE=baseline/_ref python scripts/ref_gpu_bar.py > gpurun_out/ref_gpu_bar.json'
Segments at 4096 envs (device time, CUDA events, 20 repetitions after warm-up):
  post_step   the TorchScript observation / reward / reset / AMP-observation functions driven as the task code drives them
              (oracle/make_golden.post_step_run: humanoid.py:1626-1687, humanoid_amp.py:917-971, ..terrain.py:394-491,883-930,1468-1530)
  policy      RunningMeanStd + AMPSeptValueBuilder.Network eval_actor + eval_critic + eval_task_value (fp32 cuBLAS nn.Linear)
  critic      _eval_critic on the next observation;  disc: _calc_disc_rewards
  gae         discount_values python loop (T = 32);  locoval: ValuePoseNet on 1 M rows
Physics has no reference implementation to time (PhysX binaries absent).  Ours: the same segments from the live rollout."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import make_golden as G
from oracle import netweights, ref_extract

N, T = 4096, 32
dev = "cuda"


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    R = ref_extract.load()
    out = {"envs": N, "device": torch.cuda.get_device_name(0), "torch": torch.__version__, "allow_tf32": torch.backends.cuda.matmul.allow_tf32}
    st = G.synth_state(N, 0, map_shape=(1080, 1080), rough=False)
    I = G.post_step_inputs(st, dev)
    with torch.no_grad():
        out["ref_post_step_ms"] = timeit(lambda: G.post_step_run(I))
    net = ref_extract.load_network().to(dev)
    net.load_state_dict({k: torch.from_numpy(v).to(dev) for k, v in netweights.synth_state_dict(0).items()})
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rms, arms, vrms = R.RunningMeanStd((1422,)).to(dev).eval(), R.RunningMeanStd((3090,)).to(dev).eval(), R.RunningMeanStd((1,)).to(dev).eval()
    H = ref_extract.load_agent_blocks()
    h = H()
    h.running_mean_std, h._amp_input_mean_std, h.value_mean_std = rms, arms, vrms
    h.model = types.SimpleNamespace(a2c_network=net, eval=lambda: None)
    h.normalize_input = h.normalize_value = h._normalize_amp_input = True
    h._disc_reward_mean_std = None; h.ppo_device = dev; h._disc_reward_scale = 2.0
    obs = torch.randn(N, 1422, device=dev); amp = torch.randn(N, 3090, device=dev); noise = torch.randn(N, 69, device=dev)

    def policy():
        x = h._preproc_obs(obs)
        mu, logstd = net.eval_actor(x)
        value = net.eval_critic(x)
        tv = net.eval_task_value(x)
        sigma = torch.exp(logstd)
        a = mu + sigma * noise                                           # rl_games ModelA2CContinuousLogStd: sample + neglogp
        nl = 0.5 * (((a - mu) / sigma) ** 2).sum(-1) + 0.5 * np.log(2 * np.pi) * 69 + logstd.sum(-1)
        return a, nl, vrms(value, True), tv
    with torch.no_grad():
        out["ref_policy_ms"] = timeit(policy)
        out["ref_critic_ms"] = timeit(lambda: h._eval_critic({"obs": obs}))
        out["ref_disc_ms"] = timeit(lambda: h._calc_amp_rewards(amp))
        ag = R.meth.AgentHolder(); ag.horizon_length = T; ag.gamma = 0.99; ag.tau = 0.95
        d, v, r, nv = torch.zeros(T, N, device=dev), torch.randn(T, N, 1, device=dev), torch.rand(T, N, 1, device=dev), torch.randn(T, N, 1, device=dev)
        out["ref_gae_ms"] = timeit(lambda: ag.discount_values(d, v, r, nv))
        amp_all = torch.randn(T * N, 3090, device=dev)
        out["ref_post_horizon_disc_ms"] = timeit(lambda: h._calc_amp_rewards(amp_all), reps=5)
        with contextlib.redirect_stdout(io.StringIO()):
            vp = R.ValuePoseNet(use_pose=True, use_vel=True).to(dev).eval()
        from emloco_b200.synthetic import synthetic_locoval_batch
        B = 1 << 20
        traj, pose, vel = (torch.from_numpy(a).to(dev) for a in synthetic_locoval_batch(B, seed=0))
        out["ref_locoval_1m_ms"] = timeit(lambda: vp(traj, pose.clone(), vel), reps=5)
    # ---- ours, same box, same sizes: live segment timings of the rollout + LocoVal ----
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_traj_pool
    from emloco_b200.value_pose_net import ValuePoseNet
    Ro = Rollout(N, seed=0, tensor_cores=True, traj_flags=7, traj_pool=synthetic_traj_pool(2048, 0))
    for k in range(3):
        Ro.step(k)
    Ro.finish()
    for k in range(3, 3 + T):
        Ro.step_graphed(k % T)
    for i in range(8):
        Ro.step_segments_graphed(i % 8)
    torch.cuda.synchronize()
    Ro.enable_segment_timing(True)
    for i in range(24):
        Ro.step_segments_graphed(i % 8)
    torch.cuda.synchronize()
    seg, _ = Ro.segment_ms()
    out["ours_segments_ms"] = seg
    out["ours_post_horizon_ms"] = timeit(lambda: Ro.finish_graphed(), reps=5)
    lv = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
    out["ours_locoval_1m_ms"] = timeit(lambda: lv(traj, pose, vel), reps=10)
    ref_step = out["ref_post_step_ms"] + out["ref_policy_ms"] + out["ref_critic_ms"] + out["ref_disc_ms"]
    ours_step = seg["policy"] + seg["post_step"] + seg["critic+disc+locoval"] + seg["record"]
    out["summary"] = {"ref_step_without_physics_ms": ref_step, "ours_same_segments_ms": ours_step, "speedup": ref_step / ours_step,
                      "note": "reference = its TorchScript / eager / cuBLAS-fp32 code on this GPU, no PhysX; ours = policy + post_step + critic/disc/LocoVal + record segments (physics and reset excluded on both sides)"}
    print(json.dumps(out))
    Ro.close()


if __name__ == "__main__":
    main()
