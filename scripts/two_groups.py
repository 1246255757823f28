"""Device-resident throughput of G independent env groups stepping concurrently on one GPU (streams), vs one group of all envs."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from emloco_b200.rollout import Rollout
from emloco_b200.synthetic import synthetic_traj_pool
pool = synthetic_traj_pool(bench.TRAJ_POOL, 0)
N, T = 4096, 32
out = {}
for G in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
    Rs = [Rollout(N // G, seed=g, tensor_cores=True, traj_flags=bench.TRAJ_FLAGS, traj_pool=pool) for g in range(G)]
    sts = [torch.cuda.Stream() for _ in range(G)]
    def step(k):
        for R, st in zip(Rs, sts):
            with torch.cuda.stream(st):
                R.step_graphed(k % T)
                if k % T == T - 1:
                    R.finish_graphed()
    for R, st in zip(Rs, sts):
        with torch.cuda.stream(st):
            for k in range(3):
                R.step(k)
            R.finish()
    torch.cuda.synchronize()
    for k in range(3, 3 + 2 * T - 3):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in sts:
        st.wait_event(e0)
    K = 2 * T
    for k in range(K):
        step(k)
    for st in sts:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[f"groups_{G}"] = {"ms_per_step": ms, "env_steps_per_s": N / (ms * 1e-3)}
    for R in Rs:
        R.close()
    del Rs
    torch.cuda.empty_cache()
print(json.dumps(out))
