"""End-to-end (host buffers) throughput of HostRolloutPipeline for several group counts.  python scripts/e2e_sweep.py [groups...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from emloco_b200 import dist as D
from emloco_b200.host_pipeline import HostRolloutPipeline
from emloco_b200.synthetic import synthetic_traj_pool
numa = D.bind_to_gpu_numa_node(0)
pool = synthetic_traj_pool(bench.TRAJ_POOL, 0)
N, K = 4096, 64
out = {"numa": numa}
for G in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
    if N % G:
        continue
    pipe = HostRolloutPipeline(N, groups=G, seed=101, tensor_cores=True, traj_flags=bench.TRAJ_FLAGS, traj_pool=pool)
    pipe.warm()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run(K)
    for st in pipe.stream:
        torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out[f"groups_{G}"] = {"ms_per_step": ms, "env_steps_per_s": N / (ms * 1e-3)}
    pipe.close(); del pipe
    torch.cuda.empty_cache()
print(json.dumps(out))
