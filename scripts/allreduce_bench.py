"""Gradient average of the training step over NCCL: dist.FlatGrads.average() on the 11.2 M-parameter network,
timed with CUDA events, max over ranks.  torchrun --nproc-per-node N scripts/allreduce_bench.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emloco_b200 import dist as D
from emloco_b200.policy import AMPSeptValueNetwork
rank, local, world = D.init("nccl")
torch.manual_seed(0)
net = AMPSeptValueNetwork().cuda()
fg = D.FlatGrads(net.parameters())
fg.flat.normal_()
for _ in range(5):
    fg.average()
torch.cuda.synchronize(); D.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for _ in range(reps):
    fg.average()
e1.record(); torch.cuda.synchronize()
ms = D.max_over_ranks(e0.elapsed_time(e1) / reps, device="cuda")
if rank == 0:
    nb = fg.nbytes()
    print(json.dumps({"collective": "all_reduce(avg) of the flat fp32 gradient", "bytes": nb, "n_gpus": world, "ms": ms,
                      "algbw_GBps": nb / ms / 1e6, "busbw_GBps": nb / ms / 1e6 * 2 * (world - 1) / world}))
D.finalize()
