"""Feasibility probe: torch symmetric memory (peer pointers over NVLink) + a kernel of libemloco_b200 reading a peer's buffer.
torchrun --nproc-per-node 2 scripts/symm_probe.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from emloco_b200 import dist as D, _lib
rank, local, world = D.init("nccl")
torch.cuda.set_device(local)
n = 1 << 20
t = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
t.fill_(float(rank + 1))
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok: world", h.world_size, "ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else None, flush=True)
h.barrier(channel=0)
y = torch.zeros(n, device="cuda")
peer = (rank + 1) % world
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
_lib.check(_lib.load().emloco_axpy(C.c_void_p(y.data_ptr()), C.c_void_p(h.buffer_ptrs[peer]), 1.0, n, st), "emloco_axpy")
torch.cuda.synchronize()
assert float(y[0]) == float(peer + 1) and float(y.sum()) == float(peer + 1) * n, (float(y[0]), float(y.sum()))
# bandwidth of peer reads through our kernel
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = symm.empty(16 << 20, dtype=torch.float32, device=torch.device("cuda", local)); big.fill_(1.0)
hb = symm.rendezvous(big, dist.group.WORLD); hb.barrier(channel=0)
yy = torch.zeros(16 << 20, device="cuda")
for _ in range(3):
    _lib.check(_lib.load().emloco_axpy(C.c_void_p(yy.data_ptr()), C.c_void_p(hb.buffer_ptrs[peer]), 1.0, 16 << 20, st), "emloco_axpy")
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    _lib.check(_lib.load().emloco_axpy(C.c_void_p(yy.data_ptr()), C.c_void_p(hb.buffer_ptrs[peer]), 1.0, 16 << 20, st), "emloco_axpy")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(rank, "peer read of 64 MB through axpy kernel: %.3f ms = %.0f GB/s" % (ms, 64 * 1.048576 / ms), flush=True)
hb.barrier(channel=0)
dist.barrier()
print(rank, "PROBE_OK", flush=True)
D.finalize()
