"""GPU: per-ticket trace and timing of the merged 12-layer launch (RolloutNets.merged_pass) at 4096 rows."""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from emloco_b200 import _lib
from emloco_b200.policy import AMP_OBS, OBS, AMPSeptValueNetwork, RolloutNets, RunningMeanStd, tiles_of
tag = sys.argv[1] if len(sys.argv) > 1 else "t"
M = 4096
torch.manual_seed(0)
net = AMPSeptValueNetwork().cuda()
on, an = RunningMeanStd(OBS).cuda(), RunningMeanStd(AMP_OBS).cuda()
obs, amp, noise = torch.randn(M, OBS, device="cuda"), torch.randn(M, AMP_OBS, device="cuda"), torch.randn(M, 69, device="cuda")
nets = RolloutNets(net, on, an, M, tensor_cores=True, concurrent=True, chain=True)
nets.sync_weights()
nets.action_values(obs, noise); nets.critic_disc(obs, amp)
q = nets.next_obs_set()
q["tin"].hi.copy_(nets.s_tin.hi); q["tin"].lo.copy_(nets.s_tin.lo); q["ain"].hi.copy_(nets.s_ain.hi); q["ain"].lo.copy_(nets.s_ain.lo)
f = lambda *s: torch.zeros(*s, device="cuda")
mu, val, act, nlp, tv = f(M, 69), f(M, 1), f(M, 69), f(M), f(M, 1)
run = lambda: nets.merged_pass(noise, 0, mu, val, act, nlp, tv, obs)
run(); torch.cuda.synchronize()
def timed(fn, reps=30):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))
out = {"merged_graph_us": timed(run)}
T_ = lambda L, i: tiles_of(L[i])
orders = {
    "seq_policy_first": lambda tm, L: [(i, 0, T_(L, i)) for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11)],
    "d0_first": lambda tm, L: [(10, 0, T_(L, 10)), (0, 0, T_(L, 0)), (6, 0, T_(L, 6)), (1, 0, T_(L, 1)), (7, 0, T_(L, 7)), (2, 0, T_(L, 2)), (8, 0, T_(L, 8)),
                               (3, 0, T_(L, 3)), (4, 0, T_(L, 4)), (9, 0, T_(L, 9)), (11, 0, T_(L, 11)), (5, 0, T_(L, 5))],
    "long_first": lambda tm, L: [(0, 0, T_(L, 0)), (6, 0, T_(L, 6)), (1, 0, T_(L, 1)), (7, 0, T_(L, 7)), (10, 0, T_(L, 10)), (2, 0, T_(L, 2)), (3, 0, T_(L, 3)), (8, 0, T_(L, 8)),
                                 (4, 0, T_(L, 4)), (9, 0, T_(L, 9)), (11, 0, T_(L, 11)), (5, 0, T_(L, 5))],
}
base = RolloutNets.merged_order
for name, o in orders.items():
    RolloutNets.merged_order = staticmethod(o)
    out[name + "_graph_us"] = timed(run)
RolloutNets.merged_order = staticmethod(base)
tr = torch.zeros(8192, 8, dtype=torch.int64, device="cuda")
_lib.load().emloco_linear_chain_trace(C.c_void_p(tr.data_ptr()))
run(); torch.cuda.synchronize()
_lib.load().emloco_linear_chain_trace(None)
t = tr.cpu().numpy()
np.save(f"gpurun_out/{tag}_trace_merged.npy", t[t[:, 2] > 0])
json.dump(out, open(f"gpurun_out/{tag}_merged.json", "w"), indent=1)
print(out)
