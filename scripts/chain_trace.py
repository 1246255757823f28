"""GPU: where the layer-chain launches spend their time.  Per pass (policy / critic + discriminator, 4096 rows) the per-ticket trace
(emloco_linear_chain_trace) and event timings of the chain with several ticket orders, of the per-layer launches, and of
single-layer chains (steady-state time per 128x128x64 k-block).  Writes gpurun_out/<tag>_chain_trace.npz / .json."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from emloco_b200 import _lib
from emloco_b200.policy import (AMP_OBS, OBS, SELF_OBS, AMPSeptValueNetwork, RolloutNets, RunningMeanStd, _Split, chain_layer,
                                linear_chain, split_bf16, tiles_of)

tag = sys.argv[1] if len(sys.argv) > 1 else "t"
M = 4096
torch.manual_seed(0)
net = AMPSeptValueNetwork().cuda()
on, an = RunningMeanStd(OBS).cuda(), RunningMeanStd(AMP_OBS).cuda()
obs, amp, noise = torch.randn(M, OBS, device="cuda"), torch.randn(M, AMP_OBS, device="cuda"), torch.randn(M, 69, device="cuda")
nets = RolloutNets(net, on, an, M, tensor_cores=True, concurrent=True, chain=True)
nets.sync_weights()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=20, cold=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


out = {}
# --- passes: chain vs per-layer (per-layer path needs a nets object without chain) ---
nets0 = RolloutNets(net, on, an, M, tensor_cores=True, concurrent=True, chain=False)
nets0.sync_weights()
nets.action_values(obs, noise); nets.critic_disc(obs, amp); nets0.action_values(obs, noise); nets0.critic(obs); nets0.disc_logits(amp)
torch.cuda.synchronize()
out["policy_chain_us"] = timed(lambda: nets.action_values(obs, noise, operands_ready=True))
out["policy_layers_us"] = timed(lambda: nets0.action_values(obs, noise, operands_ready=True))
out["post_chain_us"] = timed(lambda: nets.critic_disc(obs, amp, operands_ready=True))
out["post_layers_us"] = timed(lambda: (nets0.fork.run(lambda: nets0.disc_logits(amp, operands_ready=True), lambda: nets0.critic(obs, operands_ready=True))))

# --- graph-replayed (no host launch gaps) ---
def graphed(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay

for name, fn in (("policy_chain", lambda: nets.action_values(obs, noise, operands_ready=True)),
                 ("policy_layers", lambda: nets0.action_values(obs, noise, operands_ready=True)),
                 ("post_chain", lambda: nets.critic_disc(obs, amp, operands_ready=True)),
                 ("post_layers", lambda: nets0.fork.run(lambda: nets0.disc_logits(amp, operands_ready=True), lambda: nets0.critic(obs, operands_ready=True)))):
    out[name + "_graph_us"] = timed(graphed(fn))

# --- alternative ticket orders ---
seq = staticmethod(lambda tm, L, *a: [(i, 0, tiles_of(l)) for i, l in enumerate(L)])
pol0, post0 = RolloutNets.policy_order, RolloutNets.post_order
RolloutNets.policy_order = seq
out["policy_chain_seq_us"] = timed(lambda: nets.action_values(obs, noise, operands_ready=True))
T_ = lambda L, i: tiles_of(L[i])
RolloutNets.policy_order = staticmethod(lambda tm, L, *a: [(0, 0, T_(L, 0)), (1, 0, T_(L, 1)), (2, 0, T_(L, 2)), (3, 0, T_(L, 3)), (5, 0, T_(L, 5)), (4, 0, T_(L, 4))])
out["policy_chain_mu_before_c2_graph_us"] = timed(graphed(lambda: nets.action_values(obs, noise, operands_ready=True)))
RolloutNets.policy_order = seq
out["policy_chain_seq_graph_us"] = timed(graphed(lambda: nets.action_values(obs, noise, operands_ready=True)))
# critic tiles interleaved with the actor's second layer so that the pass ends on short (mu) tiles
RolloutNets.policy_order = staticmethod(lambda tm, L, *a: [(0, 0, T_(L, 0)), (1, 0, T_(L, 1)), (2, 0, T_(L, 2)), (4, 0, T_(L, 4)), (3, 0, T_(L, 3)), (5, 0, T_(L, 5))])
out["policy_chain_c2_first_graph_us"] = timed(graphed(lambda: nets.action_values(obs, noise, operands_ready=True)))
RolloutNets.policy_order = staticmethod(pol0)
RolloutNets.post_order = seq
out["post_chain_seq_graph_us"] = timed(graphed(lambda: nets.critic_disc(obs, amp, operands_ready=True)))
out["post_chain_seq_us"] = timed(lambda: nets.critic_disc(obs, amp, operands_ready=True))
RolloutNets.post_order = staticmethod(lambda tm, L, *a: [(4, 0, tiles_of(L[4])), (0, 0, tiles_of(L[0])), (1, 0, tiles_of(L[1])), (2, 0, tiles_of(L[2])),
                                                       (3, 0, tiles_of(L[3])), (5, 0, tiles_of(L[5]))])
out["post_chain_discfirst_us"] = timed(lambda: nets.critic_disc(obs, amp, operands_ready=True))
RolloutNets.post_order = staticmethod(post0)

# --- single-layer chains: steady-state time per k-block ---
def single(Mr, N, K, reps=20):
    a, w = _Split(Mr, K, "cuda"), _Split(N, K, "cuda")
    split_bf16(torch.randn(Mr, K, device="cuda"), a); split_bf16(torch.randn(N, K, device="cuda") / K ** 0.5, w)
    y = _Split(Mr, N, "cuda")
    b = torch.zeros(N, device="cuda")
    L = [chain_layer(a, w, b, True, y16=y)]
    ws = torch.zeros(64 + Mr // 128, dtype=torch.int32, device="cuda")
    us = timed(lambda: linear_chain(L, None, ws), reps)
    units = tiles_of(L[0]) * ((K + 63) // 64)
    from emloco_b200.policy import linear_bf16x3
    us128 = timed(lambda: linear_bf16x3(a, w, b, True, y16=y, tile=128), reps)
    us256 = timed(lambda: linear_bf16x3(a, w, b, True, y16=y, tile=256), reps) if N >= 256 else None
    us2c = timed(lambda: linear_bf16x3(a, w, b, True, y16=y, tile=256 | 0x800), reps) if N >= 256 else None
    return dict(M=Mr, N=N, K=K, chain_us=us, units_per_sm=units / 148, us_per_unit=us / (units / 148), layer128_us=us128, layer256_us=us256, pair256_us=us2c)

out["single"] = [single(4096, 1024, 2048), single(4096, 1024, 3090), single(4096, 2048, 624), single(4096, 4096, 624), single(4096, 512, 1054),
                 single(37 * 128, 1024, 2048), single(148 * 128, 1024, 2048), single(148 * 128, 2048, 2048, 5)]

# --- traces ---
for name, fn, nl in (("policy", lambda: nets.action_values(obs, noise, operands_ready=True), 6), ("post", lambda: nets.critic_disc(obs, amp, operands_ready=True), 6)):
    tr = torch.zeros(4096, 8, dtype=torch.int64, device="cuda")
    fn(); torch.cuda.synchronize()
    _lib.load().emloco_linear_chain_trace(C.c_void_p(tr.data_ptr()))
    fn(); torch.cuda.synchronize()
    _lib.load().emloco_linear_chain_trace(None)
    out_tr = tr.cpu().numpy()
    np.save(f"gpurun_out/{tag}_trace_{name}.npy", out_tr[out_tr[:, 2] > 0])

json.dump(out, open(f"gpurun_out/{tag}_chain_trace.json", "w"), indent=1)
print(json.dumps(out, indent=1))
