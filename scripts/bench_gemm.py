"""Per-layer timing of the bf16x3 tcgen05 kernels for every tile variant (run on the GPU box)."""
import sys, json
sys.path.insert(0, ".")
import torch
from emloco_b200.policy import _Split, linear_bf16x3, split_bf16
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
shapes = [("task1", 512, 1054), ("task2", 256, 512), ("ac1", 4096, 624), ("c1", 2048, 624), ("a2/c2", 1024, 2048),
          ("mu", 69, 1024), ("value", 1, 1024), ("d1", 1024, 3090), ("d2", 512, 1024), ("dlogit", 1, 512)]
tiles = [("1cta128", 128), ("1cta256", 256), ("2cta128", 0x800 + 128), ("2cta256", 0x800 + 256)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
for name, N, K in shapes:
    a, w = _Split(M, K, "cuda"), _Split(N, K, "cuda")
    split_bf16(torch.randn(M, K, device="cuda"), a); split_bf16(torch.randn(N, K, device="cuda") / K ** 0.5, w)
    bias = torch.zeros(N, device="cuda")
    y16 = _Split(M, N, "cuda") if N % 32 == 0 else None
    y32 = None if y16 is not None else torch.empty(M, N, device="cuda")
    res = {}
    for tn, t in tiles:
        for _ in range(3):
            linear_bf16x3(a, w, bias, True, y32=y32, y16=y16, tile=t)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            linear_bf16x3(a, w, bias, True, y32=y32, y16=y16, tile=t)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res[tn] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 1))
    out[name] = res
    print(f"{name:8s} N={N:5d} K={K:5d} " + "  ".join(f"{k}: {v[0]:7.1f}us {v[1]:6.1f}TF" for k, v in res.items()), flush=True)
json.dump(out, open("gpurun_out/gemm_tiles.json", "w"))
