"""CPU: the N>1 plumbing (world_size 2, gloo): rank seeds, env sharding, max-over-ranks timing, scalar averaging, and the
bench's reference arm under a 2-rank launch (rank 0 prints one JSON line, the others exit 0 without work)."""
import json
import os
import subprocess
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from emloco_b200 import dist as D
    r, lr, w = D.init("gloo")
    lo, hi = D.shard_envs(4097, r, w)
    out = dict(rank=r, world=w, seed=D.rank_seed(7, r), shard=(lo, hi), tmax=D.max_over_ranks(10.0 + r),
               tsum=D.sum_over_ranks(1.0 + r), kl=D.average_scalar(0.5 * (r + 1)))
    # the gradient average of the training step: one all-reduce over one flat buffer that the .grad views alias
    import torch
    torch.manual_seed(0)                                      # same weights on every rank (hvd broadcast)
    from emloco_b200.policy import AMPSeptValueNetwork
    net = AMPSeptValueNetwork()
    fg = D.FlatGrads(net.parameters())
    x = torch.full((4, 1422), 0.1 * (r + 1))                  # rank-local batch
    loss = net.mu(net.actor_mlp(torch.cat([x[:, :368], net._task_mlp(x[:, 368:])], -1))).sum() + net.mu.bias.sum() * (r + 1)
    loss.backward()
    k = [i for i, p in enumerate(fg.params) if p is net.mu.weight][0]
    assert net.mu.weight.grad.data_ptr() == fg.flat[sum(p.numel() for p in fg.params[:k]):].data_ptr()     # .grad aliases the bucket
    local = fg.flat.clone()
    fg.average()
    out.update(nparams=fg.flat.numel(), local_sum=float(local.double().sum()), avg_sum=float(fg.flat.double().sum()),
               sigma_grad=float(net.mu.bias.grad.mean()), nbytes=fg.nbytes())
    fg.zero()
    assert float(net.mu.weight.grad.abs().sum()) == 0.0       # the views survive zeroing
    # the update step's bucketed reduction (emloco_b200.update.PPOUpdate -> dist.BucketedAllReduce): first bucket asynchronous
    flat = torch.arange(1000, dtype=torch.float32) * (r + 1)
    red = D.BucketedAllReduce(flat, 333)
    red.start_first()
    flat[333:] += 1.0                                          # "backward kernels" of the second bucket issued meanwhile
    red.finish()
    want = torch.arange(1000, dtype=torch.float32) * 3
    want[333:] += 2.0
    out["bucketed_ok"] = bool(torch.equal(flat, want))
    flat2 = torch.ones(10) * (r + 1)
    D.BucketedAllReduce(flat2, 4, overlap=False).finish()
    out["single_ok"] = bool(torch.equal(flat2, torch.full((10,), 3.0)))
    D.finalize()
    q.put(out)


def test_two_rank_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda d: d["rank"])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [d["seed"] for d in res] == [7, 8]
    assert res[0]["shard"] == (0, 2049) and res[1]["shard"] == (2049, 4097)      # disjoint, covering
    assert all(d["tmax"] == 11.0 for d in res) and all(d["tsum"] == 3.0 for d in res)
    assert all(abs(d["kl"] - 0.75) < 1e-12 for d in res)
    # gradient average: both ranks hold the mean of the two local gradients
    assert res[0]["nparams"] == res[1]["nparams"] > 11_000_000 and res[0]["nbytes"] == 4 * res[0]["nparams"]
    mean = 0.5 * (res[0]["local_sum"] + res[1]["local_sum"])
    assert all(abs(d["avg_sum"] - mean) <= 1e-6 * max(1.0, abs(mean)) for d in res)
    assert all(abs(d["sigma_grad"] - 5.5) < 1e-5 for d in res)                   # 4 rows + (1 + 2) / 2
    assert all(d["bucketed_ok"] and d["single_ok"] for d in res)


def test_bench_reference_arm_under_two_ranks():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(31500 + os.getpid() % 2000), WORLD_SIZE="2", OMP_NUM_THREADS="2")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        outs.append(subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                                    "--warmup", "1", "--envs", "16"], env=e, capture_output=True, text=True, timeout=600))
    assert all(o.returncode == 0 for o in outs), [o.stderr[-500:] for o in outs]
    lines = [l for l in outs[0].stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and outs[1].stdout.strip() == ""
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "env-steps/s"
