"""GPU (needs >= 2 devices; skipped on the single-GPU test box, run with `gpurun --gpus 2`): the data-parallel update step.
Two ranks, different minibatches, one NCCL all-reduce (two buckets, the first overlapped with the backward pass) - the
parameters after the step must equal a single-process step on the AVERAGE of the two ranks' gradients, and be identical on
both ranks (Horovod `optimizer.synchronize()` semantics, amp_continuous_value.py:386-394)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("reducer", ["nccl", "nvls"])       # bucketed NCCL all-reduce / fused NVSwitch-multicast sharded step
def test_two_rank_update_equals_single_process_step_on_averaged_gradients(tmp_path, reducer):
    script = tmp_path / "worker.py"
    script.write_text(WORKER_ONE_STEP)
    env = dict(os.environ, EMLOCO_ROOT=ROOT, EMLOCO_REDUCER=reducer)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29700 + os.getpid() % 200), str(script)], env=env, capture_output=True, text=True, timeout=150)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DDP_OK" in out.stdout, out.stdout[-2000:]


WORKER_ONE_STEP = r'''
import os, sys
sys.path.insert(0, os.environ["EMLOCO_ROOT"]); sys.path.insert(0, os.path.join(os.environ["EMLOCO_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
from emloco_b200 import dist as D
rank, local_rank, world = D.init("nccl")
from test_gpu_update import _setup, _dev
from oracle.make_golden import synth_update_batch
B, Ba = 512, 256
up, net, sd, _, stats = _setup(B, Ba, 5, 7, world=None, reducer=os.environ["EMLOCO_REDUCER"])   # same weights on every rank (hvd broadcast)
assert up.world == world == 2 and up.reducer_name == os.environ["EMLOCO_REDUCER"]
batches = [_dev(synth_update_batch(B, Ba, 100 + r)[0]) for r in range(world)]
for it in range(2):                                               # two steps: moments (sharded with nvls) carry over
    up.step(batches[rank], dropout_u=batches[rank]["dropout_u"])
torch.cuda.synchronize()
mine = up.flat.p.clone()
both = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(both, mine)
assert torch.equal(both[0], both[1]), "ranks diverged"
norm2 = up.info()["total_norm"]
if rank == 0:
    ref, net2, _, _, _ = _setup(B, Ba, 5, 7, world=1)             # single process, both minibatches from the same weights / statistics
    # normaliser statistics are rank-local (the reference does not synchronise them either): keep one copy per emulated rank
    stats_r = [[{k: v.clone() for k, v in m.state_dict().items()} for m in (ref.obs_norm, ref.amp_norm)] for _ in range(world)]
    for it in range(2):
        gs = []
        for r in range(world):
            for m, s in zip((ref.obs_norm, ref.amp_norm), stats_r[r]):
                m.load_state_dict(s)
            ref.forward_backward(batches[r], dropout_u=batches[r]["dropout_u"])
            gs.append(ref.flat.g.clone())
            stats_r[r] = [{k: v.clone() for k, v in m.state_dict().items()} for m in (ref.obs_norm, ref.amp_norm)]
        ref.flat.g.copy_((gs[0] + gs[1]) / 2)
        ref.flat.state[0] = float(it + 1); ref.flat.state[1] = 0.0
        ref.reduce_and_apply()
    torch.cuda.synchronize()
    d = (ref.flat.p - mine).abs().max().item()
    n1 = ref.info()["total_norm"]
    assert abs(n1 - norm2) <= 1e-4 * n1, (n1, norm2)
    assert d <= 5e-7, d                                           # same sums up to the order of two additions
    print("DDP_OK max |p_2rank - p_single| =", d, "grad norm", n1, norm2)
dist.barrier()
D.finalize()
'''
