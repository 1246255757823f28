"""ncu target: a few update steps at the bench's minibatch (run under `ncu --metrics gpu__time_duration.sum`); also prints its
own CUDA-event timing when run plainly.  python scripts/profile_update.py [B] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_update import _setup, _dev
from oracle.make_golden import synth_update_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
up, net, sd, _, stats = _setup(B, B, 5, 7)
rng = np.random.default_rng(0)
f = lambda *s: torch.from_numpy(rng.normal(0, 1, s).astype(np.float32)).cuda()
batch = dict(obs=f(B, 1422), actions=f(B, 69), old_logp_actions=f(B), advantages=f(B), returns=f(B, 1), mu=f(B, 69), sigma=torch.full((B, 69), 0.055).cuda(),
             amp_obs=f(B, 3090), amp_obs_replay=f(B, 3090), amp_obs_demo=f(B, 3090))
up.step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    up.step(batch)
e1.record()
torch.cuda.synchronize()
print("ms per step", e0.elapsed_time(e1) / steps)
