"""Quick device timings of the individual kernels (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from emloco_b200.sim import EmlocoSim, gae
from emloco_b200.model import build_model_arrays, rest_root_height
from emloco_b200.value_pose_net import ValuePoseNet

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = build_model_arrays()
sim = EmlocoSim(N, model_arrays=A)
sim.root_state[:, 6] = 1; sim.root_state[:, 2] = rest_root_height(A) + 0.01; sim.root_state[:, 0:2] = 54.0
sim.reset_indexed(None)
act = (torch.rand(N, 69, device="cuda") - 0.5) * 0.2

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

t = timeit(lambda: sim.step(act)); print(f"env step (physics+post) {t*1e3:.1f} us -> {N/t*1e3:.3e} env-steps/s")
t = timeit(lambda: sim.simulate()); print(f"simulate (2 substeps)   {t*1e3:.1f} us")
t = timeit(lambda: sim.post_step(False)); print(f"post_step               {t*1e3:.1f} us  ({N*37e3/t/1e6:.0f} GB/s at ~37KB/env)")
net = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
B = 1 << 20
traj = torch.randn(B, 13, 2, device="cuda"); pose = torch.randn(B, 24, 3, device="cuda"); vel = torch.randn(B, 2, device="cuda")
t = timeit(lambda: net(traj, pose, vel)); print(f"locoval fwd 1M          {t*1e3:.1f} us -> {B/t*1e3:.3e} scores/s")
tr = traj.clone().requires_grad_(True)
def fb():
    v, l = net.calc_embodied_motion_loss(tr, pose, vel); l.backward()
t = timeit(fb, n=5, warm=2); print(f"locoval fwd+bwd 1M      {t*1e3:.1f} us")
d = torch.zeros(32, N, device="cuda"); v = torch.randn(32, N, device="cuda")
t = timeit(lambda: gae(d, v, v, v)); print(f"gae 32x{N}             {t*1e3:.1f} us")
print("root z mean", sim.root_state[:, 2].mean().item(), "finite", torch.isfinite(sim.rb_state).all().item())
