"""LocoVal 1M-batch timing, three ways: (a) one call per graph replay, (b) 20 calls captured in ONE graph, (c) 20 direct calls."""
import torch, sys
sys.path.insert(0, ".")
from emloco_b200.value_pose_net import ValuePoseNet
from emloco_b200.synthetic import synthetic_locoval_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
traj, pose, vel = (torch.from_numpy(a).cuda() for a in synthetic_locoval_batch(B, seed=0))
net = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = 20
with torch.no_grad():
    for _ in range(3): net(traj, pose, vel)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = net(traj, pose, vel)
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(R): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms_a = e0.elapsed_time(e1) / R
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for _ in range(R): out = net(traj, pose, vel)
    g2.replay(); torch.cuda.synchronize()
    e0.record(); g2.replay(); e1.record(); torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / R
    e0.record()
    for _ in range(R): net(traj, pose, vel)
    e1.record(); torch.cuda.synchronize()
    ms_c = e0.elapsed_time(e1) / R
for tag, ms in (("graph/call", ms_a), ("one graph of 20", ms_b), ("direct", ms_c)):
    print(f"locoval {B} [{tag}]: {ms:.3f} ms  {B / ms / 1e6:.2f} G scores/s  {B * 404 / ms / 1e6:.0f} GB/s")

# sustained loop with clock sampling (is the steady-state number a clock effect?)
import subprocess, time
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "100"],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
e0.record()
for _ in range(150): g2.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (150 * R)
time.sleep(0.2); p.terminate()
print(f"sustained 3000 kernels: {ms:.3f} ms each")
print(p.stdout.read())
# isolated launches: idle gap before each
tot = 0
for _ in range(10):
    torch.cuda.synchronize(); time.sleep(0.05)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
print(f"isolated (50 ms idle before each): {tot / 10:.3f} ms")
