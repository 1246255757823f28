import torch, sys
sys.path.insert(0, ".")
from emloco_b200.value_pose_net import ValuePoseNet
from emloco_b200.synthetic import synthetic_locoval_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
traj, pose, vel = (torch.from_numpy(a).cuda() for a in synthetic_locoval_batch(B, seed=0))
net = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): net(traj, pose, vel)
tot = 0
for _ in range(10):
    e0.record(); net(traj, pose, vel); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
ms = tot / 10
print(f"locoval {B}: {ms:.3f} ms  {B / ms / 1e6:.2f} G scores/s  {B * 404 / ms / 1e6:.0f} GB/s")
