import torch, sys, ctypes as C
sys.path.insert(0, ".")
from emloco_b200 import _lib
from emloco_b200.value_pose_net import ValuePoseNet
from emloco_b200.synthetic import synthetic_locoval_batch
B = 1 << 20
net = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(tag, traj, pose, vel):
    with torch.no_grad():
        for _ in range(3): net(traj, pose, vel)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): net(traj, pose, vel)
        e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (direct, back to back)")
syn = [torch.from_numpy(a).cuda() for a in synthetic_locoval_batch(B, seed=0)]
rnd = [torch.rand_like(t) - 0.5 for t in syn]
timeit("synthetic", *syn)
timeit("random", *rnd)
timeit("synthetic traj, random pose/vel", syn[0], rnd[1], rnd[2])
timeit("random traj, synthetic pose/vel", rnd[0], syn[1], syn[2])
z = torch.zeros_like(syn[0]); z[:, 1, 0] = 1.0
timeit("zero traj (heading along x), synthetic pose", z, syn[1], syn[2])
print("traj abs max", syn[0].abs().max().item(), "pose abs max", syn[1].abs().max().item(), "min |w1|", syn[0][:, 1].norm(dim=-1).min().item())
