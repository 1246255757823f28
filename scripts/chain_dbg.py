import sys, torch, numpy as np
sys.path.insert(0, ".")
from emloco_b200.policy import AMP_OBS, OBS, AMPSeptValueNetwork, RolloutNets, RunningMeanStd
M = 4096
torch.manual_seed(0)
net = AMPSeptValueNetwork().cuda()
on, an = RunningMeanStd(OBS).cuda(), RunningMeanStd(AMP_OBS).cuda()
obs, amp, noise = torch.randn(M, OBS, device="cuda"), torch.randn(M, AMP_OBS, device="cuda"), torch.randn(M, 69, device="cuda")
nets = RolloutNets(net, on, an, M, tensor_cores=True, concurrent=True, chain=True)
nets.sync_weights()
nets.action_values(obs, noise); nets.critic_disc(obs, amp); torch.cuda.synchronize()
def timed(fn, reps=30):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)*1e3)
    return float(np.median(ts))
print("policy", timed(lambda: nets.action_values(obs, noise, operands_ready=True)), "post", timed(lambda: nets.critic_disc(obs, amp, operands_ready=True)))
