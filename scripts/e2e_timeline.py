"""GPU: where a group-step of the host-buffer pipeline spends its time.  Two env groups as in HostRolloutPipeline, but each step issued
as H2D copies | graph A (reset .. post-step) | D2H obs (copy stream) | graph B (critic / discriminator / bookkeeping) | D2H rest, with
timing events between the pieces and host timestamps around submit / wait.  Prints mean durations (us) per group."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from emloco_b200.host_pipeline import HostRolloutPipeline
from emloco_b200.synthetic import synthetic_traj_pool

N, G, K = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 2, 96
pipe = HostRolloutPipeline(N, groups=G, seed=101, tensor_cores=True, traj_flags=bench.TRAJ_FLAGS, traj_pool=synthetic_traj_pool(bench.TRAJ_POOL, 0))
pipe.warm()
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
rec = {g: [] for g in range(G)}
host = {g: [] for g in range(G)}
base = E(); base.record(); torch.cuda.synchronize(); t_base = time.perf_counter()

def submit(g):
    R, h, st, cp = pipe.R[g], pipe.host[g], pipe.stream[g], pipe.copy[g]
    n = pipe.next[g] % pipe.T; pipe.next[g] = n + 1
    ev = [E() for _ in range(6)]
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        if n == 0: R.sync_weights()
        ev[0].record()
        R.sim.obs.copy_(h["obs"], non_blocking=True); R.noise.copy_(h["noise"], non_blocking=True)
        ev[1].record()
        def read_back():
            ev[2].record()
            with torch.cuda.stream(cp):
                cp.wait_event(ev[2])
                h["obs"].copy_(R.sim.obs, non_blocking=True); h["rew"].copy_(R.sim.rew, non_blocking=True); h["reset"].copy_(R.sim.reset, non_blocking=True)
                ev[3].record()
        R.step_graphed_host_noise(n, after_env_step=read_back)
        if n == pipe.T - 1: R.finish_graphed()
        ev[4].record()
        h["actions"].copy_(R.mb["actions"][n], non_blocking=True); h["neglogp"].copy_(R.mb["neglogpacs"][n], non_blocking=True); h["values"].copy_(R.mb["values"][n], non_blocking=True)
        st.wait_stream(cp)
        ev[5].record()
        pipe.done[g].record()
    rec[g].append(ev); host[g].append([t0, time.perf_counter()])

def wait(g):
    pipe.done[g].synchronize(); host[g][-1].append(time.perf_counter())

for g in range(G): submit(g)
for i in range(1, K):
    for g in range(G):
        wait(g); submit(g)
for g in range(G): wait(g)
torch.cuda.synchronize()
tot = time.perf_counter() - t_base
print(f"groups {G}: {tot / K * 1e3:.3f} ms per step of {N} envs = {N * K / tot / 1e6:.2f} M env-steps/s")
for g in range(G):
    ev = rec[g][8:]; hs = np.array(host[g][8:])
    d = lambda a, b: np.mean([e[a].elapsed_time(e[b]) for e in ev]) * 1e3
    start = np.array([base.elapsed_time(e[0]) for e in ev]); end = np.array([base.elapsed_time(e[5]) for e in ev])
    print(f" group {g}: H2D {d(0,1):.0f} | graph A {d(1,2):.0f} | D2H obs {d(2,3):.0f} | graph B(+finish) {d(2,4):.0f} | tail copies + join {d(4,5):.0f} | device total {d(0,5):.0f} us")
    print(f"          host: submit call {np.mean(hs[:,1]-hs[:,0])*1e6:.0f} us | submit->results {np.mean(hs[:,2]-hs[:,0])*1e6:.0f} us | cycle (submit to submit) {np.mean(np.diff(hs[:,0]))*1e6:.0f} us | device idle between steps of the group {np.mean(start[1:]-end[:-1])*1e3:.0f} us")
pipe.close()
