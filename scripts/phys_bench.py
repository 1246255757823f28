"""Physics kernel timing vs. adaptive refinement and action scale (4096 envs, physics_step only, graph replay)."""
import sys, torch
sys.path.insert(0, ".")
from emloco_b200.sim import EmlocoSim
from emloco_b200.synthetic import synthetic_env_state
N = 4096
for max_turn in (0.3, 0.0):
    for scale in (0.05, 0.3, 1.0):
        sim = EmlocoSim(N, max_turn=max_turn)
        st = synthetic_env_state(N, 0, sim.rest_height)
        root, dof = torch.from_numpy(st["root"]).cuda(), torch.from_numpy(st["dof"]).cuda()
        sim.root_state.copy_(root); sim.dof_state.copy_(dof); sim.reset_indexed(None)
        g = torch.Generator(device="cuda").manual_seed(0)
        acts = [(torch.rand(N, 69, device="cuda", generator=g) * 2 - 1) * scale for _ in range(8)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for it in range(160):
            a = acts[it % 8]
            if it >= 40 and it % 20 == 0:
                e0.record(); sim.physics_step(a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            else:
                sim.physics_step(a)
            sim.post_step(True)
            sim.reset_done(root, dof.view(N * 69, 2))
        z = sim.root_state.view(N, 13)[:, 2]
        print(f"max_turn {max_turn} action scale {scale}: physics {sum(ts)/len(ts):.1f} us (min {min(ts):.1f} max {max(ts):.1f}), mean root z {z.mean().item():.2f}")
        sim.close()
