"""Measured impact of the two PhysX features the physics kernel does not model: joint range limits
(smpl_humanoid.xml hinge `range` attributes) and self-collision (humanoid.py:917-944 filter masks).  Runs the benched rollout
(4096 envs, tensor-core policy with Xavier weights) and reports, over all env-steps: how often an exp-map joint coordinate lies
outside its MJCF range (by how much), and how often two bodies that PhysX would collide (non-adjacent, different limbs) have
their collision primitives interpenetrating (sphere-swept approximations from the MJCF geoms).
    python scripts/physics_limits_impact.py [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from emloco_b200.mjcf import default_model
from emloco_b200.rollout import Rollout
from emloco_b200.synthetic import synthetic_traj_pool

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = 4096
mdl = default_model()
lo, hi = torch.tensor(mdl.limit_lo, dtype=torch.float32).cuda(), torch.tensor(mdl.limit_hi, dtype=torch.float32).cuda()
R = Rollout(N, seed=0, tensor_cores=True, traj_flags=bench.TRAJ_FLAGS, traj_pool=synthetic_traj_pool(bench.TRAJ_POOL, 0))
# bounding spheres of the collision primitives (centre in the body frame, radius)
gt, ga, gb, gr = mdl.geom_type, np.asarray(mdl.geom_a), np.asarray(mdl.geom_b), np.asarray(mdl.geom_r)
cen = np.where((gt == 1)[:, None], 0.5 * (ga + gb), ga)
rad = np.where(gt == 0, gr, np.where(gt == 1, gr + 0.5 * np.linalg.norm(ga - gb, axis=1), np.linalg.norm(gb, axis=1) * 0.5))
cen_t, rad_t = torch.tensor(cen, dtype=torch.float32).cuda(), torch.tensor(rad, dtype=torch.float32).cuda()
parent = np.asarray(mdl.parent)
pairs = [(i, j) for i in range(24) for j in range(i + 1, 24) if parent[j] != i and parent[i] != j and not (parent[i] == parent[j])]
pi, pj = torch.tensor([p[0] for p in pairs]).cuda(), torch.tensor([p[1] for p in pairs]).cuda()
out_frac, out_max, out_mean, pen_frac, pen_env = [], [], [], [], []

def qrot(q, v):
    qv = q[..., :3]
    t = 2 * torch.cross(qv, v, dim=-1)
    return v + q[..., 3:4] * t + torch.cross(qv, t, dim=-1)

for k in range(steps):
    R.step(k % R.T)
    if k % R.T == R.T - 1:
        R.finish()
    q = R.sim.dof_state.view(N, 69, 2)[..., 0]
    over = torch.clamp(q - hi, min=0) + torch.clamp(lo - q, min=0)
    out_frac.append((over > 0).float().mean().item()); out_max.append(over.max().item()); out_mean.append(over[over > 0].mean().item() if (over > 0).any() else 0.0)
    rb = R.sim.rb_state.view(N, 24, 13)
    c = rb[..., 0:3] + qrot(rb[..., 3:7], cen_t.expand(N, 24, 3))
    d = (c[:, pi] - c[:, pj]).norm(dim=-1) - (rad_t[pi] + rad_t[pj])
    pen = d < -0.02                                          # bounding spheres over-estimate: count clear overlaps only
    pen_frac.append(pen.float().mean().item()); pen_env.append(pen.any(dim=1).float().mean().item())
torch.cuda.synchronize()
print(json.dumps({"steps": steps, "envs": N, "joint_coords_outside_range_frac": float(np.mean(out_frac)), "mean_excess_rad_when_outside": float(np.mean(out_mean)),
                  "max_excess_rad": float(np.max(out_max)), "body_pairs_checked": len(pairs), "pair_overlap_frac": float(np.mean(pen_frac)),
                  "envs_with_any_overlap_frac": float(np.mean(pen_env))}))
R.close()
