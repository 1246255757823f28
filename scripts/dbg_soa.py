import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_gpu_physics import _setup, _random_state, _gpu_step, _oracle_step
A, M, PO, h0 = _setup()
for N in (250, 256):
    root, dof_pos, dof_vel, actions = _random_state(N, 7, 2.0, 3.0)
    a = _gpu_step(N, root, dof_pos, dof_vel, actions, impl=0)
    b = _gpu_step(N, root, dof_pos, dof_vel, actions, impl=1)
    o = _oracle_step(A, M, PO, root, dof_pos, dof_vel, actions)
    for k in ("root", "rb", "dof", "dof_force"):
        da = np.abs(a[k] - o[k]).reshape(N, -1).max(1); db = np.abs(b[k] - o[k]).reshape(N, -1).max(1)
        bad = np.nonzero(da > 0.02)[0]
        print(N, k, "soa max err", da.max(), "warp max err", db.max(), "bad envs", bad[:20], "lanes", (bad % 32)[:20])
    bad = np.nonzero(np.abs(a["rb"] - o["rb"]).reshape(N, 24, 13).max(2) > 0.02)
    print("bad (env, body):", list(zip(bad[0][:30], bad[1][:30])))
    e = bad[0][0] if len(bad[0]) else 0
    print("env", e, "root soa", a["root"][e], "\n oracle", o["root"][e])
    print("jw soa", a["dof"][e, :12, 1], "\n oracle", o["dof"][e, :12, 1])
