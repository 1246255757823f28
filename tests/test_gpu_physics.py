"""GPU: physics kernel vs its fp64 CPU restatement (oracle/physics_oracle.c) on identical inputs, plus
physical invariants.  PARITY AGAINST PhysX IS UNPINNED (the reference ships neither binaries nor tests for
the physics step) - these tests pin the kernel to the builder's own algorithm."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup():
    from emloco_b200.model import build_model_arrays, rest_root_height
    from oracle import physics_oracle as PO
    A = build_model_arrays()
    M = PO.make_model(A["parent"], A["offset"], A["mass"], A["com"], A["inertia6"], A["kp_joint"], A["kd_joint"],
                      A["arm_joint"], A["geom_type"], A["geom_a"], A["geom_b"], A["geom_r"])
    return A, M, PO, rest_root_height(A)


def _random_state(N, seed, z_lo, z_hi, vel=1.0, pose=0.4):
    rng = np.random.default_rng(seed)
    root = np.zeros((N, 13))
    root[:, 0:2] = rng.uniform(52, 56, (N, 2))
    root[:, 2] = rng.uniform(z_lo, z_hi, N)
    yaw = rng.uniform(-np.pi, np.pi, N)
    tilt = rng.normal(0, 0.05, (N, 2))
    q = np.stack([tilt[:, 0], tilt[:, 1], np.sin(yaw / 2), np.cos(yaw / 2)], -1)
    root[:, 3:7] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    root[:, 7:13] = rng.normal(0, vel, (N, 6))
    dof_pos = rng.uniform(-pose, pose, (N, 69))
    dof_vel = rng.normal(0, vel, (N, 69))
    actions = rng.uniform(-0.3, 0.3, (N, 69))
    return root, dof_pos, dof_vel, actions


def _gpu_step(N, root, dof_pos, dof_vel, actions, steps=1, impl=0):
    from emloco_b200.sim import EmlocoSim
    sim = EmlocoSim(N, physics_impl=impl)
    sim.root_state.copy_(torch.from_numpy(root).float().cuda())
    ds = np.stack([dof_pos, dof_vel], -1).reshape(N * 69, 2)
    sim.dof_state.copy_(torch.from_numpy(ds).float().cuda())
    sim.reset_indexed(None)
    act = torch.from_numpy(actions).float().cuda().contiguous()
    for _ in range(steps):
        sim.step(act)
    torch.cuda.synchronize()
    out = dict(root=sim.root_state, rb=sim.rb_state.reshape(N, 24, 13), dof=sim.dof_state.reshape(N, 69, 2),
               contact=sim.contact.reshape(N, 24, 3), dof_force=sim.dof_force.reshape(N, 69), pd=sim.pd_target)
    out = {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
    sim.close()
    return out


def _oracle_step(A, M, PO, root, dof_pos, dof_vel, actions, steps=1):
    from oracle import oracle_np as O
    N = root.shape[0]
    cfg = PO.make_cfg(1.0 / 120.0)
    r = root.astype(np.float32).astype(np.float64).copy()
    jq = PO.expmap_to_quat(dof_pos.astype(np.float32).astype(np.float64)).reshape(N, 23, 4).copy()
    jw = dof_vel.astype(np.float32).astype(np.float64).copy()
    tgt = O.action_to_pd_targets(actions.astype(np.float32), A["pd_offset"], A["pd_scale"]).astype(np.float64)
    for _ in range(steps):
        rb, dp, ct, df = PO.step(M, cfg, 4, r, jq, jw, tgt)
    return dict(root=r, rb=rb, dof=np.stack([dp, jw], -1), contact=ct, dof_force=df, pd=tgt)


@pytest.mark.parametrize("impl", [0, 1])     # 0: lane-per-env kernel (physics_soa.cu), 1: warp-per-env kernel (physics.cu)
@pytest.mark.parametrize("case,z_lo,z_hi", [("airborne", 2.0, 3.0), ("contact", 0.80, 0.93)])
def test_single_env_step_matches_fp64_oracle(case, z_lo, z_hi, impl):
    A, M, PO, h0 = _setup()
    N = 250                                   # not a multiple of 32: exercises the masked tail lanes of the lane-per-env kernel
    root, dof_pos, dof_vel, actions = _random_state(N, 7 if case == "airborne" else 8, z_lo, z_hi)
    g = _gpu_step(N, root, dof_pos, dof_vel, actions, impl=impl)
    o = _oracle_step(A, M, PO, root, dof_pos, dof_vel, actions)
    np.testing.assert_allclose(g["pd"], o["pd"], rtol=1e-6, atol=1e-6)
    # 1e-3 relative (north_star) with an absolute floor scaled to each quantity's magnitude
    np.testing.assert_allclose(g["root"], o["root"], rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(g["rb"], o["rb"], rtol=1e-3, atol=5e-3)
    np.testing.assert_allclose(g["dof"], o["dof"], rtol=1e-3, atol=5e-3)
    np.testing.assert_allclose(g["dof_force"], o["dof_force"], rtol=2e-3, atol=0.5)
    np.testing.assert_allclose(g["contact"], o["contact"], rtol=5e-3, atol=2.0)
    if case == "contact":
        assert (np.abs(o["contact"]).sum(axis=(1, 2)) > 0).mean() > 0.5, "test must exercise contacts"


def test_fast_spin_triggers_refinement_and_matches_oracle():
    """Envs spinning faster than max_turn / dt per sub-step are refined (per-env piece count, CTA loops to the largest)."""
    A, M, PO, h0 = _setup()
    N = 96
    root, dof_pos, dof_vel, actions = _random_state(N, 11, 2.0, 3.0, vel=1.0)
    root[::3, 10:13] *= 25.0                  # every third env: ~40-70 rad/s root spin -> 2-3 pieces, its neighbours 1
    o = _oracle_step(A, M, PO, root, dof_pos, dof_vel, actions)
    for impl in (0, 1):
        g = _gpu_step(N, root, dof_pos, dof_vel, actions, impl=impl)
        np.testing.assert_allclose(g["root"], o["root"], rtol=2e-3, atol=5e-3)
        np.testing.assert_allclose(g["rb"], o["rb"], rtol=2e-3, atol=2e-2)
        np.testing.assert_allclose(g["dof"], o["dof"], rtol=2e-3, atol=2e-2)


def test_two_kernels_agree_over_a_rollout():
    """The two GPU mappings are the same arithmetic: 20 env steps with contacts stay within fp32 round-off of each other."""
    A, M, PO, h0 = _setup()
    N = 200
    root, dof_pos, dof_vel, actions = _random_state(N, 5, h0, h0 + 0.05, vel=0.3, pose=0.2)
    a = _gpu_step(N, root, dof_pos, dof_vel, actions * 0.3, steps=5, impl=0)
    b = _gpu_step(N, root, dof_pos, dof_vel, actions * 0.3, steps=5, impl=1)
    for k in ("root", "rb", "dof"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-3, atol=5e-3, err_msg=k)
    np.testing.assert_allclose(a["contact"], b["contact"], rtol=1e-2, atol=3.0)


def test_free_fall_and_rest_pose_invariants():
    A, M, PO, h0 = _setup()
    N = 64
    root = np.zeros((N, 13)); root[:, 6] = 1; root[:, 0:2] = 54.0
    # free fall from 5 m, zero action: semi-implicit Euler gives z = z0 - g/2 * t * (t + dt)
    root[:, 2] = 5.0
    z = np.zeros((N, 69))
    g = _gpu_step(N, root, z, z, z, steps=10)
    t, dt = 10 * 4 / 120.0, 1 / 120.0
    np.testing.assert_allclose(g["root"][:, 2], 5.0 - 0.5 * 9.81 * t * (t + dt), rtol=1e-4)
    np.testing.assert_allclose(g["root"][:, 9], -9.81 * t, rtol=1e-4)
    assert np.abs(g["dof"][..., 0]).max() < 1e-3            # joints stay at the PD target
    assert np.abs(g["contact"]).max() == 0
    # standing on flat ground for 1 s: stays upright, ground carries the weight
    root[:, 2] = h0 + 0.005
    g = _gpu_step(N, root, z, z, z, steps=30)
    assert np.all(g["root"][:, 2] > h0 - 0.05)
    np.testing.assert_allclose(g["contact"][:, :, 2].sum(1), A["total_mass"] * 9.81, rtol=0.05)
    assert np.all(np.isfinite(g["rb"]))


def test_rollout_stays_finite_with_random_actions():
    A, M, PO, h0 = _setup()
    N = 512
    root, dof_pos, dof_vel, actions = _random_state(N, 3, h0, h0 + 0.1, vel=0.5, pose=0.2)
    actions = np.random.default_rng(0).uniform(-1, 1, (N, 69))
    g = _gpu_step(N, root, dof_pos, dof_vel, actions, steps=60)
    assert np.all(np.isfinite(g["rb"])) and np.all(np.isfinite(g["dof"]))
    assert np.abs(g["rb"][..., 0:3] - g["rb"][:, :1, 0:3]).max() < 2.5   # bodies stay attached
    assert np.abs(g["dof"][..., 1]).max() <= 100.0 * np.sqrt(3) + 1e-3  # max_ang_vel clamp


def test_reset_indexed_only_touches_listed_envs():
    from emloco_b200.sim import EmlocoSim
    N = 16
    sim = EmlocoSim(N)
    sim.root_state[:, 2] = 1.0
    sim.reset_indexed(None)
    torch.cuda.synchronize()
    before = sim.rb_state.clone()
    sim.root_state[:, 0] = 3.0
    sim.dof_state.view(N, 69, 2)[:, 3, 0] = 0.5
    sim.reset_indexed(torch.tensor([2, 5]))
    torch.cuda.synchronize()
    after = sim.rb_state.view(N, 24, 13)
    changed = (after != before.view(N, 24, 13)).flatten(1).any(1).cpu().numpy()
    assert changed.tolist() == [i in (2, 5) for i in range(N)]
    assert abs(after[2, 0, 0].item() - 3.0) < 1e-6
    sim.close()


@pytest.mark.parametrize("impl", [0, 1])
def test_per_env_body_models_match_fp64_oracle_per_shape(impl):
    """SURVEY 8 row f3 (has_shape_variation): three body models (0.88x, 1x, 1.15x: lengths, masses, inertias, mass-scaled PD
    gains) assigned env i -> shape i % 3 as the reference does (humanoid.py:606); every env must step like the fp64 oracle
    run with ITS shape's model - in the air and in contact - and differently from the shared-model step.  Switching the per-env
    models off restores the shared-model result bit for bit."""
    from emloco_b200.model import rest_root_height, scaled_model_arrays
    from emloco_b200.sim import EmlocoSim
    A, M0, PO, h0 = _setup()
    shapes = [scaled_model_arrays(A, s) for s in (0.88, 1.0, 1.15)]
    oracles = [PO.make_model(S["parent"], S["offset"], S["mass"], S["com"], S["inertia6"], S["kp_joint"], S["kd_joint"], S["arm_joint"],
                             S["geom_type"], S["geom_a"], S["geom_b"], S["geom_r"]) for S in shapes]
    N = 250
    for case, z_lo, z_hi in (("airborne", 2.0, 3.0), ("contact", 0.72, 1.02)):
        root, dof_pos, dof_vel, actions = _random_state(N, 21 if case == "airborne" else 22, z_lo, z_hi)
        sim = EmlocoSim(N, physics_impl=impl)
        betas = np.concatenate([np.zeros((3, 1)), np.linspace(-1, 1, 3)[:, None] * np.ones((3, 16))], 1)
        sim.set_env_models(shapes, betas=betas)
        idx = np.arange(N) % 3

        def run():
            sim.root_state.copy_(torch.from_numpy(root).float().cuda())
            sim.dof_state.copy_(torch.from_numpy(np.stack([dof_pos, dof_vel], -1).reshape(N * 69, 2)).float().cuda())
            sim.reset_indexed(None)
            sim.step(torch.from_numpy(actions).float().cuda().contiguous())
            torch.cuda.synchronize()
            return {k: v.cpu().numpy().astype(np.float64) for k, v in dict(root=sim.root_state, rb=sim.rb_state.reshape(N, 24, 13),
                    dof=sim.dof_state.reshape(N, 69, 2), contact=sim.contact.reshape(N, 24, 3), obs=sim.obs).items()}
        g = run()
        for k in range(3):
            sel = idx == k
            o = _oracle_step(shapes[k], oracles[k], PO, root[sel], dof_pos[sel], dof_vel[sel], actions[sel])
            np.testing.assert_allclose(g["root"][sel], o["root"], rtol=1e-3, atol=2e-3, err_msg=f"{case} shape {k}")
            np.testing.assert_allclose(g["rb"][sel], o["rb"], rtol=1e-3, atol=5e-3, err_msg=f"{case} shape {k}")
            np.testing.assert_allclose(g["dof"][sel], o["dof"], rtol=1e-3, atol=5e-3, err_msg=f"{case} shape {k}")
            np.testing.assert_allclose(g["contact"][sel], o["contact"], rtol=5e-3, atol=2.0, err_msg=f"{case} shape {k}")
            np.testing.assert_allclose(g["obs"][sel, 357:368], np.broadcast_to(betas[k, :11], (sel.sum(), 11)), atol=1e-6)   # shape obs
        # the shapes matter: envs of shape 0 / 2 deviate from what the shared model gives, shape 1 does not
        sim.set_env_models(None)
        shared = run()
        assert np.abs(shared["rb"][idx == 0] - g["rb"][idx == 0]).max() > 1e-2
        np.testing.assert_array_equal(shared["rb"][idx == 1], g["rb"][idx == 1])
        ref = _gpu_step(N, root, dof_pos, dof_vel, actions, impl=impl)
        np.testing.assert_array_equal(shared["rb"], ref["rb"])
        if case == "contact":
            assert (np.abs(g["contact"]).sum(axis=(1, 2)) > 0).mean() > 0.3
        sim.close()
