"""GPU: device-side trajectory reset (emloco_traj_reset) against goldens produced by the REFERENCE's
TrajGenerator.reset on recorded uniform draws (oracle/make_golden.py:reference_traj_reset), plus the numpy oracle for the
LocoVal inputs captured at reset and properties of the Philox stream.  Tolerance: 1e-5 relative to the polyline's extent
(fp32 cumulative sums in a different association order); masks bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _sim(n):
    from emloco_b200.sim import EmlocoSim
    return EmlocoSim(n, device=0)


def _flag_reset(sim, env_ids):
    sim.reset.zero_()
    sim.reset[torch.as_tensor(env_ids, device="cuda", dtype=torch.long)] = 1


@pytest.mark.parametrize("name", ["traj_reset_plain.npz", "traj_reset_train.npz", "traj_reset_inv.npz"])
def test_traj_reset_matches_reference_golden(name):
    from oracle import oracle_np as O
    g = np.load(os.path.join(GOLDEN, name))
    N, ids = g["verts0"].shape[0], g["env_ids"]
    sim = _sim(N)
    sim.traj_verts.copy_(torch.from_numpy(g["verts0"]).cuda())
    root = sim.root_state.view(N, 13)
    root[torch.from_numpy(ids).cuda(), 0:3] = torch.from_numpy(g["init_pos"]).cuda()
    root[torch.from_numpy(ids).cuda(), 7:10] = torch.from_numpy(g["root_vel"]).cuda()
    rb = torch.randn(N, 24, 13, device="cuda")
    sim.rb_state.view(N, 24, 13).copy_(rb)
    U = torch.rand(N, 408, device="cuda")
    U[torch.from_numpy(ids).cuda(), :405] = torch.from_numpy(g["U"]).cuda()
    way = torch.full((N, 15, 3), -7.0, device="cuda"); pose = torch.full((N, 24, 3), -7.0, device="cuda")
    vel = torch.full((N, 2), -7.0, device="cuda"); inv = torch.full((N,), 9, device="cuda", dtype=torch.uint8)
    cfg = sim.traj_cfg(flags=int(g["flags"]), pool=torch.from_numpy(g["pool"]).cuda().contiguous(), uniform=U,
                       waypoint_traj=way, init_pose=pose, init_vel=vel, inverted=inv)
    _flag_reset(sim, ids)
    sim.traj_reset(cfg)
    torch.cuda.synchronize()
    v = sim.traj_verts.cpu().numpy()
    scale = np.abs(g["verts"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(v - g["verts"]) <= 1e-5 * scale + 2e-6).all(), np.abs(v - g["verts"]).max()
    others = np.setdiff1d(np.arange(N), ids)
    np.testing.assert_array_equal(v[others], g["verts0"][others])                  # untouched envs, bit for bit
    np.testing.assert_array_equal(inv.cpu().numpy()[ids], g["inverted"].astype(np.uint8))
    assert (inv.cpu().numpy()[others] == 9).all() and (way.cpu().numpy()[others] == -7).all()
    assert (sim.reset.cpu().numpy()[ids] == 1).all()                           # the explicit call leaves the flags alone
    w_o, p_o, v_o = O.reset_task_outputs(g["verts"].copy(), ids, rb.cpu().numpy()[ids][:, :, 0:3], g["root_vel"])
    ws = np.abs(w_o).max(axis=(1, 2), keepdims=True)
    assert (np.abs(way.cpu().numpy()[ids] - w_o) <= 2e-5 * np.maximum(ws, scale[ids]) + 2e-6).all()
    np.testing.assert_allclose(pose.cpu().numpy()[ids], p_o, rtol=0, atol=1e-6)
    np.testing.assert_array_equal(vel.cpu().numpy()[ids], v_o)
    sim.close()


def test_philox_stream_properties():
    """No explicit draws: deterministic per (seed, env, reset count), bounded segment lengths, headings spread over the
    circle, a second reset of the same env gives a new polyline, another seed gives another one."""
    N = 512
    sim = _sim(N)
    root = sim.root_state.view(N, 13)
    root[:, 0:2] = 50 + 8 * torch.rand(N, 2, device="cuda")
    root[:, 7:9] = torch.randn(N, 2, device="cuda")

    def run(seed):
        cfg = sim.traj_cfg(flags=0, seed=seed)
        _flag_reset(sim, np.arange(N))
        sim.traj_reset(cfg)
        torch.cuda.synchronize()
        return sim.traj_verts.cpu().numpy().copy()
    a = run(5)
    b = run(5)                    # reset count advanced -> different draws
    sim2 = _sim(N)
    sim2.root_state.copy_(sim.root_state)
    cfg = sim2.traj_cfg(flags=0, seed=5)
    _flag_reset(sim2, np.arange(N)); sim2.traj_reset(cfg); torch.cuda.synchronize()
    a2 = sim2.traj_verts.cpu().numpy()
    np.testing.assert_array_equal(a, a2)                                           # same seed, same count: bit-identical
    assert np.abs(a - b)[:, 1:].max() > 0.1
    _flag_reset(sim2, np.arange(N)); sim2.traj_reset(sim2.traj_cfg(flags=0, seed=6)); torch.cuda.synchronize()
    assert np.abs(sim2.traj_verts.cpu().numpy() - b)[:, 1:].max() > 0.1
    np.testing.assert_allclose(a[:, 0, :2], root[:, 0:2].cpu().numpy(), rtol=0, atol=0)
    seg = np.linalg.norm(np.diff(a[:, :, :2], axis=1), axis=-1)
    dt = 168 * (2 / 60) / 100
    assert seg.max() <= 3.0 * dt * (1 + 1e-4) and seg.min() >= 0.0005 * dt - 1e-5     # positions ~50 m: fp32 ulp 4e-6
    head = np.arctan2(a[:, 1, 1] - a[:, 0, 1], a[:, 1, 0] - a[:, 0, 0])
    hist = np.histogram(head, bins=8, range=(-np.pi, np.pi))[0]
    assert hist.min() > N / 8 * 0.5 and hist.max() < N / 8 * 1.6                   # U(-pi, pi) initial heading
    sp0 = seg[:, 0] / dt
    assert 1.2 < sp0.mean() < 1.8                                                  # U(0.0005, 3)
    turn = np.abs(np.diff(np.unwrap(np.arctan2(np.diff(a[:, :, 1], axis=1), np.diff(a[:, :, 0], axis=1)), axis=1), axis=1))
    sharp = (turn > 2.0 * dt * 1.01).mean()
    assert 0.005 < sharp < 0.04                                                    # 2 % sharp turns
    sim.close(); sim2.close()


def test_reset_done_runs_traj_reset_after_the_observations():
    """humanoid_amp_task.py:54-57: observations of a reset env are computed from the OLD polyline, then _reset_task replaces
    it; flags are cleared by the last stage; envs that did not reset keep their polyline."""
    from emloco_b200.rollout import Rollout
    n = 64
    R = Rollout(n, seed=1, tensor_cores=False, fuse_sinks=False, concurrent=False)
    sim = R.sim
    ids = np.array([3, 17, 40])
    old = sim.traj_verts.clone()
    # reference order on a twin without the stage: the reset observation to compare with
    _flag_reset(sim, ids)
    sim.reset_done(R.init_root, R.init_dof)
    torch.cuda.synchronize()
    obs_plain = sim.obs.clone()
    way = torch.zeros(n, 15, 3, device="cuda"); pose = torch.zeros(n, 24, 3, device="cuda"); vel = torch.zeros(n, 2, device="cuda")
    sim.set_traj_reset(sim.traj_cfg(flags=4, seed=2, waypoint_traj=way, init_pose=pose, init_vel=vel))
    _flag_reset(sim, ids)
    sim.reset_done(R.init_root, R.init_dof)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(sim.obs.cpu().numpy(), obs_plain.cpu().numpy())       # obs still from the old polyline
    v = sim.traj_verts.cpu().numpy()
    others = np.setdiff1d(np.arange(n), ids)
    np.testing.assert_array_equal(v[others], old.cpu().numpy()[others])
    assert np.abs(v[ids] - old.cpu().numpy()[ids])[:, 1:].max() > 0.05
    assert int(sim.reset.sum()) == 0 and int(sim.terminate.sum()) == 0
    # --init_heading: first segment along the root velocity
    rv = R.init_root[torch.from_numpy(ids).cuda(), 7:9].cpu().numpy()
    d = v[ids, 1, :2] - v[ids, 0, :2]
    cosang = (d * rv).sum(1) / (np.linalg.norm(d, axis=1) * np.linalg.norm(rv, axis=1))
    assert (cosang > 1 - 1e-4).all()
    np.testing.assert_allclose(way.cpu().numpy()[ids, 0], 0, atol=0)
    np.testing.assert_array_equal(vel.cpu().numpy()[ids], rv)
    rb = sim.rb_state.view(n, 24, 13)[:, :, 0:3].cpu().numpy()
    np.testing.assert_allclose(pose.cpu().numpy()[ids], rb[ids] - rb[ids][:, :1], atol=1e-6)
    sim.set_traj_reset(None)
    R.close()


def test_rollout_with_device_traj_reset_deferred_equals_inline():
    """The deferred stage (overlapping the policy pass on a side branch) gives the same rollout as the inline stage."""
    from emloco_b200.rollout import Rollout
    from emloco_b200.synthetic import synthetic_traj_pool
    n, outs = 128, []
    pool = synthetic_traj_pool(16, 0)
    for deferred in (False, True):
        R = Rollout(n, seed=2, tensor_cores=True, concurrent=True, traj_deferred=deferred, traj_flags=1 | 2 | 4 | 8, traj_pool=pool,
                    horizon=8)
        assert R._traj_deferred == deferred
        R.sim.progress.fill_(160)                     # every env times out inside the horizon
        R.sim.progress[: n // 2] = 163
        torch.manual_seed(0)
        for k in range(8):
            R.step(k, noise=torch.randn(n, 69, device="cuda", generator=torch.Generator(device="cuda").manual_seed(k)))
        torch.cuda.synchronize()
        outs.append(dict(verts=R.sim.traj_verts.cpu().numpy().copy(), way=R.waypoint_traj.cpu().numpy().copy(),
                         pose=R.init_pose.cpu().numpy().copy(), vel=R.init_vel.cpu().numpy().copy(), inv=R.inverted.cpu().numpy().copy(),
                         dones=R.mb["dones"].cpu().numpy().copy(), rewards=R.mb["rewards"].cpu().numpy().copy(),
                         obs=R.mb["obses"].cpu().numpy().copy()))
        R.close()
    a, b = outs
    assert a["dones"].sum() >= n                                       # the resets happened
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert 0 < a["inv"].sum() < n                                      # heading inversion drew both ways
