"""GPU parity tests: the CUDA path (through the C ABI) against the committed golden fixtures (produced
by the REFERENCE's own functions) and against the oracle on seeded inputs.
Tolerances: north_star says 1e-3 relative on floats, bit-exact on reset masks / indices."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 2e-5


def _sim(N, **kw):
    from emloco_b200.sim import EmlocoSim
    return EmlocoSim(N, **kw)


def _load_state(sim, st):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    N = st["rb"].shape[0]
    sim.rb_state.copy_(T(st["rb"]).reshape(N * 24, 13))
    sim.dof_state.copy_(T(st["dof_state"]).reshape(N * 69, 2))
    sim.contact.copy_(T(st["contact"]).reshape(N * 24, 3))
    sim.dof_force.copy_(T(st["dof_force"]).reshape(-1))
    sim.progress.copy_(T(st["progress"]))
    sim.betas.copy_(T(st["betas"]))
    sim.traj_verts.copy_(T(st["verts"]))
    sim.amp_obs.copy_(T(st["amp_buf"]))
    sim.set_height_field(st["height_samples"])


def _run_post(st):
    N = st["rb"].shape[0]
    sim = _sim(N)
    _load_state(sim, st)
    sim.post_step(advance_progress=False)
    torch.cuda.synchronize()
    out = dict(obs=sim.obs, flip_obs=sim.flip_obs, rew=sim.rew, reward_raw=sim.rew_raw, reset=sim.reset,
               terminate=sim.terminate, amp_obs=sim.amp_obs.reshape(N, -1))
    out = {k: v.cpu().numpy().copy() for k, v in out.items()}
    sim.close()
    return out


def _height_mismatch_ok(got, want, max_frac=2e-3):
    """Height samples are a discontinuous lookup: an ulp difference in sin/cos can move a grid point across a
    cell edge.  Allow a tiny fraction of mismatching samples, everything else to tolerance."""
    bad = ~np.isclose(got, want, rtol=RTOL, atol=ATOL)
    return bad.mean() <= max_frac


@pytest.mark.parametrize("name", ["post_step_rough.npz", "post_step_flat.npz"])
def test_post_step_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    st = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    out = _run_post(st)
    for k in ("rew", "reward_raw", "amp_obs"):
        np.testing.assert_allclose(out[k], g["out_" + k], rtol=RTOL, atol=ATOL, err_msg=k)
    for k in ("obs", "flip_obs"):
        np.testing.assert_allclose(out[k][:, :398], g["out_" + k][:, :398], rtol=RTOL, atol=ATOL, err_msg=k)
        assert _height_mismatch_ok(out[k][:, 398:], g["out_" + k][:, 398:]), k + " heights"
    for k in ("reset", "terminate"):
        assert out[k].dtype == np.int64
        np.testing.assert_array_equal(out[k], g["out_" + k])


@pytest.mark.parametrize("N,rough", [(1, True), (257, True), (4096, False)])
def test_post_step_matches_oracle(N, rough):
    from oracle import oracle_np as O
    from oracle.make_golden import synth_state
    st = synth_state(N, seed=100 + N, map_shape=(700, 700), rough=rough)
    out = _run_post(st)
    n = min(N, 512)     # the numpy oracle is the slow side; full-size properties are checked below
    sub = {k: (v[:n] if k != "height_samples" else v) for k, v in st.items()}
    ref = O.post_physics_step(sub["rb"], sub["dof_state"], sub["contact"], sub["dof_force"], sub["progress"],
                              sub["verts"], sub["betas"], sub["height_samples"], sub["amp_buf"])
    for k in ("rew", "reward_raw", "amp_obs"):
        np.testing.assert_allclose(out[k][:n], ref[k], rtol=RTOL, atol=ATOL, err_msg=k)
    for k in ("obs", "flip_obs"):
        np.testing.assert_allclose(out[k][:n, :398], ref[k][:, :398], rtol=RTOL, atol=ATOL, err_msg=k)
        assert _height_mismatch_ok(out[k][:n, 398:], ref[k][:, 398:]), k
    np.testing.assert_array_equal(out["reset"][:n], ref["reset"])
    np.testing.assert_array_equal(out["terminate"][:n], ref["terminate"])
    # size-independent properties at full size
    obs, flip = out["obs"], out["flip_obs"]
    np.testing.assert_array_equal(flip[:, 357:368], obs[:, 357:368])                       # shape params untouched
    np.testing.assert_array_equal(flip[:, 368:398:2], obs[:, 368:398:2])                   # traj x kept
    np.testing.assert_array_equal(flip[:, 369:398:2], -obs[:, 369:398:2])                  # traj y negated
    hm = obs[:, 398:].reshape(N, 32, 32)
    np.testing.assert_array_equal(flip[:, 398:].reshape(N, 32, 32), hm[:, :, ::-1])        # heightmap flipped
    amp = out["amp_obs"].reshape(N, 15, 206)
    np.testing.assert_array_equal(amp[:, 1:], st["amp_buf"][:, :-1])                       # ring shifted by one
    assert set(np.unique(out["reset"])) <= {0, 1}
    assert np.all(out["reset"] >= out["terminate"])


def test_locoval_matches_reference_golden():
    from emloco_b200.value_pose_net import ValuePoseNet
    g = np.load(os.path.join(GOLDEN, "locoval.npz"))
    net = ValuePoseNet(True, True).cuda().eval()
    net.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")})
    traj = torch.from_numpy(g["traj"]).cuda().requires_grad_(True)
    pose = torch.from_numpy(g["pose"]).cuda()
    vel = torch.from_numpy(g["vel"]).cuda()
    value, loss = net.calc_embodied_motion_loss(traj, pose, vel)
    loss.backward()
    np.testing.assert_allclose(value.detach().cpu().numpy(), g["out_value"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(loss.item(), g["out_loss"], rtol=RTOL)
    np.testing.assert_allclose(traj.grad.cpu().numpy(), g["out_grad_traj"], rtol=2e-3, atol=1e-7)
    # the reference rotates / zeroes the caller's pose in place; so do we
    assert pose.requires_grad            # ... with autograd history, like the reference's in-place bmm (value_pose_net.py:97)
    np.testing.assert_allclose(pose.detach().cpu().numpy(), g["out_pose_after"], rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("use_pose,use_vel,vru", [(True, True, False), (False, False, False), (False, True, False),
                                                  (True, False, False), (True, True, True)])
def test_locoval_variants_match_oracle(use_pose, use_vel, vru):
    from emloco_b200.value_pose_net import ValuePoseNet
    from oracle import oracle_np as O
    torch.manual_seed(0)
    T = 5 if vru else 13
    B = 1000
    net = ValuePoseNet(use_pose, use_vel, vru=vru, mutate_pose=False).cuda().eval()
    with torch.no_grad():
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.2, 0.2)
    traj = torch.randn(B, T, 3).cuda(); traj[:, 0] = 0
    pose = torch.randn(B, 24, 3).cuda() * 0.3
    vel = torch.randn(B, 2).cuda()
    v = net(traj, pose if use_pose else None, vel if use_vel else None).cpu().numpy()
    # oracle (full variant only has a dedicated function; build the others from its pieces)
    W = {k: (getattr(net._network, k).weight.detach().cpu().numpy(), getattr(net._network, k).bias.detach().cpu().numpy())
         for k in ("fc1", "fc2", "fc3")}
    tr, pr, vr, _ = O.locoval_normalize(traj.cpu().numpy()[..., :2], pose.cpu().numpy(), vel.cpu().numpy())
    pr[:, [4, 8]] = 0; pr[:, [9, 10, 11]] = 0
    feats = [tr.reshape(B, -1)] + ([pr.reshape(B, 72)] if use_pose else []) + ([vr] if use_vel else [])
    x = np.concatenate(feats, -1).astype(np.float32)
    h = np.maximum(x @ W["fc1"][0].T + W["fc1"][1], 0); h = np.maximum(h @ W["fc2"][0].T + W["fc2"][1], 0)
    ref = 1 / (1 + np.exp(-(h @ W["fc3"][0].T + W["fc3"][1])))
    np.testing.assert_allclose(v, ref, rtol=RTOL, atol=1e-6)


def test_locoval_edge_cases_and_full_size():
    from emloco_b200.value_pose_net import ValuePoseNet, score_host
    net = ValuePoseNet(True, True).cuda().eval()
    # empty batch
    out = net(torch.zeros(0, 13, 2).cuda(), torch.zeros(0, 24, 3).cuda(), torch.zeros(0, 2).cuda())
    assert out.shape == (0, 1)
    # batch of one (the evaluate_jta.py filter loop shape), first step exactly along +y (x == 0 -> epsilon branch)
    t1 = torch.zeros(1, 13, 2).cuda(); t1[0, :, 1] = torch.arange(13).float() * 0.4
    v1 = net(t1, torch.zeros(1, 24, 3).cuda(), torch.zeros(1, 2).cuda())
    assert torch.isfinite(v1).all() and 0 < v1.item() < 1
    # 1M batch: heading invariance (rotating traj+pose+vel together leaves the score unchanged) - a
    # size-independent property of the normalisation
    B = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(1)
    traj = torch.randn(B, 13, 2, device="cuda", generator=g).cumsum(1) * 0.3; traj -= traj[:, :1].clone()
    pose = torch.randn(B, 24, 3, device="cuda", generator=g) * 0.3
    vel = (traj[:, 1] - traj[:, 0]) * 2.5
    net2 = ValuePoseNet(True, True, mutate_pose=False).cuda().eval()
    a = net2(traj, pose, vel)
    th = torch.rand(B, device="cuda", generator=g) * 6.28
    c, s = th.cos(), th.sin()
    rot = lambda p: torch.stack([p[..., 0] * c[:, None] - p[..., 1] * s[:, None], p[..., 0] * s[:, None] + p[..., 1] * c[:, None]], -1)
    pose_r = pose.clone(); pose_r[..., :2] = rot(pose[..., :2])
    b = net2(rot(traj).contiguous(), pose_r, rot(vel[:, None])[:, 0].contiguous())
    assert (a - b).abs().max().item() < 2e-4
    # host-buffer entry point agrees with the device one
    n = 5000
    sd = {k: v.detach().cpu().numpy() for k, v in net2.state_dict().items()}
    h = score_host(traj[:n].cpu().numpy(), pose[:n].cpu().numpy(), vel[:n].cpu().numpy(), sd)
    np.testing.assert_allclose(h, a[:n, 0].cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_gae_matches_reference_golden_and_oracle():
    from emloco_b200.sim import gae
    from oracle import oracle_np as O
    g = np.load(os.path.join(GOLDEN, "gae.npz"))
    T = lambda a: torch.from_numpy(a).cuda()
    adv, ret = gae(T(g["dones"]), T(g["values"]), T(g["rewards"]), T(g["next_values"]))
    np.testing.assert_allclose(adv.cpu().numpy(), g["adv"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret.cpu().numpy(), g["adv"] + g["values"], rtol=1e-5, atol=1e-6)
    rng = np.random.default_rng(0)
    Th, N = 32, 4096
    d = (rng.uniform(size=(Th, N)) < 0.03).astype(np.float32)
    v, nv, r = (rng.normal(size=(Th, N, 1)).astype(np.float32) for _ in range(3))
    adv, _ = gae(T(d), T(v), T(r), T(nv))
    np.testing.assert_allclose(adv.cpu().numpy(), O.discount_values(d, v, r, nv), rtol=1e-5, atol=1e-5)
    # ragged: a single env, horizon 1
    a1, _ = gae(T(d[:1, :1]), T(v[:1, :1]), T(r[:1, :1]), T(nv[:1, :1]))
    np.testing.assert_allclose(a1.cpu().numpy(), r[:1, :1] + 0.99 * nv[:1, :1] - v[:1, :1], rtol=1e-6)


def test_linear_fma_matches_oracle():
    from emloco_b200.sim import linear
    from oracle import oracle_np as O
    rng = np.random.default_rng(1)
    M, K, N = 300, 1054, 512
    x = rng.normal(0, 2, (M, 1422)).astype(np.float32)
    w = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(0, 0.1, N).astype(np.float32)
    mean = rng.normal(0, 1, 1422); var = rng.uniform(0.1, 4, 1422)
    xt = torch.from_numpy(x).cuda()
    y = linear(xt[:, 368:], torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), relu=True,
               mean=torch.from_numpy(mean[368:]).cuda(), var=torch.from_numpy(var[368:]).cuda())
    ref = np.maximum(O.rms_normalize(x, mean, var)[:, 368:] @ w.T + b, 0)
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,N,K,relu", [(128, 128, 64, False), (300, 512, 1054, True), (4096, 69, 1024, False),
                                        (1000, 1, 512, False), (257, 1024, 3090, True), (64, 256, 624, True)])
def test_linear_bf16x3_tensor_core_matches_fp64(M, N, K, relu):
    """tcgen05 path: fp32 operands split into bf16 hi+lo, three MMAs per k-step.  Checked against float64 numpy; the
    documented error of the split is ~2^-16 relative to sum|a||w|, far inside north_star's 1e-3."""
    from emloco_b200.sim import linear
    rng = np.random.default_rng(M + N + K)
    x = rng.normal(0, 1.5, (M, K)).astype(np.float32)
    w = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(0, 0.1, N).astype(np.float32)
    y = linear(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), relu=relu, tensor_cores=True)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    if relu:
        ref = np.maximum(ref, 0)
    err = np.abs(y.cpu().numpy() - ref)
    scale = np.abs(x).astype(np.float64) @ np.abs(w).astype(np.float64).T + np.abs(b)
    assert (err / scale).max() < 1e-4, (err / scale).max()
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("tile", [128, 256, 0x800 + 128, 0x800 + 256])
def test_linear_bf16x3_split_output_chains_layers(tile):
    """Two chained layers through the split (hi/lo) epilogue output, with input normalisation, against float64;
    both output-tile shapes (128 x 128 x 64 / 128B swizzle and 128 x 256 x 32 / 64B swizzle)."""
    from emloco_b200.policy import _Split, linear_bf16x3, split_bf16
    rng = np.random.default_rng(5)
    M, K, H, N = 500, 1054, 512, 256
    x = rng.normal(0, 2, (M, K)).astype(np.float32)
    mean = rng.normal(0, 1, K).astype(np.float32); var = rng.uniform(0.1, 4, K).astype(np.float32)
    w1 = (rng.normal(0, 1, (H, K)) / np.sqrt(K)).astype(np.float32); b1 = rng.normal(0, 0.1, H).astype(np.float32)
    w2 = (rng.normal(0, 1, (N, H)) / np.sqrt(H)).astype(np.float32); b2 = rng.normal(0, 0.1, N).astype(np.float32)
    T = lambda a: torch.from_numpy(a).cuda()
    sx, s1, sw1, sw2 = _Split(M, K, "cuda"), _Split(M, H, "cuda"), _Split(H, K, "cuda"), _Split(N, H, "cuda")
    split_bf16(T(x), sx, T(mean), T(var), 1e-5); split_bf16(T(w1), sw1); split_bf16(T(w2), sw2)
    y = torch.empty(M, N, device="cuda")
    linear_bf16x3(sx, sw1, T(b1), True, y16=s1, tile=tile)
    linear_bf16x3(s1, sw2, T(b2), True, y32=y, tile=tile)
    xn = np.clip((x.astype(np.float64) - mean) / np.sqrt(var.astype(np.float64) + 1e-5), -5, 5)
    h = np.maximum(xn @ w1.T.astype(np.float64) + b1, 0)
    ref = np.maximum(h @ w2.T.astype(np.float64) + b2, 0)
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-3, atol=1e-4)
    # the split itself: hi + lo reproduces the fp32 value to 2^-16
    rec = (sx.hi.float() + sx.lo.float())[:, :K].cpu().numpy()
    np.testing.assert_allclose(rec, xn, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("tile", [128, 256, 0x800 + 128, 0x800 + 256])      # +0x800: CTA-pair (cta_group::2) kernels
@pytest.mark.parametrize("M,N,K", [(128, 69, 100), (4096, 1024, 2048), (130, 257, 33), (1000, 4096, 624), (300, 512, 1054)])
def test_linear_bf16x3_tiles_and_ragged_shapes(tile, M, N, K):
    """Forced tile shapes on ragged M / N / K (TMA zero fill of the out-of-range rows and of the K tail), fp32 output."""
    from emloco_b200.policy import _Split, linear_bf16x3, split_bf16
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    x = rng.normal(0, 1.5, (M, K)).astype(np.float32)
    w = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.normal(0, 0.1, N).astype(np.float32)
    T = lambda a: torch.from_numpy(a).cuda()
    sx, sw = _Split(M, K, "cuda"), _Split(N, K, "cuda")
    split_bf16(T(x), sx); split_bf16(T(w), sw)
    y = torch.full((M, N), float("nan"), device="cuda")
    linear_bf16x3(sx, sw, T(b), False, y32=y, tile=tile)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("stride,B", [(2, 5000), (3, 1024), (2, 131072 + 77)])
def test_locoval_tensor_core_kernel_matches_cuda_core_kernel_and_oracle(stride, B):
    """Batches >= 1024 of the full variant run on the tcgen05 kernel (locoval_tc.cu); flag bit 6 forces the CUDA-core kernel.
    Both against each other and against the numpy oracle, incl. the in-place pose side effect and ragged tiles."""
    import ctypes as C
    from emloco_b200 import _lib
    from emloco_b200.value_pose_net import ValuePoseNet
    from oracle import oracle_np as O
    torch.manual_seed(3)
    net = ValuePoseNet(True, True).cuda().eval()
    with torch.no_grad():
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.3, 0.3)
    g = torch.Generator(device="cuda").manual_seed(5)
    traj = torch.randn(B, 13, stride, device="cuda", generator=g).cumsum(1) * 0.3
    traj[:, 0] = 0
    traj[::7, 1, 0] = 0.0                                     # x == 0 at the heading waypoint: the 1e-10 epsilon branch (:79-84)
    pose = torch.randn(B, 24, 3, device="cuda", generator=g) * 0.3
    vel = torch.randn(B, 2, device="cuda", generator=g)
    w = net._weights()
    p = lambda t: C.c_void_p(t.data_ptr())
    out = {}
    for name, extra in (("tc", 0), ("cc", 64)):
        v = torch.empty(B, device="cuda"); pp = pose.clone()
        flags = 1 | 2 | 4 | 8 | 16 | 32 | extra
        _lib.check(_lib.load().emloco_locoval_forward(p(traj), stride, 13, p(pp), p(vel), p(w), p(v), B, flags, None), "emloco_locoval_forward")
        torch.cuda.synchronize()
        out[name] = (v.cpu().numpy(), pp.cpu().numpy())
    np.testing.assert_allclose(out["tc"][0], out["cc"][0], rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(out["tc"][1], out["cc"][1], rtol=1e-5, atol=1e-6)       # rotated / zeroed pose written back
    n = min(B, 4096)
    W = {k: (getattr(net._network, k).weight.detach().cpu().numpy(), getattr(net._network, k).bias.detach().cpu().numpy())
         for k in ("fc1", "fc2", "fc3")}
    ref, pose_after = O.locoval_forward(traj[:n, :, :2].cpu().numpy(), pose[:n].cpu().numpy(), vel[:n].cpu().numpy(), W)
    np.testing.assert_allclose(out["tc"][0][:n], ref[:, 0], rtol=1e-3, atol=2e-6)
    np.testing.assert_allclose(out["tc"][1][:n], pose_after, rtol=1e-3, atol=1e-5)


def test_locoval_multimodal_loss_matches_reference_golden():
    """BASELINE configs[4]: train_jta.py:289-299 with MULTI_MODAL = 5, batch 256 - one calc_embodied_motion_loss call per
    mode on the non-contiguous slice pred_trajs[:, :, i] with the SAME init_pose tensor.  The reference mutates that tensor in
    place with autograd history, so the gradient of mode i's loss also reaches the trajectories of the earlier modes through
    the pose; the golden gradient (reference autograd) contains that chain (it is ~19 % of the gradient norm)."""
    from emloco_b200.value_pose_net import ValuePoseNet
    g = np.load(os.path.join(GOLDEN, "locoval_mm5.npz"))
    net = ValuePoseNet(True, True).cuda().eval()
    net.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")})
    B, _, modes, _ = g["pred"].shape
    out = torch.from_numpy(g["pred"]).cuda().requires_grad_(True)
    trajs = torch.cat([torch.zeros(B, 1, modes, 2, device="cuda"), out], dim=1)
    pose = torch.from_numpy(g["pose"]).cuda()
    vel = torch.from_numpy(g["vel"]).cuda()
    losses, values = 0, []
    for i in range(modes):
        val, l = net.calc_embodied_motion_loss(trajs[:, :, i], pose, vel)
        losses = losses + l
        values.append(val.detach().cpu().numpy())
    losses = losses * float(g["weight"]) / modes
    losses.backward()
    np.testing.assert_allclose(np.stack(values, 0), g["out_values"], rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(losses.item(), g["out_loss"], rtol=RTOL)
    np.testing.assert_allclose(pose.detach().cpu().numpy(), g["out_pose_after"], rtol=RTOL, atol=2e-6)
    gr, ref = out.grad.cpu().numpy(), g["out_grad_pred"]
    np.testing.assert_allclose(gr, ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max() * 1e-2)
    assert np.linalg.norm(gr - ref) / np.linalg.norm(ref) < 1e-3


def test_locoval_finetune_step_matches_reference_golden():
    """The `_do_finetune` block of play_steps (amp_continuous_value.py:122-146; AdamW + MSELoss(sum), common_agent.py:94-96)
    as emloco_locoval_train_step, against four rounds run by the reference ValuePoseNet + torch.optim.AdamW.  Round 2 has no
    valid env: no optimiser step.  fc1 columns 3 and 99 are excluded from the tight check: those inputs are zero up to
    round-off after the heading normalisation, so their gradients are noise that Adam normalises to +-lr."""
    from emloco_b200.value_pose_net import ValuePoseNet
    g = np.load(os.path.join(GOLDEN, "locoval_finetune.npz"))
    net = ValuePoseNet(True, True).cuda()
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w0_")})
    net.enable_finetune()
    steps = 0
    for r in range(4):
        traj, pose, vel, gc = (torch.from_numpy(g[f"r{r}_{k}"]).cuda().contiguous() for k in ("traj", "pose", "vel", "gc"))
        pose0 = pose.clone()
        net.finetune_step(traj, pose, vel, gc, min_cum_rewards=float(g["r_min"]), max_cum_rewards=float(g["r_max"]))
        torch.cuda.synchronize()
        n = int(g[f"r{r}_count"]); steps += n > 0
        loss, pred, gt, cnt = net.finetune_stats()
        assert cnt == n and float(gc.abs().sum()) == 0.0 and torch.equal(pose, pose0)
        if n:
            np.testing.assert_allclose(loss * n, g[f"r{r}_loss"], rtol=1e-4)
            np.testing.assert_allclose(pred * n, g[f"r{r}_pred_sum"], rtol=1e-4); np.testing.assert_allclose(gt * n, g[f"r{r}_gt_sum"], rtol=1e-4)
        sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
        for k in sd:
            w, ref = sd[k], g[f"r{r}_w_{k}"]
            if k == "_network.fc1.weight":
                noise = [3, 99]
                assert np.abs(w[:, noise] - ref[:, noise]).max() <= 2 * 1e-3 * max(steps, 1) * 1.01
                w, ref = np.delete(w, noise, 1), np.delete(ref, noise, 1)
            np.testing.assert_allclose(w, ref, rtol=RTOL, atol=5e-6, err_msg=f"round {r} {k}")
        assert int(net._ft["step"].item()) == steps
    # the updated parameters are what the scoring kernels read (same storage), also after a load_state_dict in between
    net.load_state_dict({k: v.clone() for k, v in net.state_dict().items()})
    with torch.no_grad():
        net.eval()
        traj, pose, vel = (torch.from_numpy(g[f"r3_{k}"]).cuda() for k in ("traj", "pose", "vel"))
        v_kernel = net(traj.contiguous(), pose.clone(), vel).cpu().numpy()
    from oracle import oracle_np as O
    W = {k: (g[f"r3_w__network.{k}.weight"], g[f"r3_w__network.{k}.bias"]) for k in ("fc1", "fc2", "fc3")}
    v_ref, _ = O.locoval_forward(g["r3_traj"][..., :2], g["r3_pose"].copy(), g["r3_vel"], W)
    np.testing.assert_allclose(v_kernel, v_ref, rtol=RTOL, atol=1e-5)


def test_locoval_finetune_all_envs_valid_at_once():
    """Step 144 of a synchronised start: every env is valid in the same step (4096 samples, 64 tiles over 64 CTAs); the
    update equals the fp64 oracle's and is bit-reproducible."""
    from emloco_b200.synthetic import synthetic_locoval_batch
    from emloco_b200.value_pose_net import ValuePoseNet
    from oracle import oracle_np as O
    N = 4096
    traj2, pose, vel = synthetic_locoval_batch(N, seed=3)
    traj = np.concatenate([traj2, np.zeros((N, 13, 1), np.float32)], -1)
    rng = np.random.default_rng(0)
    gc = rng.uniform(5, 80, N).astype(np.float32)
    outs = []
    for rep in range(2):
        torch.manual_seed(1)
        net = ValuePoseNet(True, True).cuda()
        sd0 = {k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()}
        net.finetune_step(*(torch.from_numpy(a).cuda().contiguous() for a in (traj, pose, vel, gc)))
        torch.cuda.synchronize()
        outs.append({k: v.detach().cpu().numpy().copy() for k, v in net.state_dict().items()})
    for k in outs[0]:
        np.testing.assert_array_equal(outs[0][k], outs[1][k])
    W = {k: [sd0[f"_network.{k}.weight"].copy(), sd0[f"_network.{k}.bias"].copy()] for k in ("fc1", "fc2", "fc3")}
    opt = dict(step=0, m={k: [0.0, 0.0] for k in W}, v={k: [0.0, 0.0] for k in W})
    O.locoval_finetune_step(W, opt, traj, pose.copy(), vel, gc.copy())
    for k in W:
        w, ref = outs[0][f"_network.{k}.weight"], W[k][0]
        if k == "fc1":
            w, ref = np.delete(w, [3, 99], 1), np.delete(ref, [3, 99], 1)
        np.testing.assert_allclose(w, ref, rtol=RTOL, atol=5e-6)
        np.testing.assert_allclose(outs[0][f"_network.{k}.bias"], W[k][1], rtol=RTOL, atol=5e-6)


@pytest.mark.parametrize("tile", [0, 128, 256])
@pytest.mark.parametrize("M,N,K,with_y", [(4096, 1024, 2048, False), (300, 512, 1024, True), (129, 96, 200, False), (64, 1024, 624, True)])
def test_linear_bf16x3_fused_head_matches_fp64(tile, M, N, K, with_y):
    """The value / logit layer (nn.Linear(N, 1)) fused into the epilogue of the hidden layer that feeds it: per-row partial
    dot products per 64-column group + an ordered reduction.  Same result for every tile shape (bit-identical), with or
    without the hidden layer's own output, and for a compacted row count."""
    from emloco_b200.policy import _Split, linear_bf16x3, split_bf16
    rng = np.random.default_rng(M + N + K)
    x = rng.normal(0, 1.5, (M, K)).astype(np.float32)
    w = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32); b = rng.normal(0, 0.1, N).astype(np.float32)
    T = lambda a: torch.from_numpy(a).cuda()
    head = torch.nn.Linear(N, 1).cuda()
    sx, sw = _Split(M, K, "cuda"), _Split(N, K, "cuda")
    split_bf16(T(x), sx); split_bf16(T(w), sw)
    out = torch.full((M, 1), 7.0, device="cuda"); part = torch.zeros(M, (N + 63) // 64, device="cuda")
    y = torch.empty(M, N, device="cuda") if with_y else None
    linear_bf16x3(sx, sw, T(b), True, y32=y, tile=tile, head=(head, out, part))
    h = np.maximum(x.astype(np.float64) @ w.astype(np.float64).T + b, 0)
    ref = h @ head.weight.detach().cpu().numpy().astype(np.float64).T + head.bias.item()
    scale = np.abs(h) @ np.abs(head.weight.detach().cpu().numpy().astype(np.float64)).T + abs(head.bias.item())
    assert (np.abs(out.cpu().numpy() - ref) / scale).max() < 1e-4
    if with_y:
        np.testing.assert_allclose(y.cpu().numpy(), h, rtol=1e-3, atol=1e-4)
    # tile independence (what lets the compact critic of the value-reuse path stay bit-identical to the full pass)
    out2 = torch.zeros(M, 1, device="cuda")
    linear_bf16x3(sx, sw, T(b), True, tile=128 if tile != 128 else 256, head=(head, out2, part))
    assert torch.equal(out, out2)
    # device-side row count: rows beyond it are left alone
    r = max(M // 3, 1)
    out3 = torch.full((M, 1), -3.0, device="cuda")
    linear_bf16x3(sx, sw, T(b), True, tile=tile, rows=torch.tensor([r], dtype=torch.int32, device="cuda"), head=(head, out3, part))
    assert torch.equal(out3[:r], out[:r])


@pytest.mark.parametrize("use_pose,use_vel,vru", [(True, True, False), (False, False, False), (False, True, False),
                                                  (True, False, False), (True, True, True)])
def test_locoval_variant_gradients_match_float64_autograd(use_pose, use_vel, vru):
    """d loss / d traj of every ValuePoseNet variant, for two consecutive calls on ONE pose tensor (the second call's loss
    reaches the first call's trajectory through the in-place rotated pose), against torch float64 autograd of an independent
    restatement of value_pose_net.py:73-149 on the CPU."""
    from emloco_b200.value_pose_net import ValuePoseNet
    torch.manual_seed(3)
    T, B = (5 if vru else 13), 300
    net = ValuePoseNet(use_pose, use_vel, vru=vru).cuda().eval()
    with torch.no_grad():
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.2, 0.2)
    tr = [torch.randn(B, T, 2) for _ in range(2)]
    for t in tr:
        t[:, 0] = 0
    pose0 = torch.randn(B, 24, 3) * 0.3
    vel = torch.randn(B, 2)

    # --- kernels
    kt = [t.clone().cuda().requires_grad_(True) for t in tr]
    kp = pose0.clone().cuda()
    loss = 0
    for t in kt:
        _, l = net.calc_embodied_motion_loss(t, kp if use_pose else None, vel.cuda() if use_vel else None)
        loss = loss + l
    loss.backward()

    # --- float64 restatement
    W = [(getattr(net._network, k).weight.detach().cpu().double(), getattr(net._network, k).bias.detach().cpu().double())
         for k in ("fc1", "fc2", "fc3")]
    rt = [t.clone().double().requires_grad_(True) for t in tr]
    p = pose0.clone().double()
    ref_loss = 0
    for t in rt:
        x1 = t[:, 1, 0]
        near = x1.abs() < 1e-10
        x1 = x1 * (~near) + near * 1e-10
        ang = torch.atan2(t[:, 1, 1], x1)
        c, s = torch.cos(ang), torch.sin(ang)
        R = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
        feats = [torch.bmm(t, R).reshape(B, -1)]
        if use_pose:
            p = torch.cat([torch.bmm(p[:, :, :2], R), p[:, :, 2:]], -1)          # functional form of the in-place update
            mask = torch.ones(24, 1, dtype=torch.float64); mask[[4, 8, 9, 10, 11]] = 0
            p = p * mask
            feats.append(p.reshape(B, 72))
        if use_vel:
            feats.append(torch.bmm(vel.double().unsqueeze(1), R)[:, 0])
        h = torch.relu(torch.cat(feats, -1) @ W[0][0].T + W[0][1])
        h = torch.relu(h @ W[1][0].T + W[1][1])
        v = torch.sigmoid(h @ W[2][0].T + W[2][1])
        ref_loss = ref_loss + ((v - 1) ** 2).mean()
    ref_loss.backward()
    np.testing.assert_allclose(loss.item(), ref_loss.item(), rtol=RTOL)
    for a, b in zip(kt, rt):
        g, r = a.grad.cpu().numpy(), b.grad.numpy()
        assert np.abs(g - r).max() <= 2e-3 * np.abs(r).max() + 1e-9, (np.abs(g - r).max(), np.abs(r).max())
    if use_pose:
        np.testing.assert_allclose(kp.detach().cpu().numpy(), p.detach().numpy(), rtol=RTOL, atol=2e-6)


@pytest.mark.parametrize("tile", [0, 128, 256])
def test_linear_bf16x3_split_k_and_sample_actions_parts(tile):
    """Skinny layer (the mu head: 4096 x 69 x 1024) with split-K: the partial matrices add up to the fp64 product; the action
    sampler that consumes the partials gives the same mu / actions / neglogp as the unsplit path."""
    from emloco_b200.policy import _Split, linear_bf16x3, sample_actions, sample_actions_parts, split_bf16
    rng = np.random.default_rng(7)
    M, N, K, S = 1000, 69, 1024, 4
    x = rng.normal(0, 1.5, (M, K)).astype(np.float32)
    w = (rng.normal(0, 1, (N, K)) / np.sqrt(K)).astype(np.float32); b = rng.normal(0, 0.1, N).astype(np.float32)
    T = lambda a: torch.from_numpy(a).cuda()
    sx, sw = _Split(M, K, "cuda"), _Split(N, K, "cuda")
    split_bf16(T(x), sx); split_bf16(T(w), sw)
    parts = torch.full((S, M, N), 9.0, device="cuda")
    linear_bf16x3(sx, sw, T(b), False, y32=parts[0], tile=tile, splits=S)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    scale = np.abs(x).astype(np.float64) @ np.abs(w).astype(np.float64).T + np.abs(b)
    got = parts.sum(0).cpu().numpy()
    assert (np.abs(got - ref) / scale).max() < 1e-4
    y = torch.empty(M, N, device="cuda")
    linear_bf16x3(sx, sw, T(b), False, y32=y, tile=tile)
    np.testing.assert_allclose(got, y.cpu().numpy(), rtol=1e-5, atol=1e-5)
    logstd = torch.full((N,), -2.9, device="cuda"); noise = torch.randn(M, N, device="cuda")
    mu_o = torch.empty(M, N, device="cuda"); a1 = torch.empty(M, N, device="cuda"); n1 = torch.empty(M, device="cuda")
    sample_actions_parts(parts, mu_o, logstd, noise, a1, n1)
    a2, n2 = sample_actions(mu_o, logstd, noise)
    np.testing.assert_allclose(mu_o.cpu().numpy(), got, rtol=1e-6, atol=1e-6)
    assert torch.equal(a1, a2) and torch.allclose(n1, n2, rtol=1e-5, atol=1e-4)
    with pytest.raises(Exception):
        linear_bf16x3(sx, sw, T(b), True, y32=parts[0], tile=tile, splits=S)           # ReLU cannot be split


# ---- round 2: fixtures generated from the reference's own network / agent / plausibl classes ---------------------------
def _golden_net(seed):
    """AMPSeptValueNetwork carrying the synthetic parameters the reference network was run with (oracle/netweights.py)."""
    from emloco_b200.policy import AMPSeptValueNetwork
    from oracle import netweights
    sd = netweights.synth_state_dict(seed)
    net = AMPSeptValueNetwork()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return net.cuda(), sd


@pytest.mark.parametrize("tc", [False, True])
def test_networks_match_reference_network_golden(tc):
    """a11-a13 against AMPSeptValueBuilder.Network + the agent's _eval_critic / _calc_disc_rewards / _combine_rewards run
    from the reference tree (tests/golden/nets.npz), FMA and tcgen05 back ends."""
    from emloco_b200.policy import RolloutNets, RunningMeanStd, disc_reward
    from oracle import netweights
    g = np.load(os.path.join(GOLDEN, "nets.npz"))
    net, sd = _golden_net(int(g["seed"]))
    np.testing.assert_array_equal(netweights.checksum(sd), g["weights_checksum"])
    M = g["obs"].shape[0]
    on, an = RunningMeanStd(1422), RunningMeanStd(3090)
    on.running_mean.copy_(torch.from_numpy(g["obs_mean"])); on.running_var.copy_(torch.from_numpy(g["obs_var"]))
    an.running_mean.copy_(torch.from_numpy(g["amp_mean"])); an.running_var.copy_(torch.from_numpy(g["amp_var"]))
    nets = RolloutNets(net, on.cuda(), an.cuda(), M, tensor_cores=tc)
    obs, amp = torch.from_numpy(g["obs"]).cuda(), torch.from_numpy(g["amp_obs"]).cuda()
    noise = torch.zeros(M, 69, device="cuda")
    r = nets.action_values(obs, noise)
    torch.cuda.synchronize()
    np.testing.assert_allclose(r["mus"].cpu().numpy(), g["out_mu"], rtol=RTOL, atol=5e-5)
    np.testing.assert_allclose(r["actions"].cpu().numpy(), g["out_mu"], rtol=RTOL, atol=5e-5)          # zero noise
    np.testing.assert_allclose(r["values"].cpu().numpy(), g["out_value"], rtol=RTOL, atol=5e-5)
    np.testing.assert_allclose(r["task_values"].cpu().numpy(), g["out_task_value"], rtol=RTOL, atol=5e-5)
    np.testing.assert_array_equal(r["sigmas"].cpu().numpy(), g["out_sigma"][0])                        # network-level sigma = log std
    nv = nets.critic(obs).cpu().numpy()
    std = np.sqrt(np.float32(g["value_var"][0]) + np.float32(1e-5))
    np.testing.assert_allclose(std * np.clip(nv, -5, 5) + np.float32(g["value_mean"][0]), g["out_next_value_unnorm"], rtol=RTOL, atol=5e-5)
    logit = nets.disc_logits(amp).clone()
    np.testing.assert_allclose(logit.cpu().numpy(), g["out_disc_logit"], rtol=RTOL, atol=1e-4)
    dr, comb = disc_reward(logit, torch.from_numpy(g["task_rewards"]).cuda(), 2.0, 0.5, 0.5)
    np.testing.assert_allclose(dr.cpu().numpy(), g["out_disc_reward"], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(comb.cpu().numpy(), g["out_combined"], rtol=RTOL, atol=1e-4)


def test_rollout_record_matches_reference_loop_body_golden():
    """a14: emloco_rollout_record against amp_continuous_value.py:61-121 executed from the reference (play_block.npz): the
    inversion penalty, next_vals *= 1 - terminated, AMP reward with its 1e-4 floor, both snapshot conditions."""
    import ctypes as C
    from emloco_b200 import _lib
    from emloco_b200.sim import _ptr, _stream
    g = np.load(os.path.join(GOLDEN, "play_block.npz"))
    N = int(g["N"])
    T = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).cuda()
    state = torch.zeros(6, N, device="cuda")
    state[0] = T(g["state0_current_rewards"][:, 0]); state[1] = T(g["state0_current_lengths"]); state[2] = T(g["state0_current_combined"])
    state[3] = T(g["state0_discount"])
    inv = T(g["inverted"], torch.uint8)
    rcfg = _lib.RolloutCfg(0.3, 1.0, float(g["value_mean"]), float(np.sqrt(np.float32(g["value_var"]) + np.float32(1e-5))), 2.0, 0.99, 144, 1)
    f = lambda *s: torch.zeros(*s, device="cuda")
    for n in range(int(g["steps"])):
        rew, reset, term = T(g["in_rew"][n][:, 0]), T(g["in_reset"][n], torch.int64), T(g["in_terminate"][n], torch.int64)
        nv, logit, pv = T(g["in_critic_raw"][n]), T(g["in_logit"][n]), f(N, 1)
        o_val, o_rew, o_done, o_nv, o_amp = f(N, 1), f(N, 1), f(N), f(N, 1), f(N, 1)
        _lib.check(_lib.load().emloco_rollout_record(C.byref(rcfg), _ptr(rew), _ptr(reset), _ptr(term), _ptr(pv), _ptr(nv), _ptr(logit),
                                                     _ptr(inv), _ptr(o_val), _ptr(o_rew), _ptr(o_done), _ptr(o_nv), _ptr(o_amp),
                                                     _ptr(state), N, _stream()), "emloco_rollout_record")
        torch.cuda.synchronize()
        np.testing.assert_allclose(o_rew.cpu().numpy(), g["row_rewards"][n], rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(o_done.cpu().numpy(), g["row_dones"][n].astype(np.float32))
        np.testing.assert_allclose(o_nv.cpu().numpy(), g["row_next_values"][n], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(o_amp.cpu().numpy()[:, 0], g["row_amp_rewards"][n], rtol=1e-4, atol=1e-6)
        st = state.cpu().numpy()
        for row, ref in ((1, "out_current_lengths"), (2, "out_current_combined"), (3, "out_discount"), (4, "out_game_combined"),
                         (5, "out_terminated_flags")):
            np.testing.assert_allclose(st[row], g[ref][n], rtol=1e-5, atol=1e-5, err_msg=f"step {n} {ref}")
        np.testing.assert_allclose(st[0], g["out_current_rewards"][n][:, 0], rtol=1e-5, atol=1e-5)
        if n == int(g["zero_game_after_step"]):
            state[4].zero_()


def test_post_step_reset_branches_match_reference_golden():
    """a8: every branch of compute_humanoid_reset at 256 envs (post_step_branches.npz), masks bit-exact."""
    from oracle.make_golden import synth_state_branches
    g = np.load(os.path.join(GOLDEN, "post_step_branches.npz"))
    st = synth_state_branches(int(g["N"]), int(g["seed"]))
    out = _run_post(st)
    assert {(int(a), int(b)) for a, b in zip(g["out_reset"], g["out_terminate"])} == {(0, 0), (1, 0), (1, 1)}
    for k in ("reset", "terminate"):
        assert out[k].dtype == np.int64
        np.testing.assert_array_equal(out[k], g["out_" + k])
    for k in ("rew", "reward_raw"):
        np.testing.assert_allclose(out[k], g["out_" + k], rtol=RTOL, atol=ATOL, err_msg=k)
    for k in ("obs", "flip_obs"):
        np.testing.assert_allclose(out[k][:, :398], g["out_" + k][:, :398], rtol=RTOL, atol=ATOL, err_msg=k)
        assert _height_mismatch_ok(out[k][:, 398:], g["out_" + k][:, 398:]), k + " heights"
    np.testing.assert_allclose(out["amp_obs"][:, :206], g["out_amp_obs"], rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(out["amp_obs"][:, 206:], st["amp_buf"][:, :14].reshape(int(g["N"]), -1))


@pytest.mark.parametrize("B", [0, 1, 40, 100000])
def test_plausibl_mlp_matches_reference_class_golden(B):
    """a17: emloco_b200.plausibl.test_value_mlp.MLP against the reference class (plausibl.npz); larger batches against the
    oracle with the same weights; the empty batch returns an empty result."""
    from emloco_b200.plausibl.test_value_mlp import MLP
    from oracle import oracle_np as O
    g = np.load(os.path.join(GOLDEN, "plausibl.npz"))
    m = MLP()
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w_")}
    m._value_mlp.load_state_dict({k[len("_value_mlp."):]: v for k, v in sd.items() if k.startswith("_value_mlp.")})
    m._value_logits.load_state_dict({k[len("_value_logits."):]: v for k, v in sd.items() if k.startswith("_value_logits.")})
    m.cuda()
    if B == 40:
        y = m.forward(torch.from_numpy(g["x"]).cuda())
        np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=1e-5, atol=1e-6)
        return
    x = np.random.default_rng(B).normal(0, 2, (B, 24)).astype(np.float32)
    y = m(torch.from_numpy(x).cuda())
    assert y.shape == (B, 1)
    W = dict(fc1=(g["w__value_mlp.0.weight"], g["w__value_mlp.0.bias"]), fc2=(g["w__value_mlp.2.weight"], g["w__value_mlp.2.bias"]),
             fc3=(g["w__value_logits.weight"], g["w__value_logits.bias"]))
    np.testing.assert_allclose(y.cpu().numpy(), O.plausibl_mlp(x, W), rtol=1e-4, atol=1e-5)


def test_motion_lib_state_and_amp_demo_match_reference_golden():
    """f2: emloco_motion_state / emloco_amp_obs_demo against the reference's motion-library code (motion_lib.npz) and the oracle
    on 5000 fresh samples; reset-state sampling writes valid rows into a Rollout's init buffers."""
    from emloco_b200.motion_lib import MotionLibSMPL
    from emloco_b200.synthetic import synthetic_motion_lib
    from oracle import oracle_np as O
    g = np.load(os.path.join(GOLDEN, "motion_lib.npz"))
    arrays = synthetic_motion_lib(int(g["lib_motions"]), int(g["lib_seed"]))
    lib = MotionLibSMPL(arrays)
    ids, times = torch.from_numpy(g["ids"]).cuda(), torch.from_numpy(g["times"]).cuda()
    ms = lib.get_motion_state_smpl(ids, times, full=True)
    root, dof, rb = ms["root_state"].cpu().numpy(), ms["dof_state"].cpu().numpy(), ms["rb_state"].cpu().numpy()
    tol = dict(rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(root[:, 0:3], g["ms_root_pos"], **tol); np.testing.assert_allclose(root[:, 3:7], g["ms_root_rot"], **tol)
    np.testing.assert_allclose(root[:, 7:10], g["ms_root_vel"], **tol); np.testing.assert_allclose(root[:, 10:13], g["ms_root_ang_vel"], **tol)
    np.testing.assert_allclose(dof[..., 0], g["ms_dof_pos"], rtol=1e-4, atol=5e-5); np.testing.assert_allclose(dof[..., 1], g["ms_dof_vel"], **tol)
    np.testing.assert_allclose(ms["key_pos"].cpu().numpy(), g["ms_key_pos"], **tol)
    np.testing.assert_allclose(rb[..., 0:3], g["ms_rg_pos"], **tol); np.testing.assert_allclose(rb[..., 3:7], g["ms_rb_rot"], **tol)
    np.testing.assert_allclose(rb[..., 7:10], g["ms_body_vel"], **tol); np.testing.assert_allclose(rb[..., 10:13], g["ms_body_ang_vel"], **tol)
    demo = lib.fetch_amp_obs_demo(len(g["ids"]), motion_ids=ids, motion_times0=times)
    np.testing.assert_allclose(demo.cpu().numpy(), g["demo"], rtol=1e-3, atol=5e-5)
    # larger, fresh draws against the oracle
    big = MotionLibSMPL(synthetic_motion_lib(32, 5), seed=3)
    i2 = big.sample_motions(5000); t2 = big.sample_time(i2)
    assert i2.min() >= 0 and i2.max() < 32 and (t2 >= 0).all() and (t2 <= big.motion_lengths[i2.long()]).all()
    d2 = big.fetch_amp_obs_demo(5000, motion_ids=i2, motion_times0=t2).cpu().numpy()
    ref = O.amp_obs_demo(synthetic_motion_lib(32, 5), i2.cpu().numpy().astype(np.int64), t2.cpu().numpy())
    # The reference's slerp is not normalised: for two almost identical frames of an almost-identity joint rotation the factor
    # sin((1-t) h) / sin h + sin(t h) / sin h, with h = acos(c) and sin h = sqrt(1 - c^2) from a c one or two ulps below 1, is
    # off by a few per cent in ANY fp32 implementation; when it makes w >= 1 the reference's quat_to_exp_map takes its NaN /
    # default-axis branch (exactly identity), otherwise a rotation of a few 1e-3 rad comes out (the true one is ~1e-4).  Which
    # side a sample lands on depends on the last bit of acosf: a handful of joint entries per million may differ by < 0.02.
    bad = np.abs(d2 - ref) > 1e-4 + 1e-3 * np.abs(ref)
    assert bad.mean() < 2e-5 and np.abs(d2 - ref).max() < 0.03, (bad.sum(), np.abs(d2 - ref).max())
    assert big.fetch_amp_obs_demo(0).shape == (0, 3090)
    # reset from mocap states: rows written in place, xy kept, unit root quaternions
    from emloco_b200.rollout import Rollout
    R = Rollout(64, seed=1, tensor_cores=False)
    xy = R.init_root[:, :2].clone()
    big.sample_reset_state(R.init_root, R.init_dof)
    torch.cuda.synchronize()
    assert torch.equal(R.init_root[:, :2], xy)
    np.testing.assert_allclose(R.init_root[:, 3:7].norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    assert float(R.init_dof.abs().max()) > 0.01
    R.sim.reset.fill_(1)
    R.step(0)                                        # envs restart from the sampled mocap states and step without NaNs
    torch.cuda.synchronize()
    assert torch.isfinite(R.sim.obs).all() and torch.isfinite(R.sim.rb_state).all()
    R.close()
