"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by running the REFERENCE's own
functions (loaded from /root/reference by oracle/ref_extract.py) on seeded synthetic inputs.

Run once in the build container:  python -m oracle.make_golden
The fixtures are committed; the GPU box never needs the reference tree.
Input recipe follows SURVEY.md section 8c ("Golden vectors"): unit quats, |vel|<=3,
dof in +-1 rad, progress in [0,167], contact forces N(0,30).
"""
from __future__ import annotations

import os

import numpy as np

from . import oracle_np as O
from . import ref_extract
from .oracle_np import (AMP_STEP_DIM, AMP_STEPS, CONTACT_BODIES, CONTROL_DT, DOF_SUBSET, EPISODE_LEN, HEAD,
                        KEY_BODIES, NB, ND, NUM_TRAJ_SAMPLES, NUM_VERTS, TRAJ_SAMPLE_DT, center_height_points,
                        square_height_points, traj_dt)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_state(N, seed, map_shape=(1080, 1080), rough=True):
    """Seeded random-but-valid env state; also used (same code) by the GPU parity tests."""
    rng = np.random.default_rng(seed)
    f = np.float32
    root_xy = rng.uniform(50.0, 58.0, (N, 2))
    body_pos = rng.normal(0, 0.4, (N, NB, 3))
    body_pos[:, 0] = 0
    body_pos[..., 2] += 0.9
    body_pos[..., :2] += root_xy[:, None]
    q = rng.normal(0, 1, (N, NB, 4)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    # roots mostly upright with random heading so that the heading frame is well conditioned
    yaw = rng.uniform(-np.pi, np.pi, N)
    tilt = rng.normal(0, 0.15, (N, 3))
    qr = np.stack([tilt[:, 0], tilt[:, 1], np.sin(yaw / 2) + 0 * tilt[:, 2], np.cos(yaw / 2)], -1)
    qr /= np.linalg.norm(qr, axis=-1, keepdims=True)
    q[:, 0] = qr
    body_vel = rng.uniform(-3, 3, (N, NB, 3))
    body_ang = rng.uniform(-3, 3, (N, NB, 3))
    rb = np.concatenate([body_pos, q, body_vel, body_ang], -1).astype(f)
    dof_pos = rng.uniform(-1, 1, (N, ND))
    dof_pos[: max(1, N // 8), :6] = 0.0          # exercise the |angle|<1e-5 default-axis branch
    dof_vel = rng.uniform(-3, 3, (N, ND))
    dof_state = np.stack([dof_pos, dof_vel], -1).astype(f)
    contact = rng.normal(0, 30, (N, NB, 3)).astype(f)
    dof_force = rng.normal(0, 50, (N, ND)).astype(f)
    progress = rng.integers(0, EPISODE_LEN, N).astype(np.int64)
    progress[:4] = [0, 1, 166, 167][: min(4, N)]
    betas = np.concatenate([rng.integers(0, 3, (N, 1)), rng.normal(0, 1, (N, 16))], -1).astype(f)
    # JTA-shaped polyline: smooth walk starting near the root
    speed = rng.uniform(0, 3, (N, 1)); head = rng.uniform(-np.pi, np.pi, (N, 1))
    turn = rng.uniform(-0.5, 0.5, (N, 1))
    t = np.arange(NUM_VERTS)[None] * traj_dt()
    ang = head + turn * t
    vx, vy = speed * np.cos(ang), speed * np.sin(ang)
    verts = np.zeros((N, NUM_VERTS, 3))
    verts[..., 0] = root_xy[:, :1] + np.cumsum(vx, 1) * traj_dt() + rng.normal(0, 0.3, (N, 1))
    verts[..., 1] = root_xy[:, 1:] + np.cumsum(vy, 1) * traj_dt() + rng.normal(0, 0.3, (N, 1))
    verts = verts.astype(f)
    if rough:
        # blocky random steps (4x4 cells) - low entropy so the committed fixture stays small
        blk = rng.integers(-40, 40, ((map_shape[0] + 3) // 4, (map_shape[1] + 3) // 4))
        hs = np.kron(blk, np.ones((4, 4), np.int64))[:map_shape[0], :map_shape[1]].astype(np.int16)
    else:
        hs = np.zeros(map_shape, np.int16)
    amp_buf = rng.normal(0, 1, (N, AMP_STEPS, AMP_STEP_DIM)).astype(f)
    return dict(rb=rb, dof_state=dof_state, contact=contact, dof_force=dof_force, progress=progress,
                betas=betas, verts=verts, height_samples=hs, amp_buf=amp_buf)


def post_step_inputs(st, device="cpu"):
    """Tensors of one env state on `device` (the sim tensors and constants the task code holds)."""
    R = ref_extract.load()
    torch = R.torch
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    N = st["rb"].shape[0]
    I = dict(N=N, rb=T(st["rb"]), ds=T(st["dof_state"]), betas=T(st["betas"]), limb=torch.zeros(N, 10, device=device), progress=T(st["progress"]),
             verts=T(st["verts"]), heightsamples=T(st["height_samples"]), grid=T(square_height_points())[None].repeat(N, 1, 1),
             grid9=T(center_height_points())[None].repeat(N, 1, 1), dof_force=T(st["dof_force"]), contact=T(st["contact"]),
             amp_buf=T(st["amp_buf"]), env_ids=torch.arange(N, dtype=torch.long, device=device),
             ts=torch.arange(NUM_TRAJ_SAMPLES, dtype=torch.float, device=device) * TRAJ_SAMPLE_DT,
             contact_bodies=torch.tensor(CONTACT_BODIES, device=device), dof_subset=torch.from_numpy(DOF_SUBSET).to(device),
             term_heights=torch.zeros(NB, device=device), reset0=torch.zeros(N, dtype=torch.long, device=device))
    return I


def post_step_run(I):
    """Drives the reference functions exactly as the task code does (call sites cited inline); torch tensors in and out."""
    R = ref_extract.load()
    torch = R.torch
    N = I["N"]
    rb = I["rb"]
    bp, br, bv, bw = rb[..., 0:3], rb[..., 3:7], rb[..., 7:10], rb[..., 10:13]
    ds = I["ds"]; dpos, dvel = ds[..., 0].contiguous(), ds[..., 1].contiguous()
    betas, limb, progress = I["betas"], I["limb"], I["progress"]
    # --- self obs: humanoid_pedestrain_terrain.py:240-268 (root_height_obs False)
    so = R.jit.compute_humanoid_observations_smpl_max(bp.clone(), br, bv, bw, betas, limb, True, False, True, True, False)
    # --- flip self obs: humanoid.py:1066-1108
    fp, fr, fv, fw = bp.clone(), br.clone(), bv.clone(), bw.clone()
    l2r = R.left_to_right_index
    fp[..., 1] *= -1; fp = fp[..., l2r, :]
    fr[..., 0] *= -1; fr[..., 2] *= -1; fr = fr[..., l2r, :]
    fv[..., 1] *= -1; fv = fv[..., l2r, :]
    fw[..., 0] *= -1; fw[..., 2] *= -1; fw = fw[..., l2r, :]
    fso = R.jit.compute_humanoid_observations_smpl_max(fp, fr, fv, fw, betas, limb, True, False, True, True, False)
    # --- traj generator holder (traj_generator.py:24-36, 256-272)
    tg = R.meth.TrajHolder()
    tg._verts = I["verts"]; tg._verts_flat = tg._verts.view(-1, 3)
    tg._dt = (EPISODE_LEN * CONTROL_DT) / (NUM_VERTS - 1)
    tg.get_num_verts = lambda: tg._verts.shape[1]
    tg.get_num_segs = lambda: tg._verts.shape[1] - 1
    tg.get_traj_duration = lambda: tg.get_num_verts() * tg._dt
    env_ids = I["env_ids"]
    # --- _fetch_traj_samples: humanoid_traj.py:208-224
    t0 = progress * CONTROL_DT
    tt = t0.unsqueeze(-1) + I["ts"]
    ids = torch.broadcast_to(env_ids.unsqueeze(-1), tt.shape)
    samples = tg.calc_pos(ids.flatten(), tt.flatten()).reshape(N, NUM_TRAJ_SAMPLES, 3)
    root_states = torch.cat([bp[:, 0], br[:, 0], bv[:, 0], bw[:, 0]], -1)
    loc = R.jit.compute_location_observations(root_states, samples, True)
    # --- heights: humanoid_pedestrain_terrain.py:761-815, 732-759, 1212-1288
    ter = R.meth.TerrainHolder()
    ter.heightsamples = I["heightsamples"]; ter.horizontal_scale = 0.1; ter.vertical_scale = 0.005
    grid, grid9 = I["grid"], I["grid9"]
    head = torch.cat([bp[:, HEAD], br[:, HEAD]], 1)
    hq = R.ptu.calc_heading_quat(head[:, 3:7])
    pts = R.itu.quat_apply(hq.repeat(1, grid.shape[1]).reshape(-1, 4), grid) + head[:, :3].unsqueeze(1)
    meas = ter.sample_height_points(pts.clone()).view(N, -1)
    base_quat = root_states[:, 3:7]
    pts9 = R.jit.quat_apply_yaw(base_quat.repeat(1, 9), grid9) + root_states[:, :3].unsqueeze(1)
    ch = ter.sample_height_points(pts9.clone()).view(N, -1).mean(dim=-1, keepdim=True)
    heights = torch.clip(ch - meas, -3, 3.) * 5
    tobs = torch.cat([loc, heights], dim=1)
    # --- flip task obs: humanoid_pedestrain_terrain.py:455-491
    nt = tobs.clone()
    tr = nt[:, :30].view(N, 15, 2); tr[..., 1] *= -1
    hsamp = nt[..., 30:30 + 1024].view(N, 32, 32).flip(2)
    ftobs = torch.cat([tr.view(N, -1), hsamp.reshape(N, -1)], dim=1)
    # --- reward: humanoid_pedestrain_terrain.py:907-930
    tar = tg.calc_pos(env_ids, progress * CONTROL_DT)
    loc_r = 1 * R.jit.compute_location_reward(bp[:, 0], tar)
    power = torch.abs(torch.multiply(I["dof_force"], dvel)).sum(dim=-1)
    pow_r = -0.0005 * power
    rew = loc_r + pow_r
    rew_raw = torch.cat([loc_r[:, None], pow_r[:, None]], dim=-1)
    # --- reset: humanoid_pedestrain_terrain.py:883-905
    reset, term = R.jit.compute_humanoid_reset(
        I["reset0"], progress, I["contact"], I["contact_bodies"],
        ch, bp, tar, float(EPISODE_LEN), 4.0, True, I["term_heights"], False)
    # --- AMP obs: humanoid_amp.py:585-657
    key = bp[:, KEY_BODIES, :]
    cur = R.jit.build_amp_observations_smpl(bp[:, 0], br[:, 0], bv[:, 0], bw[:, 0], dpos, dvel, key, betas, limb,
                                            I["dof_subset"], True, False, True, True, False, True)
    amp = I["amp_buf"].clone()
    amp[:, 1:] = amp[:, 0:AMP_STEPS - 1].clone()
    amp[:, 0] = cur
    return dict(obs=torch.cat([so, tobs], -1), flip_obs=torch.cat([fso, ftobs], -1), rew=rew, reward_raw=rew_raw,
                reset=reset, terminate=term, amp_obs=amp.view(N, -1), tar_pos=tar, traj_samples=samples)


def reference_post_step(st):
    return {k: v.cpu().numpy() for k, v in post_step_run(post_step_inputs(st)).items()}


def synth_locoval(B, seed):
    rng = np.random.default_rng(seed)
    f = np.float32
    speed = rng.uniform(0, 3, (B, 1)); head = rng.uniform(-np.pi, np.pi, (B, 1)); turn = rng.uniform(-0.5, 0.5, (B, 1))
    t = np.arange(13)[None] * 0.4
    traj = np.zeros((B, 13, 2))
    traj[..., 0] = np.cumsum(speed * np.cos(head + turn * t) * 0.4, 1)
    traj[..., 1] = np.cumsum(speed * np.sin(head + turn * t) * 0.4, 1)
    traj -= traj[:, :1]
    traj[0, 1, 0] = 0.0                      # exercises the |x|<1e-10 epsilon branch
    pose = rng.normal(0, 0.3, (B, 24, 3))
    vel = (traj[:, 1] - traj[:, 0]) * 2.5
    return traj.astype(f), pose.astype(f), vel.astype(f)


def reference_locoval(traj, pose, vel, seed=0):
    R = ref_extract.load()
    torch = R.torch
    torch.manual_seed(seed)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = R.ValuePoseNet(use_pose=True, use_vel=True)
    with torch.no_grad():   # non-zero biases so the test is not blind to them
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.1, 0.1)
    W = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    t = torch.from_numpy(traj.copy()).requires_grad_(True)
    p = torch.from_numpy(pose.copy())
    v = torch.from_numpy(vel.copy())
    value, loss = net.calc_embodied_motion_loss(t, p, v)
    loss.backward()
    return W, dict(value=value.detach().numpy(), loss=loss.detach().numpy(), grad_traj=t.grad.numpy(),
                   pose_after=p.detach().numpy())


def reference_locoval_multimodal(B=256, modes=5, seed=5):
    """BASELINE configs[4]: the multi-modal EmLoco loss of social-transmotion/train_jta.py:289-299 - one
    calc_embodied_motion_loss call per mode on the non-contiguous slice pred_trajs[:, :, i], the SAME init_pose tensor
    passed every time (so mode i sees the pose already rotated / zeroed by modes 0..i-1, value_pose_net.py:97,141-144),
    losses summed, scaled and averaged, gradient taken w.r.t. the predictor output."""
    R = ref_extract.load()
    torch = R.torch
    torch.manual_seed(seed)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = R.ValuePoseNet(use_pose=True, use_vel=True)
    with torch.no_grad():
        for m in net._network:
            if hasattr(m, "bias"):
                m.bias.uniform_(-0.1, 0.1)
    W = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    rng = np.random.default_rng(seed)
    f = np.float32
    step = rng.normal(0.35, 0.25, (B, 12, modes, 2)).astype(f) * rng.choice([-1, 1], (B, 1, modes, 2)).astype(f)
    pred = np.cumsum(step, 1).astype(f)                                   # [B,12,modes,2] predicted future positions
    _, pose, vel = synth_locoval(B, seed + 1)
    p = torch.from_numpy(pose.copy()); v = torch.from_numpy(vel.copy())
    out = torch.from_numpy(pred.copy()).requires_grad_(True)              # "pred_joints[:, in_F:]"
    trajs = torch.cat([torch.zeros(B, 1, modes, 2), out], dim=1)          # :291
    weight = 0.5                                                          # config TRAIN.valuenet_weight
    losses, values = 0, []
    for i in range(modes):                                                # :294-296
        val, l = net.calc_embodied_motion_loss(trajs[:, :, i], p, v)
        losses = losses + l
        values.append(val.detach().numpy().copy())
    losses = losses * weight / modes                                      # :297-298
    losses.backward()
    return dict(pred=pred, pose=pose, vel=vel, weight=f(weight), **{f"w_{k}": a for k, a in W.items()},
                out_values=np.stack(values, 0), out_loss=losses.detach().numpy(), out_grad_pred=out.grad.numpy(),
                out_pose_after=p.detach().numpy())


def reference_locoval_finetune(N=96, seed=9):
    """The `_do_finetune` block of play_steps (amp_continuous_value.py:122-146) driven on the reference ValuePoseNet with the
    optimiser and criterion of common_agent.py:94-96; four rounds, the third with no valid env (no optimiser step)."""
    R = ref_extract.load()
    torch = R.torch
    torch.manual_seed(seed)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = R.ValuePoseNet(use_pose=True, use_vel=True)
    net.train()
    W0 = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.0001)              # common_agent.py:94
    crit = torch.nn.MSELoss(reduction='sum')                                             # :96
    opt.zero_grad()
    rng = np.random.default_rng(seed)
    f = np.float32
    rmin, rmax = -10.0, 100.0                                                            # :154-155
    out = dict(N=N, r_min=f(rmin), r_max=f(rmax), **{f"w0_{k}": a for k, a in W0.items()})
    for r, frac in enumerate((0.3, 1.0, 0.0, 0.05)):
        traj2, pose, vel = synth_locoval(N, seed + 10 + r)
        traj = np.concatenate([traj2, rng.normal(0, 1, (N, 13, 1)).astype(f)], -1)        # waypoint_traj[:, :13, :] is [N,13,3]
        gc = (rng.uniform(5, 80, N) * (rng.random(N) < frac)).astype(f)
        t, p, v = torch.from_numpy(traj.copy()), torch.from_numpy(pose.copy()), torch.from_numpy(vel.copy())
        game = torch.from_numpy(gc.copy())
        valid = torch.nonzero(game, as_tuple=True)                                       # :123
        loss_v = 0.0
        if len(valid[0]) > 0:                                                            # :124
            pred = net(t[:, :13, :], p, v).squeeze()                                     # :129
            norm = (game[valid] - rmin) / (rmax - rmin)                                  # :135
            loss = crit(pred[valid], norm)                                               # :137
            loss.backward(); opt.step(); opt.zero_grad()                                 # :138-140
            loss_v = loss.item()
            out[f"r{r}_pred_sum"] = f(pred[valid].sum().item()); out[f"r{r}_gt_sum"] = f(norm.sum().item())
        out.update({f"r{r}_traj": traj, f"r{r}_pose": pose, f"r{r}_vel": vel, f"r{r}_gc": gc, f"r{r}_loss": f(loss_v),
                    f"r{r}_count": len(valid[0])})
        out.update({f"r{r}_w_{k}": a.detach().numpy().copy() for k, a in net.state_dict().items()})
    return out


def reference_gae(T_, N, seed):
    R = ref_extract.load()
    torch = R.torch
    rng = np.random.default_rng(seed)
    f = np.float32
    dones = (rng.uniform(0, 1, (T_, N)) < 0.05).astype(f)
    values = rng.normal(0, 1, (T_, N, 1)).astype(f)
    nvalues = rng.normal(0, 1, (T_, N, 1)).astype(f)
    rewards = rng.uniform(0, 1, (T_, N, 1)).astype(f)
    ag = R.meth.AgentHolder(); ag.horizon_length = T_; ag.gamma = 0.99; ag.tau = 0.95
    adv = ag.discount_values(*[torch.from_numpy(a) for a in (dones, values, rewards, nvalues)])
    return dict(dones=dones, values=values, rewards=rewards, next_values=nvalues, adv=adv.numpy())


def reference_rms(x, mean, var):
    R = ref_extract.load()
    torch = R.torch
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RunningMeanStd(x.shape[-1])
    m.running_mean[:] = torch.from_numpy(mean); m.running_var[:] = torch.from_numpy(var)
    m.eval()
    return m(torch.from_numpy(x)).numpy()


def reference_traj_reset(N, n_reset, seed, flags, pool_size=7):
    """TrajGenerator.reset + the _reset_task outputs driven with recorded uniform draws (torch.rand / torch.bernoulli /
    random.sample of the hosted module are replaced by replay proxies - the arithmetic is the reference's)."""
    import types
    R = ref_extract.load()
    torch = R.torch
    mod = R.trajreset
    rng = np.random.default_rng(seed)
    f = np.float32
    U = rng.random((n_reset, O.TRAJ_RAND_COLS)).astype(f)
    U[:, 200:300] = np.where(rng.random((n_reset, 100)) < 0.1, 0.01, U[:, 200:300])     # make sharp turns common enough to matter
    env_ids = np.sort(rng.choice(N, n_reset, replace=False))
    verts0 = rng.normal(0, 1, (N, NUM_VERTS, 3)).astype(f)
    init_pos = np.concatenate([50 + 8 * rng.random((n_reset, 2)), np.full((n_reset, 1), 0.9)], 1).astype(f)
    root_vel = rng.normal(0, 1, (n_reset, 3)).astype(f)
    root_vel[0] = 0                                                                      # the zero-velocity branch (:184)
    pool = np.cumsum(rng.normal(0, 0.03, (pool_size, NUM_VERTS, 3)), 1).astype(f)
    pool[..., 2] = 0
    pool[1, 1] = pool[1, 0]                                                              # zero first segment (:185, :150)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    S = NUM_VERTS - 1
    queue = [T(U[:, 0:S]), T(U[:, 100:100 + S]), None, T(U[:, 300]), T(U[:, 301:301 + S]), T(U[:, 401]), T(U[:, 402]), T(U[:, 404])]
    order = iter([0, 1, 3, 4, 5, 6, 7])

    class TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        def rand(self, *a, **kw):
            t = queue[next(order)].clone()
            shape = list(a[0]) if a and isinstance(a[0], (list, tuple)) else list(a)
            assert list(t.shape) == shape, (t.shape, shape)
            return t

        def bernoulli(self, p):
            return (T(U[:, 200:200 + S]) < p).float()

    picks = np.minimum((U[:, 403] * f(pool_size)).astype(np.int64), pool_size - 1)

    class RandomProxy:
        def sample(self, population, k):
            real = U[:, 402] > f(0.5)
            assert k == int(real.sum())
            return [int(i) for i in picks[real]]

    mod.torch, mod.random = TorchProxy(), RandomProxy()
    h = mod.TrajResetHolder()
    h._device = "cpu"
    h._dt = (EPISODE_LEN * CONTROL_DT) / (NUM_VERTS - 1)
    h._dtheta_max, h._speed_min, h._speed_max, h._accel_max, h._sharp_turn_prob = 2.0, 0.0005, 3.0, 2.0, 0.02
    h._hybrid_init_prob = 0.5
    h._verts = T(verts0.copy()); h._verts_flat = h._verts.view(-1, 3)
    h.inverted = torch.zeros(N, dtype=torch.bool)
    data = [{"traj": pool[i].copy()} for i in range(pool_size)]
    h.traj_data = [data]; h.traj_data_jta = data
    F_ = lambda b: bool(flags & b)
    h._flags = types.SimpleNamespace(fixed_path=False, slow=F_(O.TRAJ_F_SLOW), adjust_root_vel=F_(O.TRAJ_F_ADJUST_VEL),
                                     real_path=F_(O.TRAJ_F_REAL), jta_path=True, jrdb_path=False, pred_path=False,
                                     init_heading=F_(O.TRAJ_F_INIT_HEADING), heading_inversion=F_(O.TRAJ_F_INVERSION),
                                     add_noise=False)
    import contextlib, io, warnings
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()) as so:
        warnings.simplefilter("ignore")
        h.reset(T(env_ids), T(init_pos), T(root_vel))
    assert "alignment failed" not in so.getvalue(), so.getvalue()
    mod.torch, mod.random = torch, __import__("random")
    return dict(U=U, env_ids=env_ids, verts0=verts0, init_pos=init_pos, root_vel=root_vel, pool=pool, flags=flags,
                verts=h._verts.numpy().copy(), inverted=h.inverted.numpy()[env_ids].copy())


def synth_state_branches(N, seed):
    """synth_state on flat terrain with every branch of compute_humanoid_reset populated: contact sums on both sides of the 50 N
    threshold, progress in the exempt range (<= 1), at the episode limit (>= 167) and in between, and trajectories nearer / farther
    than the 4 m fail distance - so that alive, time-out-only, fallen and target-lost envs all occur (asserted by the generator)."""
    st = synth_state(N, seed, map_shape=(700, 700), rough=False)
    rng = np.random.default_rng(seed + 1000)
    masked = st["contact"].copy(); masked[:, CONTACT_BODIES] = 0
    norm = np.linalg.norm(masked.sum(1), axis=-1)
    want = rng.uniform(0, 100, N)
    want[: N // 8] = rng.uniform(49.5, 50.5, N // 8)                 # close to the threshold
    st["contact"] = (st["contact"] * (want / norm)[:, None, None]).astype(np.float32)
    prog = rng.integers(2, 166, N)
    k = N // 8
    prog[0:k] = rng.integers(0, 2, k)                                 # exempt from the contact test
    prog[k:2 * k] = rng.integers(166, 169, k)                         # 166 alive, 167 / 168 timed out
    st["progress"] = prog.astype(np.int64)
    # bodies moved next to the trajectory point of the env's progress (a walker that follows its path), a few left behind
    tar = O.calc_pos(st["verts"], np.arange(N), O.progress_time(st["progress"]))
    shift = tar[:, :2] - st["rb"][:, 0, :2] + rng.normal(0, 1.0, (N, 2))
    far = rng.random(N) < 0.15
    shift[far] += rng.uniform(3.5, 6.0, (int(far.sum()), 1)) * np.array([[1.0, 0.0]])
    st["rb"][:, :, 0:2] += shift[:, None, :].astype(np.float32)
    return st


def reference_nets(M=24, seed=21):
    """a11-a13: the reference's own AMPSeptValueBuilder.Network (ref_extract.load_network) with the synthetic parameters of
    oracle/netweights.py, driven as the agent drives it: running_mean_std (eval) -> eval_actor / eval_critic / eval_task_value
    (amp_sept_value_models.py:22-30 -> amp_models.py:20-44), _eval_critic with value un-normalisation (common_agent.py:647-655),
    _eval_disc / _calc_disc_rewards / _combine_rewards (amp_continuous.py:659-692)."""
    from . import netweights
    R = ref_extract.load()
    torch = R.torch
    net = ref_extract.load_network()
    sd = netweights.synth_state_dict(seed)
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    rng = np.random.default_rng(seed)
    f = np.float32
    obs = rng.normal(0, 2.0, (M, 1422)).astype(f); obs[:, ::7] *= 4                    # some features beyond the +-5 clamp
    amp = rng.normal(0, 2.0, (M, 3090)).astype(f); amp[:, ::11] *= 4
    obs_mean, obs_var = rng.normal(0, 1, 1422), rng.uniform(0.05, 4, 1422)
    amp_mean, amp_var = rng.normal(0, 1, 3090), rng.uniform(0.05, 4, 3090)
    v_mean, v_var = np.array([0.7]), np.array([2.3])
    task_r = rng.uniform(0, 1, (M, 1)).astype(f)
    import contextlib, io
    H = ref_extract.load_agent_blocks()
    h = H()
    with contextlib.redirect_stdout(io.StringIO()):
        h.running_mean_std, h._amp_input_mean_std, h.value_mean_std = R.RunningMeanStd((1422,)), R.RunningMeanStd((3090,)), R.RunningMeanStd((1,))
    for m, (mu_, var_) in ((h.running_mean_std, (obs_mean, obs_var)), (h._amp_input_mean_std, (amp_mean, amp_var)),
                           (h.value_mean_std, (v_mean, v_var))):
        m.running_mean[:] = torch.from_numpy(mu_); m.running_var[:] = torch.from_numpy(var_); m.eval()
    import types
    h.model = types.SimpleNamespace(a2c_network=net, eval=lambda: None)
    h.normalize_input = h.normalize_value = h._normalize_amp_input = True
    h._disc_reward_mean_std = None; h.ppo_device = "cpu"; h._disc_reward_scale = 2.0
    h._task_reward_w = h._disc_reward_w = 0.5
    with torch.no_grad():
        x = h._preproc_obs(torch.from_numpy(obs))
        mu, sigma = net.eval_actor(x)
        value = net.eval_critic(x)
        tv = net.eval_task_value(x)
        next_value = h._eval_critic({"obs": torch.from_numpy(obs)})
        logit = h._eval_disc(torch.from_numpy(amp))
        amp_r = h._calc_amp_rewards(torch.from_numpy(amp))
        comb = h._combine_rewards(torch.from_numpy(task_r), amp_r)
    return dict(seed=seed, weights_checksum=netweights.checksum(sd), obs=obs, amp_obs=amp, obs_mean=obs_mean, obs_var=obs_var,
                amp_mean=amp_mean, amp_var=amp_var, value_mean=v_mean, value_var=v_var, task_rewards=task_r,
                out_mu=mu.numpy(), out_sigma=sigma.numpy(), out_value=value.numpy(), out_task_value=tv.numpy(),
                out_next_value_unnorm=next_value.numpy(), out_disc_logit=logit.numpy(), out_disc_reward=amp_r["disc_rewards"].numpy(),
                out_combined=comb.numpy())


def reference_play_block(N=96, steps=8, seed=31):
    """a14 (+ the reward / value plumbing around it): the body of the horizon loop of AMPValueAgent.play_steps from env_step to the
    end of the no_grad block (amp_continuous_value.py:61-121, incl. the inversion penalty :62-64 and next_vals *= 1-terminated
    :87-89) run `steps` times on a holder whose env, critic and discriminator return recorded tensors (the networks themselves
    are pinned by nets.npz).  State after every step and the rows written to the experience buffer are the fixture."""
    import types
    R = ref_extract.load()
    torch = R.torch
    rng = np.random.default_rng(seed)
    f = np.float32
    H = ref_extract.load_agent_blocks()
    h = H()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        h.value_mean_std = R.RunningMeanStd((1,))
    v_mean, v_var = 0.4, 1.7
    h.value_mean_std.running_mean[:] = v_mean; h.value_mean_std.running_var[:] = v_var; h.value_mean_std.eval()
    inverted = rng.random(N) < 0.3
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    rec = {}
    cur = {}
    h.vec_env = types.SimpleNamespace(env=types.SimpleNamespace(task=types.SimpleNamespace(inverted=T(inverted), viewer=None)))
    h.env_step = lambda actions: (dict(obs=cur["obs"]), cur["rew"].clone(), cur["dones"], cur["infos"])
    h.inversion_penalty_scale = 0.3
    h.rewards_shaper = lambda r: r * 1.0                                # rl_games DefaultRewardsShaper(scale_value=1)
    h.experience_buffer = types.SimpleNamespace(update_data=lambda k, n, v: rec.setdefault(k, []).append(v.clone().numpy()))
    h.motion_sym_loss = True
    h.normalize_input = False; h.normalize_value = True; h._normalize_amp_input = False
    h.running_mean_std = None
    h._preproc_obs = lambda o: o
    h.model = types.SimpleNamespace(eval=lambda: None, a2c_network=types.SimpleNamespace(
        eval_critic=lambda o: cur["critic_raw"].clone(), eval_disc=lambda a: cur["logit"].clone()))
    h._disc_reward_mean_std = None; h.ppo_device = "cpu"; h._disc_reward_scale = 2.0
    h.num_agents = 1; h.gamma = 0.99; h.step_to_pred = 144
    h.game_rewards = h.game_lengths = types.SimpleNamespace(update=lambda x: None)
    h.algo_observer = types.SimpleNamespace(process_infos=lambda i, d: None)
    # state as CommonAgent.init_tensors leaves it (common_agent.py:137-141), lengths moved close to step_to_pred
    len0 = rng.integers(0, 150, N).astype(f); len0[:12] = 143; len0[12:20] = 144; len0[20:26] = 142
    h.current_rewards = T(rng.normal(0, 5, (N, 1)).astype(f)); h.current_lengths = T(len0.copy())
    h.current_combined_rewards = T(rng.normal(0, 5, N).astype(f)); h.game_combined_rewards = torch.zeros(N)
    h.discount_coefs = T((0.99 ** len0).astype(f))
    state0 = dict(current_rewards=h.current_rewards.numpy().copy(), current_lengths=len0.copy(),
                  current_combined=h.current_combined_rewards.numpy().copy(), discount=h.discount_coefs.numpy().copy())
    ins = dict(rew=[], reset=[], terminate=[], critic_raw=[], logit=[])
    outs = dict(game_combined=[], current_combined=[], discount=[], current_lengths=[], current_rewards=[], terminated_flags=[])
    tf, rr = torch.zeros(N), torch.zeros(1)
    for n in range(steps):
        term = rng.random(N) < 0.08
        reset = term | (rng.random(N) < 0.08)
        if n == 1:
            reset[:16] = [True, False] * 8; term[:16] = False
        cur.update(obs=torch.zeros(N, 1), rew=T(rng.uniform(-0.2, 1.0, (N, 1)).astype(f)), dones=T(reset.astype(np.int64)),
                   critic_raw=T(rng.normal(0, 3, (N, 1)).astype(f)), logit=T(rng.normal(0, 4, (N, 1)).astype(f)))
        cur["logit"][:3, 0] = torch.tensor([12.0, -12.0, 9.3])                      # the 1e-4 floor of the log (:683-685)
        cur["infos"] = dict(amp_obs=torch.zeros(N, 1), flip_obs=torch.zeros(N, 1), terminate=T(term.astype(np.int64)),
                            reward_raw=torch.zeros(N, 2))
        for k in ins:
            src = dict(rew=cur["rew"], reset=cur["dones"], terminate=cur["infos"]["terminate"], critic_raw=cur["critic_raw"], logit=cur["logit"])[k]
            ins[k].append(src.numpy().copy())
        with torch.no_grad():
            loc = h.play_block(n, dict(actions=None), tf, rr)
        outs["game_combined"].append(h.game_combined_rewards.numpy().copy()); outs["current_combined"].append(h.current_combined_rewards.numpy().copy())
        outs["discount"].append(h.discount_coefs.numpy().copy()); outs["current_lengths"].append(h.current_lengths.numpy().copy())
        outs["current_rewards"].append(h.current_rewards.numpy().copy()); outs["terminated_flags"].append(tf.numpy().copy())
        rec.setdefault("amp_rewards", []).append(loc["amp_rewards"].numpy().copy())
        if n == 4:                                                                  # what the finetune block does after consuming them (:145)
            h.game_combined_rewards = torch.zeros(N)
    out = dict(N=N, steps=steps, inverted=inverted, value_mean=f(v_mean), value_var=f(v_var), zero_game_after_step=4,
               **{f"state0_{k}": v for k, v in state0.items()}, **{f"in_{k}": np.stack(v) for k, v in ins.items()},
               **{f"out_{k}": np.stack(v) for k, v in outs.items()})
    for k in ("rewards", "dones", "next_values", "amp_rewards"):
        out[f"row_{k}"] = np.stack(rec[k])
    return out

UPDATE_MU_GAIN = 3.0
UPDATE_CFG = dict(e_clip=0.2, actor_coef=1.0, critic_coef=5.0, tv_coef=5.0, bounds_loss_coef=10.0, entropy_coef=0.0, disc_coef=5.0,
                  disc_logit_reg=0.01, disc_grad_penalty=5.0, disc_weight_decay=0.0001, dropout_rate=0.3, lr=2e-5, grad_norm=50.0)
# the values of data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml:84-122 (checked against the file by reference_update_step)


def synth_update_batch(B, Ba, seed):
    """Seeded minibatch of the PPO / AMP update (the `input_dict` of calc_gradients): also used, same code, by the GPU tests."""
    rng = np.random.default_rng(seed)
    f = np.float32
    obs = rng.normal(0, 2.0, (B, 1422)).astype(f); obs[:, ::7] *= 4
    # `actions` / `old_logp_actions` here are placeholders: reference_update_step replaces them by samples around the
    # network's own mu (so that PPO ratios straddle the clip range) and stores those in the fixture
    d = dict(obs=obs, actions=rng.normal(0, 0.4, (B, 69)).astype(f), old_logp_actions=rng.normal(-140, 3, B).astype(f),
             advantages=rng.normal(0, 1, B).astype(f), returns=rng.normal(0, 1, (B, 1)).astype(f), old_values=rng.normal(0, 1, (B, 1)).astype(f),
             mu=rng.normal(0, 0.3, (B, 69)).astype(f), sigma=np.full((B, 69), np.exp(-2.9), f))
    for k in ("amp_obs", "amp_obs_replay", "amp_obs_demo"):
        a = rng.normal(0, 2.0, (Ba, 3090)).astype(f); a[:, ::11] *= 4
        d[k] = a
    d["dropout_u"] = rng.random((19, Ba, 3)).astype(f)               # get_dropout_mask: one torch.rand(B, 3) per joint
    stats = dict(obs_mean=rng.normal(0, 1, 1422), obs_var=rng.uniform(0.05, 4, 1422), obs_count=np.float64(1000.0),
                 amp_mean=rng.normal(0, 1, 3090), amp_var=rng.uniform(0.05, 4, 3090), amp_count=np.float64(500.0))
    return d, stats


def reference_update_step(B=48, Ba=40, seed=61, wseed=21):
    """f1: one `calc_gradients` call (amp_continuous_value.py:276-428) on the reference's own network and loss code - body of
    the method from set_train() to the assembled loss executed by line range (ref_extract `Holder.calc_loss`), dropout masks from
    the reference's get_dropout_mask with recorded draws, `loss.backward()`, nn.utils.clip_grad_norm_(50) and one
    torch.optim.Adam(lr 2e-5, eps 1e-8) step (common_agent.py:84-87).  rl_games' model wrapper (ModelA2CContinuousLogStd, 1.1.4,
    not in the tree) is restated inline: mu / logstd / value from the network, neglogp and entropy of Normal(mu, exp(logstd)).
    The 11.2 M-element gradient does not fit a fixture: per parameter tensor the fixture keeps the L2 norm, the sum and 48 sampled
    entries of the gradient and of the Adam update; the tests recompute all of them."""
    import types
    import yaml
    from . import netweights
    R = ref_extract.load()
    torch = R.torch
    cfgy = yaml.safe_load(open(os.path.join(ref_extract.PACER, "data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml")))["params"]["config"]
    U = UPDATE_CFG
    for a, b in (("e_clip", "e_clip"), ("actor_coef", "actor_coef"), ("critic_coef", "critic_coef"), ("tv_coef", "tv_coef"),
                 ("bounds_loss_coef", "bounds_loss_coef"), ("entropy_coef", "entropy_coef"), ("disc_coef", "disc_coef"),
                 ("disc_logit_reg", "disc_logit_reg"), ("disc_grad_penalty", "disc_grad_penalty"), ("disc_weight_decay", "disc_weight_decay"),
                 ("lr", "learning_rate"), ("grad_norm", "grad_norm")):
        assert float(cfgy[b]) == float(U[a]), (a, cfgy[b])
    net = ref_extract.load_network(seed=1)
    sd = netweights.synth_state_dict(wseed, mu_gain=UPDATE_MU_GAIN)
    net.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()}, strict=True)
    net.train()
    batch, stats = synth_update_batch(B, Ba, seed)
    # the rollout that "collected" this minibatch: actions sampled around the current policy's mean, old neglogp a little off
    rng = np.random.default_rng(seed + 1)
    with torch.no_grad():
        import contextlib as _c, io as _io
        with _c.redirect_stdout(_io.StringIO()):
            tmp = R.RunningMeanStd((1422,))
        tmp.running_mean[:] = torch.from_numpy(stats["obs_mean"]); tmp.running_var[:] = torch.from_numpy(stats["obs_var"]); tmp.eval()
        mu0, ls0 = net.eval_actor(tmp(torch.from_numpy(batch["obs"])))
        mu0, sg0 = mu0.numpy(), np.exp(ls0.numpy())
    batch["actions"] = (mu0 + sg0 * rng.normal(0, 1, mu0.shape)).astype(np.float32)
    nl0 = 0.5 * (((batch["actions"] - mu0) / sg0) ** 2).sum(-1) + 0.5 * np.log(2 * np.pi) * 69 + np.log(sg0).sum(-1)
    batch["old_logp_actions"] = (nl0 + rng.normal(0, 0.15, B)).astype(np.float32)
    batch["mu"] = (mu0 + rng.normal(0, 0.02, mu0.shape)).astype(np.float32)
    H = ref_extract.load_agent_blocks()
    h = H()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        h.running_mean_std, h._amp_input_mean_std = R.RunningMeanStd((1422,)), R.RunningMeanStd((3090,))
    for m, pre in ((h.running_mean_std, "obs"), (h._amp_input_mean_std, "amp")):
        m.running_mean[:] = torch.from_numpy(stats[pre + "_mean"]); m.running_var[:] = torch.from_numpy(stats[pre + "_var"])
        m.count.fill_(float(stats[pre + "_count"]))
    h.set_train = lambda: (h.running_mean_std.train(), h._amp_input_mean_std.train())
    h.normalize_input = h._normalize_amp_input = True
    h._amp_minibatch_size = Ba
    h.last_lr, h.e_clip = U["lr"], U["e_clip"]
    h.config = {"amp_dropout": True}
    h.vec_env = types.SimpleNamespace(env=types.SimpleNamespace(task=types.SimpleNamespace(_num_amp_obs_steps=15, cfg={"env": {}})))
    h.motion_sym_loss = h.is_rnn = h.mixed_precision = h.clip_value = False
    h.critic_coef, h._actor_coef, h.entropy_coef, h.bounds_loss_coef = U["critic_coef"], U["actor_coef"], U["entropy_coef"], U["bounds_loss_coef"]
    h._disc_coef, h._tv_coef = U["disc_coef"], U["tv_coef"]
    h._disc_logit_reg, h._disc_grad_penalty, h._disc_weight_decay = U["disc_logit_reg"], U["disc_grad_penalty"], U["disc_weight_decay"]
    # dropout draws replayed: the hosted module's `torch.rand` returns the recorded uniforms, one [Ba, 3] per joint
    mod = sys_modules_get("emloco_ref_agent")
    draws = iter(torch.from_numpy(batch["dropout_u"][j]) for j in range(19))

    class TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        def rand(self, *a, **kw):
            t = next(draws)
            assert tuple(a) == tuple(t.shape), (a, t.shape)
            return t
    captured = {}

    def model(bd):
        # rl_games 1.1.4 ModelA2CContinuousLogStd.Network.forward (is_train branch), then learning/amp_models.py:20-44 and
        # learning/amp_sept_value_models.py:22-30
        mu, logstd, value, states = net(bd)                                   # AMPBuilder.Network.forward (amp_network_builder.py:39-48)
        sigma = torch.exp(logstd)
        distr = torch.distributions.Normal(mu, sigma)
        entropy = distr.entropy().sum(dim=-1)
        x = bd["prev_actions"]
        neglogp = 0.5 * (((x - mu) / sigma) ** 2).sum(dim=-1) + 0.5 * np.log(2.0 * np.pi) * x.size()[-1] + logstd.sum(dim=-1)
        res = {"prev_neglogp": torch.squeeze(neglogp), "values": value, "entropy": entropy, "rnn_states": states, "mus": mu, "sigmas": sigma}
        amp_obs, amp_replay, amp_demo = bd["amp_obs"], bd["amp_obs_replay"], bd["amp_obs_demo"]
        if bd.get("amp_dropout", False):
            mod.torch = TorchProxy()
            try:
                dm = h.get_dropout_mask(bd["amp_obs"], bd["amp_steps"], bd["env_cfg"])
            finally:
                mod.torch = torch
            captured["mask"] = dm.numpy().copy()
            amp_obs, amp_replay, amp_demo = (h.dropout_amp_obs(amp_obs, dm[..., 0]), h.dropout_amp_obs(amp_replay, dm[..., 1]),
                                             h.dropout_amp_obs(amp_demo, dm[..., 2]))
        res["disc_agent_logit"] = net.eval_disc(amp_obs)
        res["disc_agent_replay_logit"] = net.eval_disc(amp_replay)
        res["disc_demo_logit"] = net.eval_disc(amp_demo)
        res["task_values"] = net.eval_task_value(bd["obs"])
        return res
    model.a2c_network = net
    model.parameters = net.parameters
    h.model = model
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    inp = {k: T(batch[k]) for k in ("old_values", "old_logp_actions", "advantages", "mu", "sigma", "returns", "actions", "obs", "amp_obs",
                                    "amp_obs_replay", "amp_obs_demo")}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loc = h.calc_loss(inp)
    loss = loc["loss"]
    params = dict(net.named_parameters())
    opt = torch.optim.Adam(net.parameters(), float(U["lr"]), eps=1e-08, weight_decay=0)        # common_agent.py:84-87
    for p in net.parameters():
        p.grad = None
    loss.backward()
    grads = {k: (p.grad.detach().numpy().copy() if p.grad is not None else None) for k, p in params.items()}
    total_norm = float(torch.nn.utils.clip_grad_norm_(net.parameters(), U["grad_norm"]))
    before = {k: p.detach().numpy().copy() for k, p in params.items()}
    opt.step()
    with torch.no_grad():                                                               # torch_ext.policy_kl (rl_games 1.1.4), reduce=True
        p0m, p0s, p1m, p1s = loc["mu"].detach(), loc["sigma"].detach(), inp["mu"], inp["sigma"]
        c1 = torch.log(p1s / p0s + 1e-5); c2 = (p0s ** 2 + (p1m - p0m) ** 2) / (2.0 * (p1s ** 2 + 1e-5))
        kl = (c1 + c2 - 0.5).sum(dim=-1).mean().item()
    out = dict(B=B, Ba=Ba, seed=seed, wseed=wseed, in_actions=batch["actions"], in_old_logp_actions=batch["old_logp_actions"], in_mu=batch["mu"],
               kl=np.float32(kl), weights_checksum=netweights.checksum(sd), loss=np.float32(loss.item()),
               a_loss=np.float32(loc["a_loss"].item()), c_loss=np.float32(loc["c_loss"].item()), b_loss=np.float32(loc["b_loss"].item()),
               tv_loss=np.float32(loc["tv_loss"].item()), entropy=np.float32(loc["entropy"].item()),
               disc_loss=np.float32(loc["disc_loss"].item()), disc_grad_penalty=np.float32(loc["disc_info"]["disc_grad_penalty"].item()),
               disc_logit_loss=np.float32(loc["disc_info"]["disc_logit_loss"].item()),
               disc_agent_acc=np.float32(loc["disc_info"]["disc_agent_acc"].item()), disc_demo_acc=np.float32(loc["disc_info"]["disc_demo_acc"].item()),
               a_clip_frac=np.float32(loc["a_info"]["actor_clipped"].float().mean().item()), total_norm=np.float32(total_norm),
               mus=loc["mu"].detach().numpy(), values=loc["values"].detach().numpy(), task_values=loc["task_values"].detach().numpy(),
               neglogp=loc["action_log_probs"].detach().numpy(), disc_agent_logit=loc["disc_agent_cat_logit"].detach().numpy(),
               disc_demo_logit=loc["disc_demo_logit"].detach().numpy(), dropout_mask=captured["mask"][:, :206, :].astype(np.uint8),
               obs_mean_after=h.running_mean_std.running_mean.numpy().copy(), obs_var_after=h.running_mean_std.running_var.numpy().copy(),
               obs_count_after=np.float64(h.running_mean_std.count.item()),
               amp_mean_after=h._amp_input_mean_std.running_mean.numpy().copy(), amp_var_after=h._amp_input_mean_std.running_var.numpy().copy(),
               amp_count_after=np.float64(h._amp_input_mean_std.count.item()))
    assert (captured["mask"].reshape(Ba, 15, 206, 3) == captured["mask"][:, None, :206, :]).all()      # the 15 steps share one mask
    srng = np.random.default_rng(seed + 5)
    for k, p in params.items():
        g = grads[k]
        if g is None:
            assert k == "sigma"
            continue
        idx = srng.integers(0, g.size, 48)
        out[f"gidx_{k}"] = idx
        out[f"gnorm_{k}"] = np.float64(np.linalg.norm(g.astype(np.float64)))
        out[f"gsum_{k}"] = np.float64(g.astype(np.float64).sum())
        out[f"gval_{k}"] = g.reshape(-1)[idx].copy()
        out[f"step_{k}"] = (p.detach().numpy().reshape(-1)[idx] - before[k].reshape(-1)[idx]).astype(np.float32)
    return out


def sys_modules_get(name):
    import sys
    return sys.modules[name]



def reference_rms_update(seed=51):
    """RunningMeanStd in training mode (utils/running_mean_std.py:86-96): statistics after each of three batches."""
    R = ref_extract.load()
    torch = R.torch
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RunningMeanStd((7,))
    m.train()
    rng = np.random.default_rng(seed)
    out = {}
    for i, b in enumerate((16, 5, 64)):
        x = (rng.normal(0, 1 + i, (b, 7)) + i).astype(np.float32)
        m(torch.from_numpy(x))
        out.update({f"x{i}": x, f"mean{i}": m.running_mean.numpy().copy(), f"var{i}": m.running_var.numpy().copy(),
                    f"count{i}": np.asarray(m.count.numpy()).copy()})
    return out


def reference_player(seed=71, step_to_pred=6):
    """SURVEY component 14: the step loop of AMPPlayerContinuousValue.run (amp_value_players.py:123-198) executed from the
    reference on a holder whose env / value net return recorded tensors - three games of one env: one that runs past
    step_to_pred, one that ends early, one that ends exactly at step_to_pred.  The per-game locals are re-initialised as
    :76-100 does."""
    import types
    R = ref_extract.load()
    torch = R.torch
    H = ref_extract.load_player_block()
    h = H()
    rng = np.random.default_rng(seed)
    f = np.float32
    cur = {}
    h.env = types.SimpleNamespace(task=types.SimpleNamespace(_humanoid_root_states=torch.zeros(1, 13)), get_waypoint_traj=lambda: torch.zeros(1, 13, 3),
                                  get_init_pose=lambda: torch.zeros(1, 24, 3), get_init_vel=lambda: torch.zeros(1, 2),
                                  raw_reward=lambda: cur["raw"])
    h.env_step = lambda env, a: ({"obs": None}, cur["r"].clone(), cur["done"], {})
    h.inversion_penalty_scale = 0.3
    h.device = "cpu"
    h.valuenet = lambda w, p, v: cur["score"]
    h.visualized_pose = True
    h._post_step = lambda info: (0.0, 0.0, float(cur["disc"]))
    h.gamma, h.step_to_pred, h.num_agents, h.plot_val_reward, h.is_rnn = 0.99, step_to_pred, 1, True, False
    h.criterion = torch.nn.MSELoss()
    bar = types.SimpleNamespace(set_description=lambda s: None, update=lambda k: None)
    games, steps_in = [], []
    rl, rp, rd = [], [], []
    for length, inverted in ((10, False), (4, True), (step_to_pred + 1, False)):
        L = dict(real_traj=[], inverted_envs=inverted, cr=torch.zeros(1), steps=torch.zeros(1), c_task_value=0, c_critic_value=0,
                 c_disc_reward=0, c_loc_reward=0, c_pow_reward=0, rew_disc_coef=1.0, max_frame_rew=-100, min_frame_rew=100,
                 rew_lists={'total': [], 'loc': [], 'disc': [], 'pow': []}, bar=bar, t=len(games), games_played=0, rewards_loc=rl, rewards_pow=rp,
                 rewards_disc=rd, min_reward=-10, max_reward=100, total_value_loss=0)
        score = f(rng.uniform(0.2, 0.9))
        ins = []
        import contextlib, io
        for n in range(length):
            raw = rng.uniform([0.0, -0.3], [1.0, 0.0]).astype(f)
            logit = f(rng.normal(0, 3))
            prob = 1 / (1 + np.exp(-logit)); disc = f(-np.log(max(1 - prob, 0.0001)) * 2)
            cur.update(raw=torch.from_numpy(raw[None]), r=torch.tensor([raw.sum()]), done=torch.tensor([int(n == length - 1)]), disc=disc,
                       score=torch.tensor([[score]]))
            with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
                L = h.player_block(n, L)
            ins.append(dict(raw=raw, logit=logit, disc=disc, done=int(n == length - 1)))
        games.append(dict(score=score, inverted=inverted, cr_to_pred=f(L["cr_to_pred"].item()), norm_reward=f(L["norm_rewards"].item()),
                          value_loss=f(L["value_loss"].item()), steps=length))
        steps_in.append(ins)
    out = dict(step_to_pred=step_to_pred, n_games=len(games), rewards_loc=np.array(rl, f), rewards_pow=np.array(rp, f), rewards_disc=np.array(rd, f))
    for i, (gm, ins) in enumerate(zip(games, steps_in)):
        out.update({f"g{i}_{k}": v for k, v in gm.items()})
        out[f"g{i}_raw"] = np.stack([x["raw"] for x in ins]); out[f"g{i}_logit"] = np.array([x["logit"] for x in ins], f)
        out[f"g{i}_disc"] = np.array([x["disc"] for x in ins], f)
    return out


def reference_motion_lib(n=40, seed=81):
    """f2: MotionLibSMPL.get_motion_state_smpl (utils/motion_lib_smpl.py:485-563) and HumanoidAMP.build_amp_obs_demo
    (env/tasks/humanoid_amp.py:186-211, smpl branch -> build_amp_observations_smpl) executed from the reference on the synthetic
    clips of emloco_b200.synthetic.synthetic_motion_lib(6, seed) - the fixture stores sampled ids / times and the outputs."""
    from emloco_b200.synthetic import synthetic_motion_lib
    R = ref_extract.load()
    torch = R.torch
    lib = synthetic_motion_lib(6, seed)
    H = ref_extract.load_motion_lib_block()
    h = H()
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    h.num_bodies = 24
    h._key_body_ids = torch.tensor(KEY_BODIES)
    h._motion_lengths, h._motion_num_frames, h._motion_dt = T(lib["motion_lengths"]), T(lib["motion_num_frames"].astype(np.int64)), T(lib["motion_dt"])
    h.length_starts = T(lib["length_starts"].astype(np.int64))
    h.lrs, h.gvs, h.gavs, h.gts, h.dvs, h.grs = (T(lib[k]) for k in ("lrs", "gvs", "gavs", "gts", "dvs", "grs"))
    h._motion_aa = torch.zeros(lib["gts"].shape[0], 72)
    h._motion_bodies, h._motion_limb_weights = T(lib["motion_bodies"]), torch.zeros(len(lib["motion_lengths"]), 10)
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, 6, n).astype(np.int64)
    times = (rng.random(n) * lib["motion_lengths"][ids]).astype(np.float32)
    times[:4] = [0.0, -0.2, lib["motion_lengths"][ids[2]], lib["motion_lengths"][ids[3]] + 0.3]        # both ends and beyond
    ms = h.get_motion_state_smpl(T(ids), T(times))
    # build_amp_obs_demo (humanoid_amp.py:186-211): 15 steps back in time, newest first
    dt = CONTROL_DT
    mi = torch.tile(T(ids).unsqueeze(-1), [1, AMP_STEPS]).view(-1)
    mt = (T(times).unsqueeze(-1) + (-dt * torch.arange(0, AMP_STEPS))).view(-1).float()
    r = h.get_motion_state_smpl(mi, mt)
    demo = R.jit.build_amp_observations_smpl(r["root_pos"], r["root_rot"], r["root_vel"], r["root_ang_vel"], r["dof_pos"], r["dof_vel"], r["key_pos"],
                                             r["motion_bodies"], r["motion_limb_weights"], torch.from_numpy(DOF_SUBSET), True, False, True, True, False, True)
    out = dict(lib_seed=seed, lib_motions=6, ids=ids, times=times, demo=demo.view(n, -1).numpy())
    for k in ("root_pos", "root_rot", "dof_pos", "root_vel", "root_ang_vel", "dof_vel", "key_pos", "rg_pos", "rb_rot", "body_vel", "body_ang_vel"):
        out["ms_" + k] = ms[k].numpy()
    return out


def reference_plausibl(B=40, seed=41):
    """a17: plausibl/test_value_mlp.py:24-113 `MLP` (24 -> 12 -> 6 -> 1, no sigmoid), biases randomised."""
    R = ref_extract.load()
    torch = R.torch
    MLP = ref_extract.load_plausibl_mlp()
    torch.manual_seed(seed)
    m = MLP()
    with torch.no_grad():
        for lay in (m._value_mlp[0], m._value_mlp[2], m._value_logits):
            lay.bias.uniform_(-0.3, 0.3)
    x = np.random.default_rng(seed).normal(0, 2, (B, 24)).astype(np.float32)
    with torch.no_grad():
        y = m.forward(torch.from_numpy(x)).numpy()
    sd = {f"_value_mlp.{k}": v.numpy().copy() for k, v in m._value_mlp.state_dict().items()}
    sd.update({f"_value_logits.{k}": v.numpy().copy() for k, v in m._value_logits.state_dict().items()})
    return dict(x=x, y=y, **{f"w_{k}": v for k, v in sd.items()})


def reference_pd_table():
    """a2: the PD target offset / scale table.  The reference's own `Humanoid._build_pd_action_offset_scale` (humanoid.py:949-1025,
    hosted by line range) on the joint ranges of the reference MJCF (what Isaac Gym reports as DOF limits), with the flags of
    data/cfg/pacer.yaml (bias_offset False, has_upright_start True, has_smpl_pd_offset False) and with has_smpl_pd_offset True."""
    import types
    R = ref_extract.load()
    torch = R.torch
    from emloco_b200.mjcf import load_mjcf
    mdl = load_mjcf(os.path.join(ref_extract.PACER, "data/assets/mjcf/smpl_humanoid.xml"))
    out = dict(limit_lo=np.asarray(mdl.limit_lo, np.float64), limit_hi=np.asarray(mdl.limit_hi, np.float64), names=np.array(mdl.names))
    for tag, smpl_off in (("", False), ("_smpl_offset", True)):
        h = ref_extract.load_pd_block()()
        h._dof_offsets = list(range(0, ND + 1, 3))
        h.dof_limits_lower = torch.tensor(out["limit_lo"], dtype=torch.float32)
        h.dof_limits_upper = torch.tensor(out["limit_hi"], dtype=torch.float32)
        h._bias_offset, h.smpl_humanoid, h._has_smpl_pd_offset, h._has_upright_start, h.device = False, True, smpl_off, True, "cpu"
        h._dof_names = list(mdl.names[1:])
        h._build_pd_action_offset_scale()
        out["offset" + tag] = h._pd_action_offset.numpy().copy()
        out["scale" + tag] = h._pd_action_scale.numpy().copy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "pd_table.npz"), **reference_pd_table())
    for name, fl in (("traj_reset_plain", 0), ("traj_reset_train", O.TRAJ_F_REAL | O.TRAJ_F_ADJUST_VEL | O.TRAJ_F_INIT_HEADING),
                     ("traj_reset_inv", O.TRAJ_F_REAL | O.TRAJ_F_INIT_HEADING | O.TRAJ_F_INVERSION | O.TRAJ_F_SLOW)):
        g = reference_traj_reset(24, 10, 11 + fl, fl)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **g)
        print(name, g["verts"].shape)
    for name, N, seed, rough in (("post_step_rough", 8, 0, True), ("post_step_flat", 6, 1, False)):
        st = synth_state(N, seed, map_shape=(700, 700), rough=rough)
        out = reference_post_step(st)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), seed=seed, N=N, rough=rough,
                            **{f"in_{k}": v for k, v in st.items()}, **{f"out_{k}": v for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})
    # every (reset, terminate) branch of compute_humanoid_reset at 256 envs; inputs are regenerated from the seed by the tests
    # (synth_state_branches), the fixture holds the reference outputs (AMP ring: only the new step - the rest is a shifted copy)
    st = synth_state_branches(256, 7)
    out = reference_post_step(st)
    fallen_only = (out["terminate"] == 1) & (((out["tar_pos"][:, :2] - st["rb"][:, 0, :2]) ** 2).sum(-1) <= 16)
    lost_only = (out["terminate"] == 1) & ~fallen_only
    combos = {(int(a), int(b)) for a, b in zip(out["reset"], out["terminate"])}
    assert combos == {(0, 0), (1, 0), (1, 1)} and fallen_only.sum() > 20 and lost_only.sum() > 5 and ((out['reset'] == 1) & (out['terminate'] == 0)).sum() > 5, (combos, fallen_only.sum(), lost_only.sum())
    assert ((st["progress"] <= 1) & (out["terminate"] == 0)).sum() > 10
    out["amp_obs"] = out["amp_obs"][:, :AMP_STEP_DIM].copy()
    np.savez_compressed(os.path.join(OUT, "post_step_branches.npz"), seed=7, N=256, **{f"out_{k}": v for k, v in out.items()})
    print("post_step_branches", sorted(combos), int(fallen_only.sum()), int(lost_only.sum()))
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **reference_nets())
    np.savez_compressed(os.path.join(OUT, "play_block.npz"), **reference_play_block())
    np.savez_compressed(os.path.join(OUT, "plausibl.npz"), **reference_plausibl())
    np.savez_compressed(os.path.join(OUT, "rms_update.npz"), **reference_rms_update())
    np.savez_compressed(os.path.join(OUT, "update_step.npz"), **reference_update_step())
    np.savez_compressed(os.path.join(OUT, "player.npz"), **reference_player())
    np.savez_compressed(os.path.join(OUT, "motion_lib.npz"), **reference_motion_lib())
    traj, pose, vel = synth_locoval(64, 2)
    W, out = reference_locoval(traj, pose, vel)
    np.savez_compressed(os.path.join(OUT, "locoval.npz"), traj=traj, pose=pose, vel=vel,
                        **{f"w_{k}": v for k, v in W.items()}, **{f"out_{k}": v for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, "locoval_mm5.npz"), **reference_locoval_multimodal())
    np.savez_compressed(os.path.join(OUT, "locoval_finetune.npz"), **reference_locoval_finetune())
    g = reference_gae(32, 16, 3)
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **g)
    rng = np.random.default_rng(4)
    x = rng.normal(0, 3, (16, 1422)).astype(np.float32)
    mean = rng.normal(0, 1, 1422); var = rng.uniform(0.0, 4, 1422)
    np.savez_compressed(os.path.join(OUT, "rms.npz"), x=x, mean=mean, var=var, y=reference_rms(x, mean, var))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
