"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's hot-path algorithm.

This module is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``emloco_b200``) never does; it fails loudly when its CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to
``pacer/pacer/`` in ImIntheMiddle/EmLoco).  Parity status:

* everything in this file except the physics step is PINNED against outputs of the
  reference's own functions executed in the build container
  (``oracle/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).
* the policy ``neglogp`` formula restates rl_games 1.1.4 (pinned dependency in
  pacer/requirements.txt, not vendored in the reference) -> "parity unpinned" for that
  one formula.
* the physics step lives in ``oracle/physics_oracle.c`` (fp64) and is "parity unpinned"
  against PhysX: the reference ships no binaries, no tests and no golden vectors.

All arithmetic is float32 unless noted, masks/counters int64, like the reference.
"""
from __future__ import annotations

import numpy as np

F = np.float32

# ----------------------------------------------------------------------------------------
# constants of the default configuration
# ----------------------------------------------------------------------------------------
NB, ND = 24, 69
NUM_TRAJ_SAMPLES = 15            # pacer.yaml:53
TRAJ_SAMPLE_DT = 0.4             # pacer.yaml:54
NUM_VERTS = 101                  # humanoid_traj.py:113
EPISODE_LEN = 168                # pacer.yaml:12
SIM_DT = 1.0 / 60.0              # utils/config.py:24
CONTROL_DT = 2 * SIM_DT          # humanoid.py:89 (controlFrequencyInv=2)
AMP_STEPS, AMP_STEP_DIM = 15, 206
HEAD = 13
KEY_BODIES = [7, 3, 22, 17]      # R_Ankle, L_Ankle, R_Wrist, L_Wrist (pacer.yaml:50)
CONTACT_BODIES = [7, 3, 8, 4]    # R_Ankle, L_Ankle, R_Toe, L_Toe (pacer.yaml:51)
LEFT_TO_RIGHT = [0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 19, 20, 21, 22, 23, 14, 15, 16, 17, 18]
# dof_subset: every joint except L/R_Hand, L/R_Toe (humanoid.py:290-326)
_DOF_NAMES_REMOVED = {3, 7, 17, 22}   # joint index (body index - 1) of L_Toe, R_Toe, L_Hand, R_Hand
DOF_SUBSET = np.concatenate([np.arange(3 * j, 3 * j + 3) for j in range(23) if j not in _DOF_NAMES_REMOVED])
H_SCALE, V_SCALE = 0.1, 0.005    # humanoid_pedestrain_terrain.py:1142-1143
HEIGHT_MEAS_SCALE = 5.0          # :80
FAIL_DIST = 4.0                  # humanoid_traj.py (fail_dist)


# ----------------------------------------------------------------------------------------
# quaternion helpers, xyzw (isaacgym/torch_utils.py, utils/torch_utils.py)
# ----------------------------------------------------------------------------------------
def quat_mul(a, b):
    """isaacgym/python/isaacgym/torch_utils.py:19-41"""
    x1, y1, z1, w1 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    x2, y2, z2, w2 = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = F(0.5) * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return np.stack([x, y, z, w], axis=-1).astype(F)


def my_quat_rotate(q, v):
    """utils/torch_utils.py:15-25"""
    q_w = q[..., 3:4]
    q_vec = q[..., :3]
    a = v * (F(2.0) * q_w ** 2 - F(1.0))
    b = np.cross(q_vec, v) * q_w * F(2.0)
    c = q_vec * np.sum(q_vec * v, axis=-1, keepdims=True) * F(2.0)
    return (a + b + c).astype(F)


def quat_apply(a, b):
    """isaacgym torch_utils.py:49-56"""
    xyz = a[..., :3]
    t = np.cross(xyz, b) * F(2)
    return (b + a[..., 3:4] * t + np.cross(xyz, t)).astype(F)


def normalize(x, eps=1e-9):
    """isaacgym torch_utils.py:44-46"""
    n = np.linalg.norm(x, axis=-1, keepdims=True).astype(F)
    return (x / np.maximum(n, F(eps))).astype(F)


def quat_from_angle_axis(angle, axis):
    """isaacgym torch_utils.py:98-103"""
    theta = (angle / F(2))[..., None]
    xyz = normalize(axis) * np.sin(theta)
    w = np.cos(theta)
    return normalize(np.concatenate([xyz, w], axis=-1).astype(F))


def normalize_angle(x):
    """isaacgym torch_utils.py:106-108"""
    return np.arctan2(np.sin(x), np.cos(x)).astype(F)


def quat_to_tan_norm(q):
    """utils/torch_utils.py:66-79"""
    ref_tan = np.zeros(q.shape[:-1] + (3,), F); ref_tan[..., 0] = 1
    ref_norm = np.zeros(q.shape[:-1] + (3,), F); ref_norm[..., 2] = 1
    return np.concatenate([my_quat_rotate(q, ref_tan), my_quat_rotate(q, ref_norm)], axis=-1)


def calc_heading(q):
    """utils/torch_utils.py:137-148"""
    ref = np.zeros(q.shape[:-1] + (3,), F); ref[..., 0] = 1
    rot = my_quat_rotate(q, ref)
    return np.arctan2(rot[..., 1], rot[..., 0]).astype(F)


def calc_heading_quat(q):
    """utils/torch_utils.py:150-161"""
    axis = np.zeros(q.shape[:-1] + (3,), F); axis[..., 2] = 1
    return quat_from_angle_axis(calc_heading(q), axis)


def calc_heading_quat_inv(q):
    """utils/torch_utils.py:163-174"""
    axis = np.zeros(q.shape[:-1] + (3,), F); axis[..., 2] = 1
    return quat_from_angle_axis(-calc_heading(q), axis)


def exp_map_to_quat(e):
    """utils/torch_utils.py:89-112 (exp_map_to_angle_axis + quat_from_angle_axis)"""
    angle = np.linalg.norm(e, axis=-1).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        axis = (e / angle[..., None]).astype(F)
    angle = normalize_angle(angle)
    default_axis = np.zeros_like(e); default_axis[..., 2] = 1
    mask = np.abs(angle) > F(1e-5)
    angle = np.where(mask, angle, F(0)).astype(F)
    axis = np.where(mask[..., None], axis, default_axis).astype(F)
    return quat_from_angle_axis(angle, axis)


def quat_apply_yaw(quat, vec):
    """env/tasks/humanoid_pedestrain_terrain.py:1533-1538"""
    qy = quat.copy()
    qy[..., :2] = 0
    return quat_apply(normalize(qy), vec)


# ----------------------------------------------------------------------------------------
# a4: self observation (+ mirrored copy)
# ----------------------------------------------------------------------------------------
def self_obs(body_pos, body_rot, body_vel, body_ang_vel, betas):
    """compute_humanoid_observations_smpl_max, env/tasks/humanoid.py:1626-1687 with the
    default flags local_root_obs=True, root_height_obs=False, upright=True,
    has_smpl_params=True, has_limb_weight_params=False.  -> [N,368]"""
    N = body_pos.shape[0]
    root_pos = body_pos[:, 0]
    root_rot = body_rot[:, 0]
    hinv = calc_heading_quat_inv(root_rot)                  # [N,4]
    hexp = np.repeat(hinv[:, None], NB, 1).reshape(-1, 4)
    local_pos = (body_pos - root_pos[:, None]).reshape(-1, 3)
    local_pos = my_quat_rotate(hexp, local_pos).reshape(N, NB * 3)[:, 3:]
    local_rot = quat_mul(hexp, body_rot.reshape(-1, 4))
    rot_obs = quat_to_tan_norm(local_rot).reshape(N, NB * 6)
    vel = my_quat_rotate(hexp, body_vel.reshape(-1, 3)).reshape(N, NB * 3)
    ang = my_quat_rotate(hexp, body_ang_vel.reshape(-1, 3)).reshape(N, NB * 3)
    return np.concatenate([local_pos, rot_obs, vel, ang, betas[:, :-6]], axis=-1).astype(F)


def flip_state(body_pos, body_rot, body_vel, body_ang_vel):
    """_compute_flip_humanoid_obs, env/tasks/humanoid.py:1066-1090: y-mirror + L<->R."""
    p = body_pos.copy(); p[..., 1] *= -1; p = p[:, LEFT_TO_RIGHT]
    r = body_rot.copy(); r[..., 0] *= -1; r[..., 2] *= -1; r = r[:, LEFT_TO_RIGHT]
    v = body_vel.copy(); v[..., 1] *= -1; v = v[:, LEFT_TO_RIGHT]
    w = body_ang_vel.copy(); w[..., 0] *= -1; w[..., 2] *= -1; w = w[:, LEFT_TO_RIGHT]
    return p, r, v, w


# ----------------------------------------------------------------------------------------
# a5/a6: task observation
# ----------------------------------------------------------------------------------------
def traj_dt():
    """TrajGenerator.__init__: _dt = episode_dur/(num_verts-1) (env/util/traj_generator.py:24),
    episode_dur = max_episode_length*dt (humanoid_traj.py:112)."""
    return (EPISODE_LEN * CONTROL_DT) / (NUM_VERTS - 1)


def calc_pos(verts, traj_ids, times):
    """TrajGenerator.calc_pos, env/util/traj_generator.py:278-296.  verts [N,101,3]."""
    traj_dur = NUM_VERTS * traj_dt()                         # get_traj_duration :269-272
    num_segs = NUM_VERTS - 1
    phase = np.clip(times.astype(F) / F(traj_dur), F(0), F(1)).astype(F)
    seg = (phase * F(num_segs)).astype(F)
    i0 = np.floor(seg).astype(np.int64)
    i1 = np.ceil(seg).astype(np.int64)
    lerp = (seg - i0.astype(F)).astype(F)[..., None]
    flat = verts.reshape(-1, 3)
    p0 = flat[traj_ids * NUM_VERTS + i0]
    p1 = flat[traj_ids * NUM_VERTS + i1]
    return ((F(1.0) - lerp) * p0 + lerp * p1).astype(F)


TRAJ_RAND_COLS = 405   # uniform draws consumed per reset env, in the order TrajGenerator.reset draws them (see traj_reset)
TRAJ_F_REAL, TRAJ_F_ADJUST_VEL, TRAJ_F_INIT_HEADING, TRAJ_F_INVERSION, TRAJ_F_SLOW = 1, 2, 4, 8, 16


def traj_reset(verts, env_ids, init_pos, root_vel, U, flags, pool=None, dtheta_max=2.0, speed_min=0.0005, speed_max=3.0,
               accel_max=2.0, sharp_turn_prob=0.02, hybrid_init_prob=0.5):
    """TrajGenerator.reset (env/util/traj_generator.py:60-237) with the random draws made explicit.

    U [n, 405] uniform [0,1): cols 0-99 torch.rand dtheta (:63), 100-199 sharp angles (:66), 200-299 the bernoulli draw
    (:68, mask = u < p), 300 heading (:71), 301-400 dspeed (:74), 401 initial speed (:76), 402 real-data draw (:118),
    403 pool pick (random.sample :128, here floor(u*P) - with replacement), 404 heading inversion (:196).
    Flags: real_path (:116), adjust_root_vel (:98,:149), init_heading (:176), heading_inversion (:195), slow (:94).
    Updates verts [N,101,3] in place for env_ids; returns inverted [n] bool."""
    f = np.float32
    n = len(env_ids)
    nv = verts.shape[1]
    S = nv - 1
    dt = f(traj_dt())
    U = U.astype(f)
    init_pos = init_pos.astype(f); root_vel = root_vel.astype(f)
    dtheta = (f(2) * U[:, 0:S] - f(1)) * f(dtheta_max) * dt
    sharp = f(np.pi) * (f(2) * U[:, 100:100 + S] - f(1))
    mask = U[:, 200:200 + S] < f(sharp_turn_prob)
    dtheta[mask] = sharp[mask]
    dtheta[:, 0] = f(np.pi) * (f(2) * U[:, 300] - f(1))
    dspeed = (f(2) * U[:, 301:301 + S] - f(1)) * f(accel_max) * dt
    dspeed[:, 0] = f(speed_max - speed_min) * U[:, 401] + f(speed_min)
    speed = np.zeros_like(dspeed)
    speed[:, 0] = dspeed[:, 0]
    for i in range(1, S):
        speed[:, i] = np.clip(speed[:, i - 1] + dspeed[:, i], f(speed_min), f(speed_max))
    if flags & TRAJ_F_SLOW:
        speed = speed / f(4)
    if flags & TRAJ_F_ADJUST_VEL:
        root_speed = np.linalg.norm(root_vel[:, :2], axis=-1)
        speed = np.clip((root_speed / speed[:, 0])[:, None] * speed, f(speed_min), f(speed_max))
    theta = np.cumsum(dtheta, -1, dtype=f)
    dpos = np.stack([np.cos(theta), -np.sin(theta), np.zeros_like(theta)], -1) * (speed * dt)[..., None]
    dpos[:, 0, 0:2] += init_pos[:, 0:2]
    verts[env_ids, 0, 0:2] = init_pos[:, 0:2]
    verts[env_ids, 1:] = np.cumsum(dpos, -2, dtype=f)
    if flags & TRAJ_F_REAL:
        real = U[:, 402] > f(hybrid_init_prob)
        P = pool.shape[0]
        pick = np.minimum((U[:, 403] * f(P)).astype(np.int64), P - 1)[real]
        traj = pool[pick][:, :nv].astype(f).copy()
        traj[..., 0:2] -= traj[:, :1, 0:2].copy()
        if flags & TRAJ_F_ADJUST_VEL:
            init_speed = np.maximum(np.linalg.norm(traj[:, 1] - traj[:, 0], axis=-1), f(speed_min) * dt)
            ratio = np.linalg.norm(root_vel[real, :2], axis=-1) / init_speed * dt
            traj[..., 0:2] *= ratio[:, None, None]
        traj[..., 0:2] += init_pos[real, None, 0:2]
        verts[np.asarray(env_ids)[real]] = traj
    inverted = np.zeros(n, bool)
    if flags & TRAJ_F_INIT_HEADING:
        v = verts[env_ids].copy()
        dinit = v[:, 1, :2] - v[:, 0, :2]
        rmag = np.sqrt((root_vel ** 2).sum(1)); dmag = np.sqrt((dinit ** 2).sum(1))
        root_rot = np.where(rmag > 0, np.arctan2(root_vel[:, 1], root_vel[:, 0]), f(0))
        init_heading = np.where(dmag > 0, np.arctan2(dinit[:, 1], dinit[:, 0]), f(0))
        rot = (init_heading - root_rot).astype(f)
        if flags & TRAJ_F_INVERSION:
            inverted = U[:, 404] > f(0.5)
            rot[inverted] = init_heading[inverted] - root_rot[inverted] + f(np.pi)
        origin = v[:, :1, 0:2].copy()
        xy = v[:, :, 0:2] - origin
        c, s_ = np.cos(rot)[:, None], np.sin(rot)[:, None]
        v[:, :, 0] = xy[..., 0] * c + xy[..., 1] * s_ + origin[..., 0]        # bmm(xy, [[c,-s],[s,c]]) (:211-216)
        v[:, :, 1] = -xy[..., 0] * s_ + xy[..., 1] * c + origin[..., 1]
        verts[env_ids] = v
    return inverted


def reset_task_outputs(verts, env_ids, rb_pos, root_vel):
    """HumanoidPedestrianTerrain._reset_task (:493-516) after the generator: waypoint_traj = _fetch_traj_samples at progress 0,
    init_pose = rigid-body positions, init_vel = root xy velocity; returned as the vec-env getters expose them
    (vec_task_wrappers.py:47-63: origin-relative waypoints and pose)."""
    n = len(env_ids)
    ids = np.repeat(np.asarray(env_ids), NUM_TRAJ_SAMPLES)
    t = np.tile(np.arange(NUM_TRAJ_SAMPLES, dtype=np.float32) * np.float32(TRAJ_SAMPLE_DT), n)
    w = calc_pos(verts, ids, t).reshape(n, NUM_TRAJ_SAMPLES, 3)
    return (w - w[:, :1]).astype(np.float32), (rb_pos - rb_pos[:, :1]).astype(np.float32), root_vel[:, :2].astype(np.float32)


def progress_time(progress):
    """`self.progress_buf * self.dt` (int64 tensor * python float -> float32)."""
    return (progress.astype(F) * F(CONTROL_DT)).astype(F)


def fetch_traj_samples(verts, progress):
    """_fetch_traj_samples, env/tasks/humanoid_traj.py:208-224 -> [N,15,3]"""
    N = progress.shape[0]
    t0 = progress_time(progress)
    ts = (np.arange(NUM_TRAJ_SAMPLES, dtype=F) * F(TRAJ_SAMPLE_DT)).astype(F)
    tt = (t0[:, None] + ts[None]).astype(F)
    ids = np.broadcast_to(np.arange(N)[:, None], tt.shape)
    return calc_pos(verts, ids.reshape(-1), tt.reshape(-1)).reshape(N, NUM_TRAJ_SAMPLES, 3)


def location_obs(root_pos, root_rot, traj_samples):
    """compute_location_observations, humanoid_pedestrain_terrain.py:1549-1577 -> [N,30]"""
    N, S, _ = traj_samples.shape
    hinv = calc_heading_quat_inv(root_rot)
    hexp = np.repeat(hinv[:, None], S, 1).reshape(-1, 4)
    delta = (traj_samples - root_pos[:, None]).reshape(-1, 3)
    local = my_quat_rotate(hexp, delta)[:, :2]
    return local.reshape(N, S * 2).astype(F)


def square_height_points(extent=2.0, res=32):
    """init_square_height_points, humanoid_pedestrain_terrain.py:650-668 (float64 grid -> f32)."""
    y = np.linspace(-extent, extent, res)
    x = np.linspace(-extent, extent, res)
    gx, gy = np.meshgrid(x, y, indexing="ij")
    pts = np.zeros((res * res, 3), F)
    pts[:, 0] = gx.reshape(-1)
    pts[:, 1] = gy.reshape(-1)
    return pts


def center_height_points():
    """init_center_height_points, humanoid_pedestrain_terrain.py:631-647"""
    y = np.linspace(-0.2, 0.2, 3)
    x = np.linspace(-0.1, 0.1, 3)
    gx, gy = np.meshgrid(x, y, indexing="ij")
    pts = np.zeros((9, 3), F)
    pts[:, 0] = gx.reshape(-1)
    pts[:, 1] = gy.reshape(-1)
    return pts


def sample_height_points(height_samples, points):
    """Terrain.world_points_to_map + sample_height_points (non-group branch),
    humanoid_pedestrain_terrain.py:1212-1226,1282-1288.  height_samples int16 [R,C]."""
    p = (points / F(H_SCALE)).astype(F)
    idx = np.trunc(p).astype(np.int64)
    px = np.clip(idx[..., 0], 0, height_samples.shape[0] - 2)
    py = np.clip(idx[..., 1], 0, height_samples.shape[1] - 2)
    h1 = height_samples[px, py]
    h2 = height_samples[px + 1, py + 1]
    return (np.minimum(h1, h2).astype(F) * F(V_SCALE)).astype(F)


def get_heights(height_samples, pose_pos, pose_rot, grid):
    """get_heights, humanoid_pedestrain_terrain.py:761-815 (head-rooted) -> [N,P]"""
    N, P = pose_pos.shape[0], grid.shape[0]
    hq = calc_heading_quat(pose_rot)
    hexp = np.repeat(hq[:, None], P, 1).reshape(-1, 4)
    pts = quat_apply(hexp, np.tile(grid, (N, 1))).reshape(N, P, 3) + pose_pos[:, None]
    return sample_height_points(height_samples, pts.astype(F))


def get_center_heights(height_samples, root_pos, root_rot, grid9):
    """get_center_heights, humanoid_pedestrain_terrain.py:732-759 -> [N,9]"""
    N, P = root_pos.shape[0], grid9.shape[0]
    qexp = np.repeat(root_rot[:, None], P, 1).reshape(-1, 4)
    pts = quat_apply_yaw(qexp, np.tile(grid9, (N, 1))).reshape(N, P, 3) + root_pos[:, None]
    return sample_height_points(height_samples, pts.astype(F))


def task_obs(body_pos, body_rot, verts, progress, height_samples, grid=None, grid9=None):
    """_compute_task_obs, humanoid_pedestrain_terrain.py:394-452 with terrain_obs=True,
    terrain_obs_root='head', use_center_height=True, velocity_map=False -> [N,1054]"""
    grid = square_height_points() if grid is None else grid
    grid9 = center_height_points() if grid9 is None else grid9
    root_pos, root_rot = body_pos[:, 0], body_rot[:, 0]
    samples = fetch_traj_samples(verts, progress)
    obs = location_obs(root_pos, root_rot, samples)
    meas = get_heights(height_samples, body_pos[:, HEAD], body_rot[:, HEAD], grid)
    ch = get_center_heights(height_samples, root_pos, root_rot, grid9).mean(axis=-1, keepdims=True, dtype=F)
    heights = (np.clip(ch - meas, F(-3), F(3.0)) * F(HEIGHT_MEAS_SCALE)).astype(F)
    return np.concatenate([obs, heights], axis=1).astype(F)


def flip_task_obs(tobs, res=32):
    """_compute_flip_task_obs, humanoid_pedestrain_terrain.py:455-491"""
    B = tobs.shape[0]
    traj = tobs[:, :2 * NUM_TRAJ_SAMPLES].reshape(B, NUM_TRAJ_SAMPLES, 2).copy()
    traj[..., 1] *= -1
    hm = tobs[:, 2 * NUM_TRAJ_SAMPLES:].reshape(B, res, res)[:, :, ::-1]
    return np.concatenate([traj.reshape(B, -1), hm.reshape(B, -1)], axis=1).astype(F)


# ----------------------------------------------------------------------------------------
# a7/a8: reward + reset
# ----------------------------------------------------------------------------------------
def reward(root_pos, tar_pos, dof_force, dof_vel, power_coef=0.0005, loc_coef=1.0):
    """_compute_reward + compute_location_reward, humanoid_pedestrain_terrain.py:907-930,1581-1592"""
    diff = tar_pos[..., 0:2] - root_pos[..., 0:2]
    err = np.sum(diff * diff, axis=-1, dtype=F)
    loc = (F(loc_coef) * np.exp(F(-2.0) * err)).astype(F)
    power = np.abs(dof_force * dof_vel).sum(axis=-1, dtype=F)
    pw = (F(-power_coef) * power).astype(F)
    return (loc + pw).astype(F), np.stack([loc, pw], axis=-1).astype(F)


def compute_reset(progress, contact, body_pos, tar_pos, max_len=EPISODE_LEN, fail_dist=FAIL_DIST):
    """compute_humanoid_reset, humanoid_pedestrain_terrain.py:1468-1530 (early termination on,
    collision check on) -> reset, terminated (int64)"""
    masked = contact.copy()
    masked[:, CONTACT_BODIES, :] = 0
    s = masked.sum(axis=-2, dtype=F)
    fallen = np.sqrt(np.square(np.abs(s)).sum(axis=-1, dtype=F)) > F(50)
    fallen = fallen & (progress > 1)
    delta = tar_pos[..., 0:2] - body_pos[:, 0, 0:2]
    d2 = np.sum(delta * delta, axis=-1, dtype=F)
    fail = d2 > F(fail_dist * fail_dist)
    term = (fallen | fail).astype(np.int64)
    reset = np.where(progress >= max_len - 1, np.int64(1), term)
    return reset, term


# ----------------------------------------------------------------------------------------
# a9: AMP observation
# ----------------------------------------------------------------------------------------
def amp_obs_step(body_pos, body_rot, body_vel, body_ang_vel, dof_pos, dof_vel, betas):
    """build_amp_observations_smpl, env/tasks/humanoid_amp.py:917-971 with local_root_obs=True,
    root_height_obs=False, has_dof_subset=True, has_shape_obs_disc=True -> [N,206]"""
    root_pos, root_rot = body_pos[:, 0], body_rot[:, 0]
    hinv = calc_heading_quat_inv(root_rot)
    root_rot_obs = quat_to_tan_norm(quat_mul(hinv, root_rot))
    lv = my_quat_rotate(hinv, body_vel[:, 0])
    la = my_quat_rotate(hinv, body_ang_vel[:, 0])
    key = body_pos[:, KEY_BODIES] - root_pos[:, None]
    N, K, _ = key.shape
    hexp = np.repeat(hinv[:, None], K, 1).reshape(-1, 4)
    key = my_quat_rotate(hexp, key.reshape(-1, 3)).reshape(N, K * 3)
    dp = dof_pos[:, DOF_SUBSET]
    dv = dof_vel[:, DOF_SUBSET]
    dof_obs = quat_to_tan_norm(exp_map_to_quat(dp.reshape(-1, 3))).reshape(N, -1)  # humanoid.py:1327-1338
    return np.concatenate([root_rot_obs, lv, la, dof_obs, dv, key, betas[:, :-6]], axis=-1).astype(F)


def amp_hist_update(amp_buf, new_step):
    """_update_hist_amp_obs + _compute_amp_observations, humanoid_amp.py:585-657:
    hist[k+1] = old[k]; slot 0 = new step.  amp_buf [N,15,206]"""
    out = np.empty_like(amp_buf)
    out[:, 1:] = amp_buf[:, :-1]
    out[:, 0] = new_step
    return out


# ----------------------------------------------------------------------------------------
# whole post-physics step (a3-a9) as the fused kernel produces it
# ----------------------------------------------------------------------------------------
def post_physics_step(rb_state, dof_state, contact, dof_force, progress, verts, betas,
                      height_samples, amp_buf):
    """BaseTask.step tail, env/tasks/base_task.py:258-265 -> post_physics_step
    (humanoid_amp.py:139-157 -> humanoid.py:1211-1232).  `progress` is the value AFTER the
    `progress_buf += 1` of humanoid.py:1213."""
    N = rb_state.shape[0]
    rb = rb_state.reshape(N, NB, 13)
    bp, br, bv, bw = rb[..., 0:3], rb[..., 3:7], rb[..., 7:10], rb[..., 10:13]
    ds = dof_state.reshape(N, ND, 2)
    dpos, dvel = ds[..., 0], ds[..., 1]
    so = self_obs(bp, br, bv, bw, betas)
    to = task_obs(bp, br, verts, progress, height_samples)
    obs = np.concatenate([so, to], axis=-1)
    fso = self_obs(*flip_state(bp, br, bv, bw), betas)
    flip = np.concatenate([fso, flip_task_obs(to)], axis=-1)
    tar = calc_pos(verts, np.arange(N), progress_time(progress))
    rew, rew_raw = reward(bp[:, 0], tar, dof_force, dvel)
    reset, term = compute_reset(progress, contact.reshape(N, NB, 3), bp, tar)
    amp = amp_hist_update(amp_buf, amp_obs_step(bp, br, bv, bw, dpos, dvel, betas))
    return dict(obs=obs.astype(F), flip_obs=flip.astype(F), rew=rew, reward_raw=rew_raw,
                reset=reset, terminate=term, amp_obs=amp.reshape(N, -1))


# ----------------------------------------------------------------------------------------
# a2: action -> PD target
# ----------------------------------------------------------------------------------------
def pd_action_offset_scale(limit_lo, limit_hi):
    """_build_pd_action_offset_scale, env/tasks/humanoid.py:950-1025 (bias_offset False,
    3-dof joints; knee-y scale forced to 5 at :1009-1013)."""
    lo, hi = limit_lo.copy(), limit_hi.copy()
    for j in range(ND // 3):
        s = slice(3 * j, 3 * j + 3)
        sc = max(np.max(np.abs(lo[s])), np.max(np.abs(hi[s])))
        sc = min(1.2 * sc, np.pi)
        lo[s], hi[s] = -sc, sc
    offset = (0.5 * (hi + lo)).astype(F)
    scale = (0.5 * (hi - lo)).astype(F)
    scale[1 * 3 + 1] = 5   # L_Knee is joint 1 -> dof 4
    scale[5 * 3 + 1] = 5   # R_Knee is joint 5 -> dof 16
    return offset, scale


def action_to_pd_targets(actions, offset, scale):
    """pre_physics_step, env/tasks/humanoid.py:1184-1202: hands and toes forced to 0."""
    tar = (offset + scale * actions).astype(F)
    for j in (3, 7, 17, 22):
        tar[:, 3 * j:3 * j + 3] = 0
    return tar


# ----------------------------------------------------------------------------------------
# a11-a13: nets (weights passed as dicts of numpy arrays; W is [out,in] like nn.Linear)
# ----------------------------------------------------------------------------------------
def rms_normalize(x, mean, var, eps=1e-5):
    """RunningMeanStd.forward eval branch, utils/running_mean_std.py:60-84 (float64 stats
    cast to float32)."""
    y = (x - mean.astype(F)) / np.sqrt(var.astype(F) + F(eps))
    return np.clip(y, F(-5.0), F(5.0)).astype(F)


_LINEAR = {"backend": "numpy", "cache": {}}


def set_linear_backend(name, threads=None):
    """'numpy' (default, the checker) or 'torch': the dense layers through torch's CPU `F.linear` (MKL / oneDNN sgemm on all
    host threads) - what the reference's own `nn.Linear` modules run when the agent sits on the CPU.
    Used by bench.py's CPU arm only; same arithmetic (fp32 GEMM + bias + ReLU), different BLAS."""
    assert name in ("numpy", "torch")
    _LINEAR["backend"] = name
    if name == "torch":
        import torch
        if threads:
            torch.set_num_threads(int(threads))


def _mlp(x, layers, relu_last=True):
    if _LINEAR["backend"] == "torch":
        import torch
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=F))
        with torch.no_grad():
            for i, (W, b) in enumerate(layers):
                key = (id(W), id(b))
                wb = _LINEAR["cache"].get(key)
                if wb is None or wb[2] is not W:
                    wb = _LINEAR["cache"][key] = (torch.from_numpy(np.ascontiguousarray(W, dtype=F)), torch.from_numpy(np.ascontiguousarray(b, dtype=F)), W)
                t = torch.nn.functional.linear(t, wb[0], wb[1])
                if relu_last or i < len(layers) - 1:
                    t = torch.relu_(t)
        return t.numpy()
    for i, (W, b) in enumerate(layers):
        x = x @ W.T.astype(F) + b.astype(F)
        if relu_last or i < len(layers) - 1:
            x = np.maximum(x, F(0))
    return x.astype(F)


def policy_forward(obs, P, noise=None):
    """ModelAMPContinuousSeptValue forward in eval mode: learning/amp_network_sept_builder.py:69-111,
    amp_network_sept_value_builder.py:31-46, rl_games ModelA2CContinuousLogStd (neglogp).
    P: dict with 'mean','var', 'task'[(W,b)x2], 'actor'[(W,b)x2], 'mu'(W,b), 'sigma'[69],
    'critic'[(W,b)x2], 'value'(W,b), 'tv'[(W,b)x2], 'tv_out'(W,b)."""
    SELF = 368
    xh = rms_normalize(obs, P["mean"], P["var"])
    t = _mlp(xh[:, SELF:], P["task"])
    ain = np.concatenate([xh[:, :SELF], t], axis=-1)
    mu = _mlp(ain, P["actor"]) @ P["mu"][0].T + P["mu"][1]
    value = _mlp(ain, P["critic"]) @ P["value"][0].T + P["value"][1]
    tv = _mlp(xh[:, SELF:SELF + 30], P["tv"]) @ P["tv_out"][0].T + P["tv_out"][1]
    logstd = P["sigma"].astype(F)
    sigma = np.exp(logstd)
    out = dict(mu=mu.astype(F), sigma=np.broadcast_to(sigma, mu.shape).astype(F),
               value=value.astype(F), task_value=tv.astype(F))
    if noise is not None:
        a = (mu + sigma * noise).astype(F)
        out["actions"] = a
        out["neglogp"] = (F(0.5) * (((a - mu) / sigma) ** 2).sum(-1) + F(0.5 * np.log(2.0 * np.pi)) * a.shape[-1]
                          + logstd.sum()).astype(F)
    return out


def critic_forward(obs, P):
    """_eval_critic, learning/common_agent.py:647-655 (value un-normalisation left to caller)."""
    SELF = 368
    xh = rms_normalize(obs, P["mean"], P["var"])
    t = _mlp(xh[:, SELF:], P["task"])
    ain = np.concatenate([xh[:, :SELF], t], axis=-1)
    return (_mlp(ain, P["critic"]) @ P["value"][0].T + P["value"][1]).astype(F)


def disc_reward(amp_obs, D, scale=2.0):
    """_calc_disc_rewards, learning/amp_continuous.py:675-692; eval_disc amp_network_builder.py:81-84.
    D: 'mean','var','mlp'[(W,b)x2],'logit'(W,b)."""
    x = rms_normalize(amp_obs, D["mean"], D["var"])
    logit = _mlp(x, D["mlp"]) @ D["logit"][0].T + D["logit"][1]
    prob = F(1) / (F(1) + np.exp(-logit))
    r = -np.log(np.maximum(F(1) - prob, F(0.0001)))
    return (r * F(scale)).astype(F), logit.astype(F)


# ----------------------------------------------------------------------------------------
# a12-a14: per-step reward / value plumbing and the LocoVal-target bookkeeping
# ----------------------------------------------------------------------------------------
def value_unnormalize(v, mean, var, eps=1e-5):
    """RunningMeanStd.forward(unnorm=True), utils/running_mean_std.py:78-80: sqrt(var + eps) * clamp(v, +-5) + mean."""
    return (np.sqrt(F(var) + F(eps)) * np.clip(v, F(-5), F(5)) + F(mean)).astype(F)


def rollout_record(st, rew, reset, terminate, next_value_raw, logit, inverted=None, value_mean=0.0, value_var=1.0,
                   inv_penalty=0.3, disc_scale=2.0, gamma=0.99, step_to_pred=144):
    """The body of the horizon loop of AMPValueAgent.play_steps after env_step (learning/amp_continuous_value.py:61-118).
    st: dict of float32 [N] arrays current_rewards, current_lengths, current_combined, discount, game_combined,
    terminated_flags (updated in place).  rew [N] task reward of the env step, reset / terminate int64 [N], next_value_raw [N]
    critic(next obs) before un-normalisation, logit [N] discriminator logit, inverted bool [N] or None.
    Returns the rows written to the experience buffer: rewards, dones, next_values, amp_rewards (all [N])."""
    r = rew.astype(F).copy()
    if inverted is not None:
        r[inverted.astype(bool)] *= F(-inv_penalty)                                            # :62-64
    shaped = r * F(1.0)                                                                        # rewards_shaper, scale_value 1
    term = terminate.astype(F)
    st["terminated_flags"] += term                                                             # :76-77
    next_values = value_unnormalize(next_value_raw, value_mean, value_var) * (F(1) - term)     # :87-89
    prob = F(1) / (F(1) + np.exp(-logit.astype(F)))
    amp_r = (-np.log(np.maximum(F(1) - prob, F(0.0001))) * F(disc_scale)).astype(F)            # amp_continuous.py:675-692
    done = (reset != 0).astype(F)
    nd = F(1) - done
    st["current_rewards"] = st["current_rewards"] + r                                          # :94
    st["current_lengths"] = st["current_lengths"] + F(1)
    st["current_combined"] = st["current_combined"] + (shaped + amp_r) * st["discount"]        # :96-97
    ln = st["current_lengths"]
    sel = ((ln <= step_to_pred) & (done != 0)) | ((ln == step_to_pred) & (nd != 0))            # :106-107
    st["game_combined"] = st["game_combined"] + st["current_combined"] * sel.astype(F)         # :109
    st["current_combined"] = st["current_combined"] * nd                                       # :113
    st["discount"] = np.where(done != 0, F(1), st["discount"] * F(gamma)).astype(F)            # :114-116
    st["current_rewards"] = st["current_rewards"] * nd                                         # :117-118
    st["current_lengths"] = st["current_lengths"] * nd
    return dict(rewards=shaped, dones=done, next_values=next_values, amp_rewards=amp_r)


def player_record(st, rew, rew_raw, reset, logit, locoval_scores, inverted=None, plot_val_reward=True, inv_penalty=0.3, disc_scale=2.0,
                  gamma=0.99, step_to_pred=144, min_reward=-10.0, max_reward=100.0):
    """One iteration of the step loop of AMPPlayerContinuousValue.run (learning/amp_value_players.py:123-198), vectorised over
    envs (the reference plays one env).  st: dict of float32 [N] arrays n (steps played in the running episode), cr, coef, pred,
    cr_to_pred, c_loc, c_pow, c_disc and their snapshots loc_to_pred, pow_to_pred, disc_to_pred (updated in place).  rew [N] task reward, rew_raw [N,2] (location, power), reset int64 [N],
    logit [N] discriminator logit, locoval_scores [N].  -> finished episodes as dict(env, pred, cr_to_pred, norm_reward, sq_err,
    c_loc, c_pow, c_disc, steps)."""
    n = st["n"]
    first = n == 0
    st["pred"] = np.where(first, locoval_scores.astype(F), st["pred"])                             # :128-137 score at the episode's first step
    prob = F(1) / (F(1) + np.exp(-logit.astype(F)))
    disc = (-np.log(np.maximum(F(1) - prob, F(0.0001))) * F(disc_scale)).astype(F)
    st["coef"] = (st["coef"] * F(gamma)).astype(F)                                                 # :144, before use
    c = st["coef"]
    if plot_val_reward:                                                                            # :145-160
        r_loc, r_pow = rew_raw[:, 0].astype(F), rew_raw[:, 1].astype(F)
        st["c_disc"] = st["c_disc"] + disc * F(0.25) * c
        st["c_loc"] = st["c_loc"] + r_loc * F(0.5) * c
        st["c_pow"] = st["c_pow"] + r_pow * F(0.5) * c
        st["cr"] = st["cr"] + ((r_loc + r_pow) * F(0.5) + disc * F(0.25)) * c
    else:                                                                                          # :161-163, with the penalty of :127
        r = rew.astype(F).copy()
        if inverted is not None:
            r[inverted.astype(bool)] *= F(-inv_penalty)
        st["cr"] = st["cr"] + r * c
    at_pred = n == step_to_pred                                                                    # :177-179
    done = reset != 0
    early = done & (n < step_to_pred)                                                              # :186-188
    snap = at_pred | early
    st["cr_to_pred"] = np.where(snap, st["cr"], st["cr_to_pred"]).astype(F)
    for a, b in (("loc_to_pred", "c_loc"), ("pow_to_pred", "c_pow"), ("disc_to_pred", "c_disc")):  # rewards_loc / _pow / _disc (:180-193)
        st[a] = np.where(snap, st[b], st[a]).astype(F)
    ids = np.nonzero(done)[0]
    norm = ((st["cr_to_pred"][ids] - F(min_reward)) / F(max_reward - min_reward)).astype(F)        # :195
    out = dict(env=ids, pred=st["pred"][ids].copy(), cr_to_pred=st["cr_to_pred"][ids].copy(), norm_reward=norm,
               sq_err=((st["pred"][ids] - norm) ** 2).astype(F), c_loc=st["loc_to_pred"][ids].copy(), c_pow=st["pow_to_pred"][ids].copy(),
               c_disc=st["disc_to_pred"][ids].copy(), steps=(n[ids] + 1).copy())
    for k in ("cr", "c_loc", "c_pow", "c_disc"):                                                   # :214-217 (new game: counters restart)
        st[k] = np.where(done, F(0), st[k]).astype(F)
    st["coef"] = np.where(done, F(1), st["coef"]).astype(F)
    st["n"] = np.where(done, F(0), n + 1).astype(F)
    return out


# ----------------------------------------------------------------------------------------
# f2: mocap reset state and AMP demo observations (utils/motion_lib_smpl.py, env/tasks/humanoid_amp.py:168-220)
# ----------------------------------------------------------------------------------------
def slerp(q0, q1, t):
    """utils/torch_utils.py:114-136 (t broadcastable to [..., 1])."""
    q0, q1 = q0.astype(F), q1.astype(F).copy()
    c = np.sum(q0 * q1, axis=-1, dtype=F)
    neg = c < 0
    q1[neg] = -q1[neg]
    c = np.abs(c)[..., None]
    half = np.arccos(np.minimum(c, F(1))).astype(F)
    s = np.sqrt(np.maximum(F(1) - c * c, F(0))).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        ra = np.sin((F(1) - t) * half) / s
        rb = np.sin(t * half) / s
        q = ra * q0 + rb * q1
    q = np.where(np.abs(s) < F(0.001), F(0.5) * q0 + F(0.5) * q1, q)
    q = np.where(np.abs(c) >= 1, q0, q)
    return q.astype(F)


def quat_to_exp_map(q):
    """utils/torch_utils.py:27-65: angle-axis with angle wrapped to (-pi, pi], default axis z below 1e-5."""
    q = q.astype(F)
    sin_t = np.sqrt(np.maximum(F(1) - q[..., 3] * q[..., 3], F(0))).astype(F)
    ang = normalize_angle((F(2) * np.arccos(np.clip(q[..., 3], F(-1), F(1)))).astype(F))
    with np.errstate(divide="ignore", invalid="ignore"):
        axis = q[..., :3] / sin_t[..., None]
    mask = np.abs(sin_t) > F(1e-5)
    default = np.zeros_like(axis); default[..., 2] = 1
    ang = np.where(mask, ang, F(0))
    axis = np.where(mask[..., None], axis, default)
    return (ang[..., None] * axis).astype(F)


def calc_frame_blend(time, length, num_frames, dt):
    """utils/motion_lib_smpl.py:596-606."""
    time = time.astype(F).copy()
    phase = np.clip(time / length, F(0), F(1))
    time[time < 0] = 0
    f0 = (phase * (num_frames - 1).astype(F)).astype(np.int64)
    f1 = np.minimum(f0 + 1, num_frames - 1)
    blend = (time - f0.astype(F) * dt) / dt
    return f0, f1, blend.astype(F)


def motion_state(lib, motion_ids, motion_times):
    """MotionLibSMPL.get_motion_state_smpl (utils/motion_lib_smpl.py:485-563) on the flat arrays of `lib`
    (emloco_b200.synthetic.synthetic_motion_lib layout)."""
    f0, f1, blend = calc_frame_blend(motion_times, lib["motion_lengths"][motion_ids], lib["motion_num_frames"][motion_ids], lib["motion_dt"][motion_ids])
    f0l, f1l = f0 + lib["length_starts"][motion_ids], f1 + lib["length_starts"][motion_ids]
    b = blend[:, None, None]
    lerp = lambda k: ((F(1) - b) * lib[k][f0l] + b * lib[k][f1l]).astype(F)
    rg_pos, body_vel, body_ang_vel, dof_vel = lerp("gts"), lerp("gvs"), lerp("gavs"), lerp("dvs")
    local_rot = slerp(lib["lrs"][f0l], lib["lrs"][f1l], b)
    rb_rot = slerp(lib["grs"][f0l], lib["grs"][f1l], b)
    dof_pos = quat_to_exp_map(local_rot[:, 1:]).reshape(len(motion_ids), -1)
    return dict(root_pos=rg_pos[:, 0], root_rot=rb_rot[:, 0], dof_pos=dof_pos, root_vel=body_vel[:, 0], root_ang_vel=body_ang_vel[:, 0],
                dof_vel=dof_vel.reshape(len(motion_ids), -1), key_pos=rg_pos[:, KEY_BODIES], rg_pos=rg_pos, rb_rot=rb_rot, body_vel=body_vel,
                body_ang_vel=body_ang_vel, motion_bodies=lib["motion_bodies"][motion_ids])


def amp_obs_demo(lib, motion_ids, motion_times0, dt=CONTROL_DT, steps=AMP_STEPS):
    """HumanoidAMP.build_amp_obs_demo (env/tasks/humanoid_amp.py:186-211): `steps` observations per sample at times
    t0 - k dt, newest first -> [n, steps * 206]."""
    n = len(motion_ids)
    ids = np.repeat(motion_ids, steps)
    times = (motion_times0[:, None].astype(F) + (-F(dt) * np.arange(steps, dtype=F))[None]).reshape(-1)
    m = motion_state(lib, ids, times)
    hinv = calc_heading_quat_inv(m["root_rot"])
    root_rot_obs = quat_to_tan_norm(quat_mul(hinv, m["root_rot"]))
    lv, la = my_quat_rotate(hinv, m["root_vel"]), my_quat_rotate(hinv, m["root_ang_vel"])
    key = m["key_pos"] - m["root_pos"][:, None]
    key = my_quat_rotate(np.repeat(hinv[:, None], 4, 1).reshape(-1, 4), key.reshape(-1, 3)).reshape(len(ids), 12)
    dof_obs = quat_to_tan_norm(exp_map_to_quat(m["dof_pos"][:, DOF_SUBSET].reshape(-1, 3))).reshape(len(ids), -1)
    obs = np.concatenate([root_rot_obs, lv, la, dof_obs, m["dof_vel"][:, DOF_SUBSET], key, m["motion_bodies"][:, :-6]], -1).astype(F)
    return obs.reshape(n, steps * AMP_STEP_DIM)


# ----------------------------------------------------------------------------------------
# a15: GAE
# ----------------------------------------------------------------------------------------
def discount_values(fdones, values, rewards, next_values, gamma=0.99, tau=0.95):
    """discount_values, learning/common_agent.py:573-587.  shapes [T,N,1] (fdones [T,N])."""
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = np.zeros_like(rewards[0])
    for t in reversed(range(T)):
        nd = (F(1.0) - fdones[t])[:, None]
        delta = rewards[t] + F(gamma) * next_values[t] - values[t]
        last = (delta + F(gamma) * F(tau) * nd * last).astype(F)
        adv[t] = last
    return adv


def normalize_advantages(returns, values):
    """_calc_advs, learning/common_agent.py:685-696 (torch.std is the unbiased estimator)."""
    adv = (returns - values).sum(axis=1)
    return ((adv - adv.mean(dtype=F)) / (adv.std(ddof=1, dtype=F) + F(1e-8))).astype(F)


# ----------------------------------------------------------------------------------------
# a16: LocoVal
# ----------------------------------------------------------------------------------------
def locoval_normalize(traj, pose, vel):
    """ValuePoseNet._rotate_normalization, learning/value_pose_net.py:73-103."""
    x = traj[:, 1, 0].copy(); y = traj[:, 1, 1]
    near = np.abs(x) < F(1e-10)
    x = (x * (~near) + near * F(1e-10)).astype(F)
    ang = np.arctan2(y, x).astype(F)
    c, s = np.cos(ang), np.sin(ang)
    R = np.zeros((len(ang), 2, 2), F)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1] = c, -s, s, c
    traj_r = np.einsum("bnk,bkj->bnj", traj[..., :2], R).astype(F)
    pose_r = pose.copy()
    pose_r[..., :2] = np.einsum("bnk,bkj->bnj", pose[..., :2], R)
    vel_r = np.einsum("bk,bkj->bj", vel[:, :2], R).astype(F)
    return traj_r, pose_r, vel_r, R


def locoval_forward(traj, pose, vel, W):
    """ValuePoseNet.forward (forward_full), learning/value_pose_net.py:105-149.
    W: dict fc1/fc2/fc3 -> (weight[out,in], bias).  Returns (value [B,1], mutated pose) -
    the reference rotates and zeroes the caller's init_pose in place (:97,:141-144)."""
    traj_r, pose_r, vel_r, _ = locoval_normalize(traj, pose, vel)
    pose_r[:, [4, 8]] = 0
    pose_r[:, [9, 10, 11]] = 0
    B = traj.shape[0]
    x = np.concatenate([traj_r.reshape(B, 26), pose_r.reshape(B, 72), vel_r.reshape(B, 2)], axis=-1).astype(F)
    h1 = np.maximum(x @ W["fc1"][0].T + W["fc1"][1], F(0))
    h2 = np.maximum(h1 @ W["fc2"][0].T + W["fc2"][1], F(0))
    z = h2 @ W["fc3"][0].T + W["fc3"][1]
    return (F(1) / (F(1) + np.exp(-z))).astype(F), pose_r


def locoval_finetune_step(W, opt, traj, pose, vel, game_combined, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                          r_min=-10.0, r_max=100.0):
    """The `_do_finetune` block of AMPValueAgent.play_steps (learning/amp_continuous_value.py:122-146) with
    torch.optim.AdamW(lr=1e-3, weight_decay=1e-4) and MSELoss(reduction='sum') (learning/common_agent.py:94-96), backward
    written out by hand in float64.  W: dict fc1/fc2/fc3 -> [weight, bias] (updated in place); opt: dict(step, m, v) with m / v
    shaped like W.  Returns (loss, sum pred, sum target, count); zeroes game_combined like :145."""
    valid = np.nonzero(game_combined)[0]
    if len(valid) == 0:
        return 0.0, 0.0, 0.0, 0
    traj_r, pose_r, vel_r, _ = locoval_normalize(traj[valid][..., :2], pose[valid], vel[valid])
    pose_r[:, [4, 8]] = 0
    pose_r[:, [9, 10, 11]] = 0
    B = len(valid)
    d = np.float64
    x = np.concatenate([traj_r.reshape(B, 26), pose_r.reshape(B, 72), vel_r.reshape(B, 2)], axis=-1).astype(d)
    (w1, b1), (w2, b2), (w3, b3) = [[a.astype(d) for a in W[k]] for k in ("fc1", "fc2", "fc3")]
    h1 = np.maximum(x @ w1.T + b1, 0); h2 = np.maximum(h1 @ w2.T + b2, 0)
    v = 1 / (1 + np.exp(-(h2 @ w3.T + b3)))[:, 0]
    target = (game_combined[valid].astype(d) - r_min) / (r_max - r_min)
    diff = v - target
    dz = (2 * diff * v * (1 - v))[:, None]
    g3w, g3b = dz.T @ h2, dz.sum(0)
    dh2 = (dz @ w3) * (h2 > 0)
    g2w, g2b = dh2.T @ h1, dh2.sum(0)
    dh1 = (dh2 @ w2) * (h1 > 0)
    g1w, g1b = dh1.T @ x, dh1.sum(0)
    opt["step"] += 1
    t = opt["step"]
    bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
    for k, gs in (("fc1", (g1w, g1b)), ("fc2", (g2w, g2b)), ("fc3", (g3w, g3b))):
        for i, g in enumerate(gs):
            p = W[k][i].astype(d) * (1 - lr * weight_decay)
            m = opt["m"][k][i] = betas[0] * opt["m"][k][i] + (1 - betas[0]) * g
            vv = opt["v"][k][i] = betas[1] * opt["v"][k][i] + (1 - betas[1]) * g * g
            W[k][i] = (p - (lr / bc1) * m / (np.sqrt(vv) / np.sqrt(bc2) + eps)).astype(F)
    game_combined[:] = 0
    return float((diff ** 2).sum()), float(v.sum()), float(target.sum()), B


def locoval_loss(value):
    """calc_embodied_motion_loss, learning/value_pose_net.py:151-159: MSE(value, 1)."""
    return np.mean((value - F(1)) ** 2, dtype=F)


def plausibl_mlp(x, W):
    """plausibl/test_value_mlp.py:24-113 MLP.forward: 24->12->6->1, no sigmoid."""
    h1 = np.maximum(x @ W["fc1"][0].T + W["fc1"][1], F(0))
    h2 = np.maximum(h1 @ W["fc2"][0].T + W["fc2"][1], F(0))
    return (h2 @ W["fc3"][0].T + W["fc3"][1]).astype(F)
