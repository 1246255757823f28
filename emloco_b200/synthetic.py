"""Seeded synthetic inputs for the benchmark and smoke runs: there is no network for the AMASS / JTA /
JRDB data the reference resets from, so initial states and trajectories are generated here with the same
distributions the reference samples from.  Host-side numpy, run once per reset batch - not on the hot path.

  * trajectories: the random-walk polylines of TrajGenerator.reset (pacer/pacer/env/util/traj_generator.py:60-113):
    101 vertices, per-vertex turn U(-1,1)*dtheta_max*dt with 2 % sharp turns U(-pi,pi), speed random walk with
    |accel| <= 2 m/s^2 clipped to [0.0005, 3] m/s (pacer.yaml:55-61), initial heading U(-pi,pi);
  * initial state: upright rest pose at the rest height, root xy uniform in the 8 x 8 m patch behind the 50 m border of the
    default terrain (humanoid_pedestrain_terrain.py:1142-1165), heading = first trajectory segment
    (--random_heading), root speed U(1,1.5) m/s along the heading (..terrain.py:568);
  * LocoVal batches: 13 waypoints at 0.4 s spacing sampled from such polylines, origin-relative; pose = rest-pose
    joint positions + N(0,0.05); vel = (w1-w0)*2.5  (social-transmotion/load_jta_traj.py:66-120 shapes).
"""
from __future__ import annotations

import numpy as np

EPISODE_LEN, NUM_VERTS, CONTROL_DT = 168, 101, 2.0 / 60.0
DTHETA_MAX, SPEED_MIN, SPEED_MAX, ACCEL_MAX, SHARP_PROB = 2.0, 0.0005, 3.0, 2.0, 0.02


def random_walk_verts(n, init_xy, rng, num_verts=NUM_VERTS, episode_dur=EPISODE_LEN * CONTROL_DT):
    dt = episode_dur / (num_verts - 1)
    dtheta = (2 * rng.random((n, num_verts - 1)) - 1.0) * DTHETA_MAX * dt
    sharp = np.pi * (2 * rng.random((n, num_verts - 1)) - 1.0)
    mask = rng.random((n, num_verts - 1)) < SHARP_PROB
    dtheta[mask] = sharp[mask]
    dtheta[:, 0] = np.pi * (2 * rng.random(n) - 1.0)
    dspeed = (2 * rng.random((n, num_verts - 1)) - 1.0) * ACCEL_MAX * dt
    dspeed[:, 0] = (SPEED_MAX - SPEED_MIN) * rng.random(n) + SPEED_MIN
    speed = np.zeros_like(dspeed)
    speed[:, 0] = dspeed[:, 0]
    for i in range(1, num_verts - 1):
        speed[:, i] = np.clip(speed[:, i - 1] + dspeed[:, i], SPEED_MIN, SPEED_MAX)
    theta = np.cumsum(dtheta, -1)
    dpos = np.stack([np.cos(theta), -np.sin(theta), np.zeros_like(theta)], -1) * (speed * dt)[..., None]
    dpos[:, 0, 0:2] += init_xy
    verts = np.zeros((n, num_verts, 3), np.float32)
    verts[:, 0, 0:2] = init_xy
    verts[:, 1:] = np.cumsum(dpos, -2)
    return verts, -theta[:, 0]


def synthetic_env_state(n, seed=0, root_height=0.93, border=50.0, patch=8.0):
    """-> dict(root [n,13] f32, dof [n*69,2] f32, verts [n,101,3] f32)."""
    rng = np.random.default_rng(seed)
    xy = border + patch * rng.random((n, 2))
    verts, heading = random_walk_verts(n, xy, rng)
    root = np.zeros((n, 13), np.float32)
    root[:, 0:2] = xy
    root[:, 2] = root_height
    root[:, 5] = np.sin(heading / 2)
    root[:, 6] = np.cos(heading / 2)
    speed = rng.uniform(1.0, 1.5, n)
    root[:, 7] = speed * np.cos(heading)
    root[:, 8] = speed * np.sin(heading)
    dof = np.zeros((n * 69, 2), np.float32)
    return dict(root=root, dof=dof, verts=verts, waypoints=waypoints_from_verts(verts))


def synthetic_traj_pool(p, seed=0):
    """Stand-in for the saved JTA / JRDB trajectory pickles (`{id: {'traj': [101,3]}}`, traj_generator.py:44-52): p
    origin-relative random-walk polylines, float32 [p,101,3]."""
    rng = np.random.default_rng(seed + 7919)
    verts, _ = random_walk_verts(p, np.zeros((p, 2)), rng)
    return np.ascontiguousarray(verts, np.float32)


def waypoints_from_verts(verts, num=13, sample_dt=0.4):
    """_fetch_traj_samples at progress 0 (humanoid_traj.py:208-224 via TrajGenerator.calc_pos :278-296), xy only,
    relative to the first waypoint (vec_task_wrappers.py:47-52).  Note calc_pos spreads the 101 vertices over
    num_verts * dt (traj_generator.py:269-272)."""
    nv = verts.shape[1]
    dur = nv * (EPISODE_LEN * CONTROL_DT / (nv - 1))
    seg = np.clip(np.arange(num) * sample_dt / dur, 0, 1) * (nv - 1)
    i0 = np.floor(seg).astype(int); i1 = np.ceil(seg).astype(int); fr = (seg - i0).astype(np.float32)
    w = verts[:, i0, :2] * (1 - fr)[None, :, None] + verts[:, i1, :2] * fr[None, :, None]
    return np.ascontiguousarray(w - w[:, :1], dtype=np.float32)     # fancy indexing above leaves a permuted memory layout


def synthetic_locoval_batch(b, seed=0, rest_joint_pos=None):
    """JTA-shaped LocoVal inputs: traj [b,13,2], pose [b,24,3], vel [b,2] (float32)."""
    rng = np.random.default_rng(seed)
    verts, _ = random_walk_verts(b, np.zeros((b, 2)), rng)
    traj = waypoints_from_verts(verts)
    if rest_joint_pos is None:
        from .model import build_model_arrays, rest_joint_positions
        rest_joint_pos = rest_joint_positions(build_model_arrays())
    pose = (np.asarray(rest_joint_pos, np.float32)[None] + rng.normal(0, 0.05, (b, 24, 3))).astype(np.float32)
    pose -= pose[:, :1].copy()
    vel = ((traj[:, 1] - traj[:, 0]) * 2.5).astype(np.float32)
    return traj, pose, vel


# ---- synthetic motion library (the AMASS clips of utils/motion_lib_smpl.py are not redistributable) ----------
def _q_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _q_rot(q, v):
    qv = q[..., :3]
    t = 2.0 * np.cross(qv, v)
    return v + q[..., 3:4] * t + np.cross(qv, t)


def _q_exp(e):
    ang = np.linalg.norm(e, axis=-1, keepdims=True)
    ax = np.where(ang > 1e-8, e / np.maximum(ang, 1e-8), np.array([0.0, 0.0, 1.0]))
    return np.concatenate([ax * np.sin(ang / 2), np.cos(ang / 2)], -1)


def synthetic_motion_lib(num_motions=16, seed=0, parent=None, offset=None, fps=30.0):
    """Walking-like clips in the layout MotionLibSMPL keeps on the device (utils/motion_lib_smpl.py:248-330): per-frame global
    translations gts [F,24,3], global / local rotations grs / lrs [F,24,4] (xyzw), global linear / angular velocities gvs / gavs
    [F,24,3], joint velocities dvs [F,23,3]; per motion: length (s), frame count, dt, first-frame index, shape parameters [17].
    Joint angles are sums of sinusoids, the root walks along its heading; velocities are finite differences, like poselib's."""
    from .model import build_model_arrays
    if parent is None:
        A = build_model_arrays()
        parent, offset = A["parent"], A["offset"]
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("gts", "grs", "lrs", "gvs", "gavs", "dvs")}
    lens, nfr, starts = [], [], []
    dt = 1.0 / fps
    start = 0
    for m in range(num_motions):
        Fm = int(rng.integers(45, 150))
        t = np.arange(Fm)[:, None, None] * dt
        amp = rng.uniform(0.05, 0.6, (1, 23, 3)); freq = rng.uniform(0.5, 2.0, (1, 23, 3)); ph = rng.uniform(0, 2 * np.pi, (1, 23, 3))
        e = amp * np.sin(2 * np.pi * freq * t + ph)                                  # joint exp-maps [F,23,3]
        yaw0, yawr = rng.uniform(-np.pi, np.pi), rng.uniform(-0.5, 0.5)
        yaw = yaw0 + yawr * t[:, 0, 0]
        tilt = 0.08 * np.sin(2 * np.pi * 1.1 * t[:, 0, :] + rng.uniform(0, 6, (1, 1)))  # [F,1]
        rq = _q_mul(np.stack([0 * yaw, 0 * yaw, np.sin(yaw / 2), np.cos(yaw / 2)], -1), _q_exp(np.concatenate([tilt, tilt * 0.5, 0 * tilt], -1)))
        lrs = np.concatenate([rq[:, None], _q_exp(e)], 1)                            # [F,24,4]
        speed = rng.uniform(0.6, 1.6)
        vel = speed * np.stack([np.cos(yaw), np.sin(yaw), 0 * yaw], -1)
        rp = np.cumsum(vel * dt, 0) + np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), 0.0])
        rp[:, 2] = 0.9 + 0.02 * np.sin(2 * np.pi * 2.0 * t[:, 0, 0])
        grs, gts = np.zeros((Fm, 24, 4)), np.zeros((Fm, 24, 3))
        grs[:, 0], gts[:, 0] = lrs[:, 0], rp
        for b in range(1, 24):
            p = int(parent[b])
            grs[:, b] = _q_mul(grs[:, p], lrs[:, b])
            gts[:, b] = gts[:, p] + _q_rot(grs[:, p], np.broadcast_to(offset[b], (Fm, 3)))
        grs /= np.linalg.norm(grs, axis=-1, keepdims=True)
        gvs = np.gradient(gts, dt, axis=0)

        def ang_vel(q):                                                              # from consecutive orientations
            dq = _q_mul(q[1:], q[:-1] * np.array([-1, -1, -1, 1.0]))
            dq *= np.sign(dq[..., 3:4] + 1e-12)
            ang = 2 * np.arctan2(np.linalg.norm(dq[..., :3], axis=-1, keepdims=True), dq[..., 3:4])
            ax = dq[..., :3] / np.maximum(np.linalg.norm(dq[..., :3], axis=-1, keepdims=True), 1e-9)
            w = ax * ang / dt
            return np.concatenate([w, w[-1:]], 0)
        gavs = ang_vel(grs)
        dvs = ang_vel(lrs[:, 1:])
        for k, v in (("gts", gts), ("grs", grs), ("lrs", lrs), ("gvs", gvs), ("gavs", gavs), ("dvs", dvs)):
            out[k].append(v.astype(np.float32))
        lens.append(dt * (Fm - 1)); nfr.append(Fm); starts.append(start)
        start += Fm
    lib = {k: np.concatenate(v, 0) for k, v in out.items()}
    lib.update(motion_lengths=np.array(lens, np.float32), motion_num_frames=np.array(nfr, np.int32), motion_dt=np.full(num_motions, dt, np.float32),
               length_starts=np.array(starts, np.int32),
               motion_bodies=np.concatenate([rng.integers(0, 2, (num_motions, 1)), rng.normal(0, 1, (num_motions, 16))], 1).astype(np.float32),
               weights=np.full(num_motions, 1.0 / num_motions, np.float32))
    return lib
