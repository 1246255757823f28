"""Flattened articulation model handed to the C ABI (``emloco_model`` in include/emloco.h).

Mirrors what the reference derives at env-creation time:
* PD gain scaling  ``stiffness *= humanoid_mass / 77 * kp_scale`` - humanoid.py:905-910
* action -> PD target offset/scale, knee-y scale forced to 5 - humanoid.py:950-1025
"""
from __future__ import annotations

import numpy as np

from .mjcf import HumanoidModel, default_model

NB, ND = 24, 69


def pd_action_offset_scale(limit_lo, limit_hi, names, smpl_pd_offset=False):
    """_build_pd_action_offset_scale (humanoid.py:950-1025), bias_offset False, 3-dof joints; float32 like the reference (Isaac Gym
    reports the DOF limits as float32).  smpl_pd_offset: cfg env.has_smpl_pd_offset with has_upright_start (shoulder x offsets
    -+pi/2, :1015-1018; default cfg: False).  Pinned to the reference function by tests/golden/pd_table.npz."""
    lo, hi = np.array(limit_lo, np.float32), np.array(limit_hi, np.float32)
    for j in range(len(lo) // 3):
        s = slice(3 * j, 3 * j + 3)
        sc = max(np.max(np.abs(lo[s])), np.max(np.abs(hi[s])))
        sc = min(1.2 * sc, np.pi)
        lo[s], hi[s] = -sc, sc
    offset = (0.5 * (hi + lo)).astype(np.float32)
    scale = (0.5 * (hi - lo)).astype(np.float32)
    dof_names = list(names[1:])
    scale[dof_names.index("L_Knee") * 3 + 1] = 5     # humanoid.py:1009-1013
    scale[dof_names.index("R_Knee") * 3 + 1] = 5
    if smpl_pd_offset:
        offset[dof_names.index("L_Shoulder") * 3] = -np.pi / 2
        offset[dof_names.index("R_Shoulder") * 3] = np.pi / 2
    return offset, scale


def build_model_arrays(mdl: HumanoidModel | None = None, kp_scale=1.0, kd_scale=None, default_mass=77.0,
                       scale_by_mass=True):
    mdl = default_model() if mdl is None else mdl
    assert mdl.num_bodies == NB and mdl.num_dof == ND, "kernel is specialised for the 24-body SMPL humanoid"
    kd_scale = kp_scale if kd_scale is None else kd_scale
    pd_scale = (mdl.total_mass / default_mass) if scale_by_mass else 1.0
    kp = mdl.kp * pd_scale * kp_scale
    kd = mdl.kd * pd_scale * kd_scale
    # the spherical-joint formulation needs isotropic gains per joint (true for smpl_humanoid.xml)
    for j in range(ND // 3):
        for a in (kp, kd, mdl.armature):
            assert np.allclose(a[3 * j:3 * j + 3], a[3 * j]), "per-joint gains must be equal on x/y/z"
    I = mdl.inertia
    inertia6 = np.stack([I[:, 0, 0], I[:, 0, 1], I[:, 0, 2], I[:, 1, 1], I[:, 1, 2], I[:, 2, 2]], -1)
    off, sc = pd_action_offset_scale(mdl.limit_lo, mdl.limit_hi, mdl.names)
    per_joint = lambda a: np.concatenate([[0.0], a[0::3]])
    return dict(
        names=mdl.names, parent=mdl.parent.astype(np.int32), offset=mdl.offset, mass=mdl.mass, com=mdl.com,
        inertia6=inertia6, kp=kp, kd=kd, armature=mdl.armature,
        kp_joint=per_joint(kp), kd_joint=per_joint(kd), arm_joint=per_joint(mdl.armature),
        geom_type=mdl.geom_type.astype(np.int32), geom_a=mdl.geom_a, geom_b=mdl.geom_b, geom_r=mdl.geom_r,
        pd_offset=off, pd_scale=sc, total_mass=mdl.total_mass)


def rest_root_height(arrs):
    """Lowest collision point of the rest pose relative to the pelvis origin -> root z that puts the feet on z=0."""
    NBn = len(arrs["parent"])
    x = np.zeros((NBn, 3))
    for i in range(1, NBn):
        x[i] = x[arrs["parent"][i]] + arrs["offset"][i]
    low = 0.0
    for i in range(NBn):
        t = arrs["geom_type"][i]
        if t == 0:
            z = x[i, 2] + arrs["geom_a"][i][2] - arrs["geom_r"][i]
        elif t == 1:
            z = x[i, 2] + min(arrs["geom_a"][i][2], arrs["geom_b"][i][2]) - arrs["geom_r"][i]
        else:
            z = x[i, 2] + arrs["geom_a"][i][2] - arrs["geom_b"][i][2]
        low = min(low, z)
    return -low


def rest_joint_positions(arrs):
    """Joint (body-origin) positions of the rest pose relative to the pelvis: [24,3]."""
    NBn = len(arrs["parent"])
    x = np.zeros((NBn, 3), np.float32)
    for i in range(1, NBn):
        x[i] = x[arrs["parent"][i]] + arrs["offset"][i]
    return x


EM_FLOATS = 576


def geom_bound(arrs):
    """Largest distance from each body origin to one of its contact points, + the primitive's radius (the bound the physics
    kernel uses to skip the contact loop of bodies that cannot reach the ground; same formula as emloco_create)."""
    out = np.zeros(NB, np.float32)
    for i in range(NB):
        a, b, r, t = np.asarray(arrs["geom_a"][i], np.float64), np.asarray(arrs["geom_b"][i], np.float64), float(arrs["geom_r"][i]), int(arrs["geom_type"][i])
        if t == 0:
            ext = np.linalg.norm(a) + r
        elif t == 1:
            ext = max(np.linalg.norm(a), np.linalg.norm(b)) + r
        else:
            ext = np.linalg.norm(np.abs(a) + np.abs(b))
        out[i] = ext * 1.0001 + 1e-5
    return out


def pack_env_model(arrs):
    """One env's body model as the 576 floats of emloco_set_env_models (include/emloco.h)."""
    f = lambda a: np.asarray(a, np.float32).reshape(-1)
    v = np.concatenate([f(arrs["offset"]), f(arrs["mass"]), f(arrs["com"]), f(arrs["inertia6"]), f(arrs["kp_joint"]), f(arrs["kd_joint"]),
                        f(arrs["arm_joint"]), f(arrs["geom_a"]), f(arrs["geom_b"]), f(arrs["geom_r"]), geom_bound(arrs)])
    assert v.size == EM_FLOATS
    return v


def scaled_model_arrays(arrs, scale, default_mass=77.0):
    """A body of the same proportions `scale` times as tall: lengths x s, masses x s^3, inertias x s^5, and the PD gains of the
    reference's mass rule `stiffness *= humanoid_mass / 77 * kp_scale` (humanoid.py:905-910) re-applied to the new mass.
    Stand-in for the per-beta bodies `Robot.load_from_skeleton` generates (uhc/smpllib/smpl_local_robot.py:1235 - needs the
    licensed SMPL model files): exercises per-env models with physically consistent numbers."""
    s = float(scale)
    out = dict(arrs)
    for k in ("offset", "com", "geom_a", "geom_b", "geom_r"):
        out[k] = np.asarray(arrs[k], np.float64) * s
    out["mass"] = np.asarray(arrs["mass"], np.float64) * s ** 3
    out["inertia6"] = np.asarray(arrs["inertia6"], np.float64) * s ** 5
    out["total_mass"] = float(arrs["total_mass"]) * s ** 3
    gain = s ** 3                                               # (new mass / 77) / (old mass / 77)
    for k in ("kp", "kd", "kp_joint", "kd_joint"):
        out[k] = np.asarray(arrs[k], np.float64) * gain
    return out
