"""Flattened articulation model handed to the C ABI (``emloco_model`` in include/emloco.h).

Mirrors what the reference derives at env-creation time:
* PD gain scaling  ``stiffness *= humanoid_mass / 77 * kp_scale`` - humanoid.py:905-910
* action -> PD target offset/scale, knee-y scale forced to 5 - humanoid.py:950-1025
"""
from __future__ import annotations

import numpy as np

from .mjcf import HumanoidModel, default_model

NB, ND = 24, 69


def pd_action_offset_scale(limit_lo, limit_hi, names):
    """_build_pd_action_offset_scale (humanoid.py:950-1025), bias_offset False, 3-dof joints."""
    lo, hi = np.array(limit_lo, np.float64), np.array(limit_hi, np.float64)
    for j in range(len(lo) // 3):
        s = slice(3 * j, 3 * j + 3)
        sc = max(np.max(np.abs(lo[s])), np.max(np.abs(hi[s])))
        sc = min(1.2 * sc, np.pi)
        lo[s], hi[s] = -sc, sc
    offset = (0.5 * (hi + lo)).astype(np.float32)
    scale = (0.5 * (hi - lo)).astype(np.float32)
    dof_names = names[1:]
    scale[dof_names.index("L_Knee") * 3 + 1] = 5     # humanoid.py:1009-1013
    scale[dof_names.index("R_Knee") * 3 + 1] = 5
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
