"""TEST INFRASTRUCTURE ONLY - ctypes wrapper over oracle/physics_oracle.c (fp64 CPU restatement of
the articulated-body step).  See the header of physics_oracle.c for scope and parity status
("parity unpinned" against PhysX).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libphysics_oracle.so")
NB, ND = 24, 69


class OModel(C.Structure):
    _fields_ = [("parent", C.c_int * NB), ("offset", C.c_double * 3 * NB), ("mass", C.c_double * NB),
                ("com", C.c_double * 3 * NB), ("inertia", C.c_double * 6 * NB),
                ("kp", C.c_double * NB), ("kd", C.c_double * NB), ("arm", C.c_double * NB),
                ("geom_type", C.c_int * NB), ("geom_a", C.c_double * 3 * NB), ("geom_b", C.c_double * 3 * NB),
                ("geom_r", C.c_double * NB)]


class OCfg(C.Structure):
    _fields_ = [("dt", C.c_double), ("gravity_z", C.c_double), ("kn", C.c_double), ("cn", C.c_double),
                ("ct", C.c_double), ("mu", C.c_double), ("max_ang_vel", C.c_double), ("max_turn", C.c_double), ("max_effort", C.c_double),
                ("hf_rows", C.c_int), ("hf_cols", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "physics_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", src, "-o", _SO, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.emloco_oracle_sizeof_model() == C.sizeof(OModel)
        assert _lib.emloco_oracle_sizeof_cfg() == C.sizeof(OCfg)
    return _lib


def _fill(dst, src):
    a = np.ascontiguousarray(src)
    C.memmove(dst, a.ctypes.data, a.nbytes)


def make_model(parent, offset, mass, com, inertia6, kp_joint, kd_joint, arm_joint, geom_type, geom_a, geom_b, geom_r):
    """kp/kd/arm are per joint, indexed by body (entry 0 unused)."""
    m = OModel()
    _fill(m.parent, np.asarray(parent, np.int32)); _fill(m.offset, np.asarray(offset, np.float64))
    _fill(m.mass, np.asarray(mass, np.float64)); _fill(m.com, np.asarray(com, np.float64))
    _fill(m.inertia, np.asarray(inertia6, np.float64))
    _fill(m.kp, np.asarray(kp_joint, np.float64)); _fill(m.kd, np.asarray(kd_joint, np.float64))
    _fill(m.arm, np.asarray(arm_joint, np.float64))
    _fill(m.geom_type, np.asarray(geom_type, np.int32)); _fill(m.geom_a, np.asarray(geom_a, np.float64))
    _fill(m.geom_b, np.asarray(geom_b, np.float64)); _fill(m.geom_r, np.asarray(geom_r, np.float64))
    return m


def make_cfg(dt, gravity_z=-9.81, kn=5e4, cn=1e3, ct=2e3, mu=1.0, max_ang_vel=100.0, max_effort=500.0, max_turn=0.3, hf_shape=(0, 0)):
    return OCfg(dt, gravity_z, kn, cn, ct, mu, max_ang_vel, max_turn, max_effort, hf_shape[0], hf_shape[1])


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def step(model, cfg, n_sub, root, jq, jw, target, height=None):
    """In-place on root [N,13], jq [N,23,4], jw [N,69] (float64).  Returns rb [N,24,13], dof_pos [N,69],
    contact [N,24,3], dof_force [N,69]."""
    N = root.shape[0]
    for a in (root, jq, jw):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    target = np.ascontiguousarray(target, np.float64)
    rb = np.zeros((N, NB, 13)); dof_pos = np.zeros((N, ND)); contact = np.zeros((N, NB, 3)); dof_force = np.zeros((N, ND))
    hf = None
    if height is not None:
        height = np.ascontiguousarray(height, np.int16)
        cfg.hf_rows, cfg.hf_cols = height.shape
        hf = _p(height, C.c_int16)
    lib().emloco_oracle_step(C.byref(model), C.byref(cfg), N, n_sub, _p(root), _p(jq), _p(jw), _p(target), hf,
                             _p(rb), _p(dof_pos), _p(contact), _p(dof_force))
    return rb, dof_pos, contact, dof_force


def refresh(model, root, jq, jw):
    N = root.shape[0]
    rb = np.zeros((N, NB, 13)); dof_pos = np.zeros((N, ND))
    lib().emloco_oracle_refresh(C.byref(model), N, _p(np.ascontiguousarray(root)), _p(np.ascontiguousarray(jq)),
                                _p(np.ascontiguousarray(jw)), _p(rb), _p(dof_pos))
    return rb, dof_pos


def expmap_to_quat(e):
    e = np.ascontiguousarray(e, np.float64).reshape(-1, 3)
    q = np.zeros((e.shape[0], 4))
    lib().emloco_oracle_expmap_to_quat(e.shape[0], _p(e), _p(q))
    return q
