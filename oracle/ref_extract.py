"""TEST INFRASTRUCTURE ONLY - loads the *reference's own* functions for the hot path.

Runs only in the build container (needs /root/reference, which does not exist on
the GPU box).  It is used by ``oracle/make_golden.py`` to produce the committed
fixtures under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference tree is absent).  Nothing here is copied into the product:
functions are pulled out of the reference files *by line range at run time* into a
temp module, because the enclosing modules import the absent Isaac Gym bindings
(SURVEY.md section 8c).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import textwrap
import types

REF = os.environ.get("EMLOCO_REFERENCE", "/root/reference")
PACER = os.path.join(REF, "pacer", "pacer")


def available() -> bool:
    return os.path.isdir(PACER)


def _lines(path, a, b):
    with open(path) as f:
        L = f.readlines()
    return "".join(L[a - 1:b])


_cache = {}


def _import_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns a namespace with the reference functions (torch, CPU)."""
    if "ns" in _cache:
        return _cache["ns"]
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float  # isaacgym/torch_utils.py:135 default arg
    if not hasattr(np, "int"):
        np.int = int
    import torch

    # stub packages so that `from isaacgym.torch_utils import *` resolves
    isaac = types.ModuleType("isaacgym")
    isaac.__path__ = []
    sys.modules.setdefault("isaacgym", isaac)
    itu = _import_path("isaacgym.torch_utils",
                       os.path.join(REF, "isaacgym/python/isaacgym/torch_utils.py"))
    isaac.torch_utils = itu
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = []
    sys.modules.setdefault("utils", utils_pkg)
    ptu = _import_path("utils.torch_utils", os.path.join(PACER, "utils/torch_utils.py"))
    utils_pkg.torch_utils = ptu

    tmp = tempfile.mkdtemp(prefix="emloco_ref_")
    hum = os.path.join(PACER, "env/tasks/humanoid.py")
    amp = os.path.join(PACER, "env/tasks/humanoid_amp.py")
    ter = os.path.join(PACER, "env/tasks/humanoid_pedestrain_terrain.py")
    header = ("import torch\nimport numpy as np\nfrom isaacgym.torch_utils import *\n"
              "from utils import torch_utils\nfrom typing import Tuple, List\n\n")
    # remove_base_rot is referenced (dead branch, upright=True) by the jit functions
    body = header
    body += ("@torch.jit.script\ndef remove_base_rot(quat):\n"
             "    base_rot = quat_conjugate(torch.tensor([[0.5, 0.5, 0.5, 0.5]]).to(quat))\n"
             "    shape = quat.shape[0]\n"
             "    return quat_mul(quat, base_rot.repeat(shape, 1))\n\n")
    body += _lines(hum, 1327, 1338)       # dof_to_obs_smpl
    body += "\n" + _lines(hum, 1626, 1687)  # compute_humanoid_observations_smpl_max
    body += "\n" + _lines(amp, 917, 971)    # build_amp_observations_smpl
    body += "\n" + _lines(ter, 1468, 1530)  # compute_humanoid_reset
    body += "\n" + _lines(ter, 1533, 1538)  # quat_apply_yaw
    body += "\n" + _lines(ter, 1549, 1592)  # compute_location_observations / reward
    p = os.path.join(tmp, "emloco_ref_jit.py")
    with open(p, "w") as f:
        f.write(body)
    jit = _import_path("emloco_ref_jit", p)

    # plain-python methods, re-hosted on small holder classes
    tg = os.path.join(PACER, "env/util/traj_generator.py")
    ca = os.path.join(PACER, "learning/common_agent.py")
    ac = os.path.join(PACER, "learning/amp_continuous.py")
    meth = header + "class TrajHolder:\n" + _lines(tg, 278, 296)
    meth += "\nclass TerrainHolder:\n" + _lines(ter, 1212, 1218) + "\n" + _lines(ter, 1221, 1226)
    # non-group branch of sample_height_points (ter:1282-1288)
    meth += ("        heights1 = heightsamples[px, py]\n"
             "        heights2 = heightsamples[px + 1, py + 1]\n"
             "        heights = torch.min(heights1, heights2)\n"
             "        return heights * self.vertical_scale\n")
    meth += "\nclass AgentHolder:\n" + _lines(ca, 573, 587)
    p2 = os.path.join(tmp, "emloco_ref_meth.py")
    with open(p2, "w") as f:
        f.write(meth)
    methm = _import_path("emloco_ref_meth", p2)

    # TrajGenerator.reset (traj_generator.py:60-237) on a holder class; its module-level `torch` / `random` names are
    # swapped by make_golden for proxies that replay recorded uniform draws
    p3 = os.path.join(tmp, "emloco_ref_trajreset.py")
    with open(p3, "w") as f:
        f.write("import numpy as np\nimport random\nimport torch\n\nclass TrajResetHolder:\n" + _lines(tg, 60, 237)
                + "\n" + _lines(tg, 261, 262))
    trajreset = _import_path("emloco_ref_trajreset", p3)

    # ValuePoseNet with a stub matplotlib
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    vpn = _import_path("emloco_ref_vpn", os.path.join(PACER, "learning/value_pose_net.py"))
    rms = _import_path("emloco_ref_rms", os.path.join(PACER, "utils/running_mean_std.py"))

    ns = types.SimpleNamespace(
        torch=torch, itu=itu, ptu=ptu, jit=jit, meth=methm, trajreset=trajreset,
        ValuePoseNet=vpn.ValuePoseNet, RunningMeanStd=rms.RunningMeanStd,
        left_to_right_index=[0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 19, 20, 21, 22, 23,
                             14, 15, 16, 17, 18],  # humanoid.py:334
    )
    _cache["ns"] = ns
    return ns
