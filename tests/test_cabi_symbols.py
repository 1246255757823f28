"""CPU: the C-ABI library loads and exports every symbol include/emloco.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "emloco.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emloco_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("emloco_create", "emloco_step", "emloco_post_step", "emloco_simulate", "emloco_reset_indexed",
                 "emloco_locoval_forward", "emloco_locoval_backward", "emloco_gae", "emloco_linear", "emloco_tensor"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from emloco_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in emloco.h but not exported"
    assert sorted(_lib.SYMBOLS) == _declared()


def test_struct_layouts_match_header_sizes():
    from emloco_b200 import _lib
    # emloco_cfg: 18 4-byte fields + 8 reserved ints; emloco_model: arrays as declared
    assert ctypes.sizeof(_lib.Cfg) == (18 + 8) * 4
    nb, nd = 24, 69
    assert ctypes.sizeof(_lib.Model) == 4 * (nb + 3 * nb + nb + 3 * nb + 6 * nb + 3 * nd + nb + 3 * nb + 3 * nb + nb + 2 * nd)


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from emloco_b200 import _lib
    from emloco_b200.sim import EmlocoSim
    with pytest.raises(_lib.EmlocoError):
        EmlocoSim(4)
    from emloco_b200.value_pose_net import ValuePoseNet
    net = ValuePoseNet(True, True)
    with pytest.raises(_lib.EmlocoError):
        net(torch.zeros(2, 13, 2), torch.zeros(2, 24, 3), torch.zeros(2, 2))


def test_state_dict_keys_match_reference_contract():
    from emloco_b200.value_pose_net import ValuePoseNet
    net = ValuePoseNet(True, True)
    assert list(net.state_dict().keys()) == [f"_network.fc{i}.{p}" for i in (1, 2, 3) for p in ("weight", "bias")]
    assert net._network.fc1.weight.shape == (49, 100) and net._network.fc2.weight.shape == (24, 49)
    assert sum(p.numel() for p in net.parameters()) == 6174


def test_mjcf_model_table():
    from emloco_b200.model import build_model_arrays
    a = build_model_arrays()
    assert len(a["names"]) == 24 and a["kp"].shape == (69,)
    assert a["names"][13] == "Head" and a["names"][7] == "R_Ankle"
    assert abs(a["pd_scale"][4] - 5) < 1e-6 and abs(a["pd_scale"][16] - 5) < 1e-6
    assert 60 < a["total_mass"] < 90


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in emloco_b200/_lib.py have the size (and, for the plain ones, the field offsets) a C compiler gives
    the structs of include/emloco.h - guards against the two drifting apart."""
    import ctypes as C
    import shutil
    import subprocess
    from emloco_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no C compiler")
    src = tmp_path / "sizes.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "emloco.h"
int main(void) {
    printf("emloco_cfg %zu\\n", sizeof(emloco_cfg));
    printf("emloco_model %zu\\n", sizeof(emloco_model));
    printf("emloco_rollout_cfg %zu\\n", sizeof(emloco_rollout_cfg));
    printf("emloco_post_sinks %zu\\n", sizeof(emloco_post_sinks));
    printf("emloco_traj_cfg %zu\\n", sizeof(emloco_traj_cfg));
    printf("emloco_motion_lib %zu\\n", sizeof(emloco_motion_lib));
    printf("emloco_chain_layer %zu\\n", sizeof(emloco_chain_layer));
    printf("chain.dep %zu\\n", offsetof(emloco_chain_layer, dep));
    printf("chain.d_head_out %zu\\n", offsetof(emloco_chain_layer, d_head_out));
    printf("rollout.d_value_stats %zu\\n", offsetof(emloco_rollout_cfg, d_value_stats));
    printf("motion.num_motions %zu\\n", offsetof(emloco_motion_lib, num_motions));
    printf("cfg.max_effort %zu\\n", offsetof(emloco_cfg, max_effort));
    printf("cfg.physics_impl %zu\\n", offsetof(emloco_cfg, physics_impl));
    printf("sinks.rows_only %zu\\n", offsetof(emloco_post_sinks, rows_only));
    printf("sinks.ld_amp %zu\\n", offsetof(emloco_post_sinks, ld_amp));
    printf("sinks.task_lo2 %zu\\n", offsetof(emloco_post_sinks, task_lo2));
    printf("traj.seed %zu\\n", offsetof(emloco_traj_cfg, seed));
    printf("traj.num_waypoints %zu\\n", offsetof(emloco_traj_cfg, num_waypoints));
    return 0;
}
''')
    exe = tmp_path / "sizes"
    inc = os.path.join(ROOT, "include")
    subprocess.run([gcc, "-std=c99", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    out = {k: int(v) for k, v in out.items()}
    assert out["emloco_cfg"] == C.sizeof(_lib.Cfg)
    assert out["emloco_model"] == C.sizeof(_lib.Model)
    assert out["emloco_rollout_cfg"] == C.sizeof(_lib.RolloutCfg)
    assert out["emloco_post_sinks"] == C.sizeof(_lib.PostSinks)
    assert out["emloco_traj_cfg"] == C.sizeof(_lib.TrajCfg)
    assert out["emloco_motion_lib"] == C.sizeof(_lib.MotionLib)
    assert out["emloco_chain_layer"] == C.sizeof(_lib.ChainLayer)
    assert out["chain.dep"] == _lib.ChainLayer.dep.offset and out["chain.d_head_out"] == _lib.ChainLayer.d_head_out.offset
    assert out["rollout.d_value_stats"] == _lib.RolloutCfg.d_value_stats.offset
    assert out["motion.num_motions"] == _lib.MotionLib.num_motions.offset
    assert out["cfg.max_effort"] == _lib.Cfg.max_effort.offset and out["cfg.physics_impl"] == _lib.Cfg.physics_impl.offset
    assert out["sinks.rows_only"] == _lib.PostSinks.rows_only.offset and out["sinks.ld_amp"] == _lib.PostSinks.ld_amp.offset
    assert out["sinks.task_lo2"] == _lib.PostSinks.task_lo2.offset
    assert out["traj.seed"] == _lib.TrajCfg.seed.offset and out["traj.num_waypoints"] == _lib.TrajCfg.num_waypoints.offset
