#!/usr/bin/env python
"""Stages the handful of reference files the GPU-eager bar needs under baseline/_ref/ (git-ignored, travels with gpurun;
NOT part of the repository - nothing under baseline/_ref is ever committed).  Run in the build container:
    python scripts/stage_reference.py
The staged tree keeps the reference's relative paths, so `EMLOCO_REFERENCE=baseline/_ref` makes oracle/ref_extract.py load the
reference's own functions on the GPU box exactly as it does here from /root/reference."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("EMLOCO_REFERENCE_SRC", "/root/reference")
FILES = [
    "isaacgym/python/isaacgym/torch_utils.py",
    "pacer/pacer/utils/torch_utils.py", "pacer/pacer/utils/running_mean_std.py",
    "pacer/pacer/env/tasks/humanoid.py", "pacer/pacer/env/tasks/humanoid_amp.py", "pacer/pacer/env/tasks/humanoid_pedestrain_terrain.py",
    "pacer/pacer/env/util/traj_generator.py",
    "pacer/pacer/learning/common_agent.py", "pacer/pacer/learning/amp_continuous.py", "pacer/pacer/learning/amp_continuous_value.py",
    "pacer/pacer/learning/network_builder.py", "pacer/pacer/learning/amp_network_builder.py", "pacer/pacer/learning/amp_network_sept_builder.py",
    "pacer/pacer/learning/amp_network_sept_value_builder.py", "pacer/pacer/learning/value_pose_net.py", "pacer/pacer/learning/amp_models.py",
    "pacer/pacer/learning/amp_value_players.py", "pacer/pacer/data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml",
    "plausibl/test_value_mlp.py",
]


def main():
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: run this in the build container")
    dst = os.path.join(ROOT, "baseline", "_ref")
    for f in FILES:
        os.makedirs(os.path.dirname(os.path.join(dst, f)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), os.path.join(dst, f))
    print(f"staged {len(FILES)} reference files under {dst} (git-ignored)")


if __name__ == "__main__":
    main()
