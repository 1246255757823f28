"""CPU: host logic of the gymapi-shaped shim - the env-creation sequence of humanoid.py:643-946 and
humanoid_pedestrain_terrain.py:860-880 recorded into a pending sim (no device work before prepare_sim)."""
import os

import numpy as np
import pytest

from emloco_b200 import _lib
from emloco_b200 import gym_shim as G

REF_XML = "/root/reference/pacer/pacer/data/assets/mjcf"


def _create(n=3, scale=1.3):
    gym = G.acquire_gym()
    sim = gym.create_sim(0, -1, G.SIM_PHYSX, G.SimParams())
    opt = G.AssetOptions(); opt.angular_damping = 0.01; opt.max_angular_velocity = 100.0; opt.default_dof_drive_mode = G.DOF_MODE_NONE
    asset = gym.load_asset(sim, "/nonexistent", "smpl_humanoid.xml", opt)          # falls back to the packaged table
    for i in range(n):
        env = gym.create_env(sim, G.Vec3(-5, -5, 0), G.Vec3(5, 5, 5), 2)
        pose = G.Transform(); pose.p = G.Vec3(50 + i, 52.0, 0.93); pose.r = G.Quat(0, 0, 0, 1)
        h = gym.create_actor(env, asset, pose, "humanoid", i, 0, 0)
        gym.enable_actor_dof_force_sensors(env, h)
        prop = gym.get_asset_dof_properties(asset)
        prop["driveMode"] = G.DOF_MODE_POS
        prop["stiffness"] *= scale; prop["damping"] *= scale                       # humanoid.py:905-910
        gym.set_actor_dof_properties(env, h, prop)
    return gym, sim, asset


def test_asset_queries_match_the_model_table():
    gym, sim, asset = _create()
    assert gym.get_asset_rigid_body_count(asset) == 24 and gym.get_asset_dof_count(asset) == 69 == gym.get_asset_joint_count(asset)
    assert gym.find_asset_rigid_body_index(asset, "Pelvis") == 0 and gym.find_asset_rigid_body_index(asset, "right_foot") == -1
    assert len(gym.get_asset_actuator_properties(asset)) == 69 and gym.get_asset_actuator_properties(asset)[0].motor_effort == 500.0
    env = sim.envs[0]
    masses = [p.mass for p in gym.get_actor_rigid_body_properties(env, 0)]
    assert abs(sum(masses) - asset.model.total_mass) < 1e-9 and len(masses) == 24
    prop = gym.get_actor_dof_properties(env, 0)
    np.testing.assert_allclose(prop["stiffness"], asset.model.kp * 1.3, rtol=1e-6)
    assert set(prop.dtype.names) >= {"driveMode", "stiffness", "damping", "lower", "upper", "effort", "armature"}
    assert gym.find_actor_rigid_body_handle(env, 0, "L_Ankle") == asset.model.names.index("L_Ankle")
    props = gym.get_actor_rigid_shape_properties(env, 0)
    assert len(props) == 24
    props[2].filter = 7
    gym.set_actor_rigid_shape_properties(env, 0, props)
    assert gym.get_actor_rigid_shape_properties(env, 0)[2].filter == 7


def test_triangle_mesh_round_trips_the_height_field():
    gym, sim, _ = _create(1)
    rng = np.random.default_rng(0)
    hs = rng.integers(-200, 200, (37, 53)).astype(np.int16)
    x, y = np.arange(37) * 0.1, np.arange(53) * 0.1
    yy, xx = np.meshgrid(y, x)                                                       # Terrain.convert_heightfield_to_trimesh layout
    verts = np.stack([xx.flatten(), yy.flatten(), hs.flatten() * 0.005], 1).astype(np.float32)
    tm = G.TriangleMeshParams(); tm.nb_vertices = verts.shape[0]
    gym.add_triangle_mesh(sim, verts.flatten(order="C"), np.zeros(3, np.uint32), tm)
    np.testing.assert_array_equal(sim.height, hs)


def test_prepare_sim_rejects_what_the_kernel_cannot_do():
    gym = G.acquire_gym()
    with pytest.raises(_lib.EmlocoError, match="exactly one actor"):
        gym.prepare_sim(gym.create_sim(0, -1, G.SIM_PHYSX, G.SimParams()))
    gym, sim, asset = _create(2)
    p = gym.get_actor_dof_properties(sim.envs[1], 0); p["stiffness"] *= 2
    gym.set_actor_dof_properties(sim.envs[1], 0, p)
    with pytest.raises(_lib.EmlocoError, match="per-env drive gains"):
        gym.prepare_sim(sim)
    gym, sim, asset = _create(1)
    p = gym.get_actor_dof_properties(sim.envs[0], 0); p["driveMode"] = G.DOF_MODE_EFFORT
    gym.set_actor_dof_properties(sim.envs[0], 0, p)
    with pytest.raises(_lib.EmlocoError, match="DOF_MODE_POS"):
        gym.prepare_sim(sim)
    with pytest.raises(_lib.EmlocoError, match="prepare_sim"):
        gym.acquire_actor_root_state_tensor(_create(1)[1])


@pytest.mark.skipif(not os.path.isdir(REF_XML), reason="reference tree not present (GPU box)")
def test_packaged_table_equals_the_reference_mjcf():
    gym = G.acquire_gym()
    sim = gym.create_sim(0, -1, G.SIM_PHYSX, G.SimParams())
    a = gym.load_asset(sim, REF_XML, "smpl_humanoid.xml", None).model
    b = gym.load_asset(sim, "/nonexistent", "smpl_humanoid.xml", None).model
    assert a.names == b.names
    for k in ("parent", "offset", "mass", "com", "inertia", "kp", "kd", "armature", "limit_lo", "limit_hi", "geom_type", "geom_a", "geom_b", "geom_r"):
        np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-9, atol=1e-12, err_msg=k)
