"""CPU: the formats either side of the path: rl_games checkpoint dict, saved-trajectory pkl, mode filter."""
import os
import pickle

import numpy as np
import pytest
import torch


def test_rl_games_checkpoint_round_trip(tmp_path):
    from emloco_b200.formats import load_rl_games_checkpoint, save_rl_games_checkpoint
    from emloco_b200.policy import AMPSeptValueNetwork, RunningMeanStd
    torch.manual_seed(0)
    net = AMPSeptValueNetwork()
    on, an, vn = RunningMeanStd(1422), RunningMeanStd(3090), RunningMeanStd(1)
    on.running_mean.normal_(); on.running_var.uniform_(0.5, 2); an.running_mean.normal_(); vn.running_var.fill_(3.0); on.count.fill_(1234.0)
    p = os.path.join(tmp_path, "Humanoid.pth")
    ck = save_rl_games_checkpoint(p, net, on, an, vn, epoch=25000, frame=123)
    assert all(k.startswith("a2c_network.") for k in ck["model"])
    assert {"a2c_network.actor_mlp.0.weight", "a2c_network._task_mlp.2.bias", "a2c_network._disc_logits.weight", "a2c_network.sigma",
            "a2c_network._value_logits.bias"} <= set(ck["model"])
    net2, on2, an2, vn2, meta = load_rl_games_checkpoint(p)
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(on.running_mean, on2.running_mean) and torch.equal(on.running_var, on2.running_var) and on2.count.item() == 1234.0
    assert torch.equal(an.running_mean, an2.running_mean) and vn2.running_var.item() == 3.0
    assert meta == {"epoch": 25000, "frame": 123}
    # normalisers kept inside the model dict (later rl_games layouts) and a shape mismatch
    ck2 = {"model": dict(ck["model"])}
    ck2["model"]["running_mean_std.running_mean"] = on.running_mean + 1
    ck2["model"]["running_mean_std.running_var"] = on.running_var
    ck2["model"]["value_mean_std.running_mean"] = torch.tensor([0.5]); ck2["model"]["value_mean_std.running_var"] = torch.tensor([2.0])
    _, on3, _, vn3, _ = load_rl_games_checkpoint(ck2)
    assert torch.allclose(on3.running_mean, on.running_mean + 1) and vn3.running_mean.item() == 0.5
    bad = {"model": dict(ck["model"])}
    bad["model"]["a2c_network.mu.weight"] = torch.zeros(3, 3)
    with pytest.raises(ValueError):
        load_rl_games_checkpoint(bad)
    del bad["model"]["a2c_network.mu.weight"]
    with pytest.raises(KeyError):
        load_rl_games_checkpoint(bad)


def test_saved_traj_pkl_and_assignment(tmp_path):
    from emloco_b200.formats import assign_trajs_to_envs, load_saved_trajs
    rng = np.random.default_rng(0)
    d = {i: {"pose": (rng.normal(size=(24, 3)) if i % 3 else None), "traj": rng.normal(size=(101, 3)).cumsum(0)} for i in range(20)}
    p = os.path.join(tmp_path, "jta_trajs.pkl")
    with open(p, "wb") as f:
        pickle.dump(d, f)
    traj, pose, ids = load_saved_trajs(p)
    assert traj.shape == (20, 101, 3) and pose.shape == (20, 24, 3) and ids == list(range(20))
    assert np.isnan(pose[0]).all() and not np.isnan(pose[1]).any()
    np.testing.assert_allclose(traj[5], d[5]["traj"].astype(np.float32))
    xy = rng.uniform(50, 58, (8, 2))
    verts, rid = assign_trajs_to_envs(traj, xy, np.random.default_rng(1))
    assert verts.shape == (8, 101, 3) and len(set(rid.tolist())) == 8
    np.testing.assert_allclose(verts[:, 0, :2], xy, atol=1e-5)                         # starts at the env's root
    np.testing.assert_allclose(verts[:, 1:, :2] - verts[:, :1, :2], traj[rid][:, 1:, :2] - traj[rid][:, :1, :2], atol=1e-4)
    with pytest.raises(ValueError):
        assign_trajs_to_envs(traj, rng.uniform(50, 58, (30, 2)), np.random.default_rng(1))
    with pytest.raises(ValueError):
        load_saved_trajs({0: {"pose": None, "traj": np.zeros((50, 3))}})


def test_filter_modes_matches_reference_loop():
    from emloco_b200.formats import filter_modes
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 1, (200, 5)); v[:10] *= 0.5                                     # some scenes with no mode above threshold
    keep = filter_modes(torch.from_numpy(v), 0.7).numpy()
    for s in range(v.shape[0]):                                                        # evaluate_jta.py:316-340
        filt = [m for m in range(5) if v[s, m] >= 0.7]
        if not filt:
            filt = [int(np.argmax(v[s]))]
        assert sorted(np.nonzero(keep[s])[0].tolist()) == filt


def test_reference_yaml_maps_onto_our_defaults():
    """kwargs_from_reference_cfg on the reference's own YAML files (when the tree is present; an excerpt otherwise): the
    constants this package hard-codes as defaults ARE the reference's configuration (pacer.yaml,
    train/rlg/amp_humanoid_smpl_sept_task.yaml)."""
    import inspect
    import os
    from emloco_b200 import _lib
    from emloco_b200.formats import kwargs_from_reference_cfg
    from emloco_b200.policy import AMPSeptValueNetwork
    from emloco_b200.rollout import Rollout
    root = "/root/reference/pacer/pacer/data/cfg"
    if os.path.isdir(root):
        import yaml
        env_cfg = yaml.safe_load(open(os.path.join(root, "pacer.yaml")))
        train_cfg = yaml.safe_load(open(os.path.join(root, "train/rlg/amp_humanoid_smpl_sept_task.yaml")))
    else:
        env_cfg = {"env": dict(numEnvs=1600, episodeLength=168, controlFrequencyInv=2, power_coefficient=0.0005, location_coefficient=1,
                               trajSampleTimestep=0.4, stepToPred=144, speedMin=0.0005, speedMax=3.0, accelMax=2.0, sharpTurnProb=0.02,
                               hybridInitProb=0.5, numAMPObsSteps=15, numTrajSamples=15, pdControl=True,
                               terrain=dict(staticFriction=1.0))}
        train_cfg = {"params": dict(
            network=dict(space=dict(continuous=dict(sigma_init=dict(val=-2.9), fixed_sigma=True, learn_sigma=False)),
                         mlp=dict(units=[2048, 1024]), task_mlp=dict(units=[512, 256]), value_mlp=dict(units=[15, 6]),
                         disc=dict(units=[1024, 512])),
            config=dict(horizon_length=32, gamma=0.99, tau=0.95, task_reward_w=0.5, disc_reward_w=0.5, disc_reward_scale=2,
                        inversion_penalty_scale=0.3, normalize_value=True, player=dict(finetune=True)))}
    k = kwargs_from_reference_cfg(env_cfg, train_cfg, dict(real_path=True, adjust_root_vel=True, init_heading=True))
    # network: the defaults of AMPSeptValueNetwork
    sig = inspect.signature(AMPSeptValueNetwork.__init__).parameters
    for name, v in k["net"].items():
        assert sig[name].default == v, (name, sig[name].default, v)
    # rollout: the defaults of Rollout
    sig = inspect.signature(Rollout.__init__).parameters
    for name in ("horizon", "gamma", "tau", "task_reward_w", "disc_reward_w", "disc_reward_scale", "inversion_penalty_scale",
                 "step_to_pred", "normalize_value"):
        assert sig[name].default == k["rollout"][name], (name, sig[name].default, k["rollout"][name])
    assert k["rollout"]["finetune"] is True and k["rollout"]["traj_flags"] == 1 | 2 | 4
    # sim: emloco_default_cfg
    d = _lib.default_cfg()
    for name, v in k["sim"].items():
        assert abs(getattr(d, name) - v) < 1e-7, (name, getattr(d, name), v)
    # trajectory generator: the defaults of EmlocoSim.traj_cfg (pacer.yaml:45,55-61)
    assert k["traj"] == dict(speed_min=0.0005, speed_max=3.0, accel_max=2.0, sharp_turn_prob=0.02, hybrid_init_prob=0.5)
    with __import__("pytest").raises(ValueError):
        kwargs_from_reference_cfg(env_cfg, train_cfg, dict(pred_path=True))


def test_running_mean_std_fp32_copies_follow_every_kind_of_update():
    """the fp32 copies the kernels read must follow in-place updates AND attribute re-assignment (fresh
    tensors whose _version is 0 again, the reference's way: utils/running_mean_std.py:93-96), at a fixed address."""
    import torch
    from emloco_b200.policy import RunningMeanStd
    m = RunningMeanStd(5)
    mean, var = m.f32()
    ptrs = (mean.data_ptr(), var.data_ptr(), m.inv_std().data_ptr())
    m.running_mean.add_(1.0)
    assert torch.equal(m.f32()[0], torch.ones(5))
    for k in (2.0, 3.0):
        m.running_mean = torch.full((5,), k, dtype=torch.float64)
        m.running_var = torch.full((5,), k * k, dtype=torch.float64)
        mean, var = m.f32()
        assert torch.equal(mean, torch.full((5,), k)) and torch.equal(var, torch.full((5,), k * k))
        torch.testing.assert_close(m.inv_std(), torch.full((5,), 1.0 / (k * k + 1e-5) ** 0.5))
    assert ptrs == (mean.data_ptr(), var.data_ptr(), m.inv_std().data_ptr())


def test_running_mean_std_update_matches_reference_golden():
    """RunningMeanStd.update against the reference's training-mode forward (utils/running_mean_std.py:86-96), three batches."""
    import os
    import numpy as np
    import torch
    from conftest import GOLDEN
    from emloco_b200.policy import RunningMeanStd
    g = np.load(os.path.join(GOLDEN, "rms_update.npz"))
    m = RunningMeanStd(g["x0"].shape[1])
    for i in range(3):
        m.update(torch.from_numpy(g[f"x{i}"]))
        np.testing.assert_allclose(m.running_mean.numpy(), g[f"mean{i}"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(m.running_var.numpy(), g[f"var{i}"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(m.count.numpy(), g[f"count{i}"])
