"""On-disk / wire formats either side of the hot path, so that trained reference weights and real
JTA / JRDB trajectories flow through the new path unchanged.  Host-side, run once - not on the hot path.

  * rl_games checkpoint dict written by `CommonAgent.save` / `get_full_state_weights`
    (pacer/pacer/learning/common_agent.py:248-265,621-631; amp_continuous.py:76-96): keys `model` (state dict with the
    `a2c_network.` prefix, plus `value_mean_std.*` / `running_mean_std.*` when rl_games keeps them inside the model),
    `running_mean_std`, `amp_input_mean_std`, `optimizer`, `epoch`, `frame`.
  * LocoVal `.pth`: a bare `ValuePoseNet.state_dict()` (common_agent.py:252,262-264) - loads with `load_state_dict` as is.
  * saved-trajectory pkl `{id: {'pose': [24,3] or None, 'traj': [101,3]}}` written by
    social-transmotion/load_jta_traj.py:111-114 and consumed at pacer/pacer/env/util/traj_generator.py:44-52,136-143.
  * the value filter of social-transmotion/evaluate_jta.py:298-340 (keep modes with value >= threshold, fall back to the
    arg-max), batched instead of a batch-of-1 triple loop.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

from .policy import AMPSeptValueNetwork, RunningMeanStd

A2C_PREFIX = "a2c_network."


def _load_rms(rms: RunningMeanStd, sd, prefix=""):
    rms.running_mean.copy_(torch.as_tensor(sd[prefix + "running_mean"], dtype=torch.float64).reshape(rms.running_mean.shape))
    rms.running_var.copy_(torch.as_tensor(sd[prefix + "running_var"], dtype=torch.float64).reshape(rms.running_var.shape))
    if prefix + "count" in sd:
        rms.count.copy_(torch.as_tensor(sd[prefix + "count"], dtype=torch.float64).reshape(()))


def load_rl_games_checkpoint(ckpt, net: AMPSeptValueNetwork = None, obs_norm: RunningMeanStd = None,
                             amp_norm: RunningMeanStd = None, value_norm: RunningMeanStd = None):
    """ckpt: path (torch.load) or the dict itself.  Returns (net, obs_norm, amp_norm, value_norm, meta)."""
    if isinstance(ckpt, (str, bytes)):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    net = net or AMPSeptValueNetwork()
    obs_norm, amp_norm, value_norm = obs_norm or RunningMeanStd(1422), amp_norm or RunningMeanStd(3090), value_norm or RunningMeanStd(1)
    model = ckpt["model"]
    own = net.state_dict()
    sd, extra = {}, {}
    for k, v in model.items():
        kk = k[len(A2C_PREFIX):] if k.startswith(A2C_PREFIX) else k
        (sd if kk in own else extra)[kk] = torch.as_tensor(v)
    missing = [k for k in own if k not in sd]
    if missing:
        raise KeyError(f"checkpoint lacks network tensors: {missing[:4]}{'...' if len(missing) > 4 else ''}")
    for k in sd:
        if tuple(sd[k].shape) != tuple(own[k].shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(sd[k].shape)} != network shape {tuple(own[k].shape)}")
    net.load_state_dict(sd)
    # normalisers: rl_games 1.1.4 keeps them next to the model (get_stats_weights), later versions inside it
    if "running_mean_std" in ckpt:
        _load_rms(obs_norm, ckpt["running_mean_std"])
    elif "running_mean_std.running_mean" in extra:
        _load_rms(obs_norm, extra, "running_mean_std.")
    if "amp_input_mean_std" in ckpt:
        _load_rms(amp_norm, ckpt["amp_input_mean_std"])
    if "reward_mean_std" in ckpt:           # rl_games names the value normaliser this way in get_stats_weights
        _load_rms(value_norm, ckpt["reward_mean_std"])
    elif "value_mean_std.running_mean" in extra:
        _load_rms(value_norm, extra, "value_mean_std.")
    meta = {k: ckpt[k] for k in ("epoch", "frame", "last_mean_rewards") if k in ckpt}
    return net, obs_norm, amp_norm, value_norm, meta


def save_rl_games_checkpoint(path, net, obs_norm, amp_norm, value_norm=None, epoch=0, frame=0, optimizer=None):
    """The inverse, same key names - so a policy rolled out here can be resumed by the reference."""
    model = {A2C_PREFIX + k: v.detach().cpu() for k, v in net.state_dict().items()}
    rms = lambda r: {"running_mean": r.running_mean.cpu(), "running_var": r.running_var.cpu(), "count": r.count.cpu()}
    ckpt = {"model": model, "running_mean_std": rms(obs_norm), "amp_input_mean_std": rms(amp_norm), "epoch": epoch, "frame": frame}
    if value_norm is not None:
        ckpt["reward_mean_std"] = rms(value_norm)
    if optimizer is not None:
        ckpt["optimizer"] = optimizer
    torch.save(ckpt, path)
    return ckpt


def load_saved_trajs(path_or_dict, num_verts=101):
    """-> (traj [M,num_verts,3] float32, pose [M,24,3] float32 with NaN rows where the entry had pose None, ids).
    Accepts the pkl path (pickle / joblib pickles are plain pickles when uncompressed) or the dict itself."""
    d = path_or_dict
    if isinstance(d, (str, bytes)):
        try:
            import joblib
            d = joblib.load(d)
        except ImportError:
            with open(d, "rb") as f:
                d = pickle.load(f)
    ids = sorted(d.keys())
    traj = np.zeros((len(ids), num_verts, 3), np.float32)
    pose = np.full((len(ids), 24, 3), np.nan, np.float32)
    for i, k in enumerate(ids):
        t = np.asarray(d[k]["traj"], np.float32)
        if t.ndim != 2 or t.shape[0] < num_verts or t.shape[1] < 2:
            raise ValueError(f"trajectory {k}: expected at least [{num_verts}, 2..3], got {t.shape}")
        traj[i, :, :t.shape[1]] = t[:num_verts, :3]
        p = d[k].get("pose")
        if p is not None:
            pose[i] = np.asarray(p, np.float32).reshape(24, 3)
    return traj, pose, ids


def assign_trajs_to_envs(traj, init_xy, rng, real_frac=1.0, synthetic=None):
    """TrajGenerator.reset with flags.real_path (traj_generator.py:130-160): every env draws a stored trajectory without
    replacement (random.sample) and it is translated to start at the env's root xy.  -> verts [N,101,3], chosen ids."""
    n = init_xy.shape[0]
    k = int(round(n * real_frac))
    if k > traj.shape[0]:
        raise ValueError(f"{k} envs need real trajectories but only {traj.shape[0]} are stored (random.sample would raise too)")
    rid = rng.choice(traj.shape[0], size=k, replace=False)
    verts = np.zeros((n, traj.shape[1], 3), np.float32) if synthetic is None else synthetic.copy()
    sel = traj[rid].copy()
    sel[..., :2] += (init_xy[:k] - sel[:, 0, :2])[:, None, :]
    verts[:k] = sel
    return verts, rid


def filter_modes(values, threshold=0.7):
    """evaluate_jta.py:316-340 batched: values [S, M] (S scenes/persons, M modes) -> keep [S, M] bool.
    A mode is kept when its value >= threshold; a scene with no such mode keeps only its arg-max mode."""
    v = torch.as_tensor(values)
    keep = v >= threshold
    none = ~keep.any(dim=1)
    if none.any():
        best = v.argmax(dim=1)
        keep[none, best[none]] = True
    return keep


def score_and_filter(valuenet, pred_trajs, init_pose, init_vel, threshold=0.7, reference_compat=False, gt_trajs=None):
    """pred_trajs [S, M, 13, 2] (origin prepended, evaluate_jta.py:293-296), init_pose [S,24,3], init_vel [S,2] ->
    (values [S,M], keep [S,M]).  One LocoVal launch over S*M rows instead of S*M batch-of-1 calls; pose / velocity are
    shared by the M modes of a scene and never mutated here.

    DEVIATION (default): every mode is scored against the scene's ORIGINAL pose.  The reference loop
    (evaluate_jta.py:298-302) passes `init_pose.unsqueeze(0)` - a view - to `calc_embodied_motion_loss` twice per mode (the
    prediction, then the ground truth), and ValuePoseNet rotates that pose in place by the heading of the trajectory it is
    given (value_pose_net.py:97,141-144), so in the reference mode p of a scene is scored against the pose already rotated by
    the headings of predictions 0..p-1 and p ground-truth calls.  Only mode 0 agrees with the default here.
    reference_compat=True reproduces the reference numbers, still in one launch: the heading depends on the trajectory alone,
    so the cumulative rotation each mode sees is applied to a private copy of the pose up front (gt_trajs [S,13,2] required:
    its heading is part of the chain).  Values then match the sequential loop to float rounding."""
    S, M = pred_trajs.shape[:2]
    traj = pred_trajs.reshape(S * M, 13, -1).contiguous()
    pose = init_pose[:, None].expand(S, M, 24, 3).reshape(S * M, 24, 3).contiguous()
    vel = init_vel[:, None].expand(S, M, 2).reshape(S * M, 2).contiguous()
    if reference_compat:
        if gt_trajs is None:
            raise ValueError("reference_compat=True needs gt_trajs: the reference also scores the ground truth on the same pose")

        def heading(t):                                                   # value_pose_net.py:76-84
            x = t[..., 1, 0]
            x = torch.where(x.abs() < 1e-10, torch.full_like(x, 1e-10), x)
            return torch.atan2(t[..., 1, 1], x)
        th = heading(pred_trajs.float()) + heading(gt_trajs.float())[:, None]          # [S, M]: rotation added by mode p's two calls
        before = (torch.cumsum(th, dim=1) - th).reshape(S * M)                          # what mode p's prediction call finds
        c, sn = torch.cos(before)[:, None], torch.sin(before)[:, None]
        x, y = pose[..., 0].clone(), pose[..., 1].clone()
        pose[..., 0], pose[..., 1] = x * c + y * sn, -x * sn + y * c                    # row vector times [[c, -s], [s, c]] (:85-97)
        if M > 1:                                                                       # zeroed by every earlier call (:141-144)
            later = (torch.arange(S * M, device=pose.device) % M) > 0
            hide = [j for j, on in ((4, valuenet.hide_toe), (8, valuenet.hide_toe), (9, valuenet.hide_spine), (10, valuenet.hide_spine),
                                    (11, valuenet.hide_spine)) if on]
            for j in hide:
                pose[later, j] = 0
    was = valuenet.mutate_pose
    valuenet.mutate_pose = False
    try:
        with torch.no_grad():
            values = valuenet(traj, pose, vel).reshape(S, M)
    finally:
        valuenet.mutate_pose = was
    return values, filter_modes(values, threshold)


# ---- reference configuration files -> constructor arguments -----------------------------------------------------------
def kwargs_from_reference_cfg(env_cfg, train_cfg, run_flags=None):
    """Maps the reference's two YAML trees - `data/cfg/pacer.yaml` (env_cfg, the dict with the `env` key) and
    `data/cfg/train/rlg/amp_humanoid_smpl_sept_task.yaml` (train_cfg, the dict with the `params` key) - and the run.py
    command-line flags (`--real_path ... --adjust_root_vel --init_heading --heading_inversion --slow`, run.py:262-331) onto the
    arguments of this package:  -> dict(net=..., rollout=..., sim=..., traj=..., num_envs=...)
        AMPSeptValueNetwork(**net); Rollout(num_envs, net=..., **rollout, sim_cfg=sim, traj_cfg=traj)
    Keys the hot path does not use (viewer, logging, optimiser) are ignored; unsupported settings raise ValueError."""
    from . import _lib
    env, prm = env_cfg["env"], train_cfg["params"]
    cfg, nw = prm["config"], prm["network"]
    if nw["space"]["continuous"].get("learn_sigma", False) or not nw["space"]["continuous"].get("fixed_sigma", True):
        raise ValueError("only fixed_sigma / learn_sigma False is implemented (amp_humanoid_smpl_sept_task.yaml:19-27)")
    if not env.get("pdControl", True):
        raise ValueError("only pdControl is implemented (humanoid.py:905-910)")
    if env.get("numAMPObsSteps", 15) != 15 or env.get("numTrajSamples", 15) != 15:
        raise ValueError("kernels are specialised for 15 AMP steps and 15 trajectory samples (pacer.yaml:46,53)")
    net = dict(mlp_units=tuple(nw["mlp"]["units"]), task_units=tuple(nw["task_mlp"]["units"]),
               value_units=tuple(nw["value_mlp"]["units"]), disc_units=tuple(nw["disc"]["units"]),
               sigma_init=float(nw["space"]["continuous"]["sigma_init"]["val"]))
    rollout = dict(horizon=int(cfg["horizon_length"]), gamma=float(cfg["gamma"]), tau=float(cfg["tau"]),
                   task_reward_w=float(cfg["task_reward_w"]), disc_reward_w=float(cfg["disc_reward_w"]),
                   disc_reward_scale=float(cfg["disc_reward_scale"]),
                   inversion_penalty_scale=float(cfg.get("inversion_penalty_scale", 0.3)),
                   step_to_pred=int(env.get("stepToPred", 144)), normalize_value=bool(cfg.get("normalize_value", True)),
                   finetune=bool(cfg.get("player", {}).get("finetune", False)))
    sim = dict(episode_length=int(env["episodeLength"]), control_freq_inv=int(env["controlFrequencyInv"]),
               power_coefficient=float(env.get("power_coefficient", 0.0005)),
               location_coefficient=float(env.get("location_coefficient", 1.0)),
               traj_sample_dt=float(env["trajSampleTimestep"]),
               friction_mu=float(env.get("terrain", {}).get("staticFriction", 1.0)))
    traj = dict(speed_min=float(env["speedMin"]), speed_max=float(env["speedMax"]), accel_max=float(env["accelMax"]),
                sharp_turn_prob=float(env["sharpTurnProb"]), hybrid_init_prob=float(env.get("hybridInitProb", 0.5)))
    f = run_flags or {}
    flags = ((_lib.TRAJ_REAL_PATH if f.get("real_path") else 0) | (_lib.TRAJ_ADJUST_ROOT_VEL if f.get("adjust_root_vel") else 0)
             | (_lib.TRAJ_INIT_HEADING if f.get("init_heading") else 0) | (_lib.TRAJ_HEADING_INVERSION if f.get("heading_inversion") else 0)
             | (_lib.TRAJ_SLOW if f.get("slow") else 0))
    for k in ("fixed_path", "pred_path", "add_noise"):
        if f.get(k):
            raise ValueError(f"run.py flag --{k} is not implemented by the device-side trajectory reset")
    rollout["traj_flags"] = flags
    return dict(net=net, rollout=rollout, sim=sim, traj=traj, num_envs=int(env["numEnvs"]))
