"""ctypes binding of libemloco_b200.so (include/emloco.h).  No fallback: if the CUDA library is
missing or does not load, importing the product path raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libemloco_b200.so")
NB, ND = 24, 69

SYMBOLS = [
    "emloco_create", "emloco_destroy", "emloco_default_cfg", "emloco_tensor", "emloco_set_height_field",
    "emloco_set_pd_targets", "emloco_simulate", "emloco_reset_indexed", "emloco_post_step", "emloco_step",
    "emloco_step_host", "emloco_locoval_forward", "emloco_locoval_backward", "emloco_locoval_forward_host",
    "emloco_plausibl_mlp_forward", "emloco_gae", "emloco_reset_done", "emloco_sample_actions",
    "emloco_disc_reward", "emloco_rollout_record", "emloco_normalize", "emloco_physics_step", "emloco_split_bf16",
    "emloco_linear_bf16x3", "emloco_set_post_sinks", "emloco_linear_bf16x3_rows", "emloco_timeout_gather",
    "emloco_rollout_record_deferred", "emloco_fill_next_values", "emloco_traj_reset", "emloco_set_traj_reset", "emloco_locoval_backward_pose", "emloco_locoval_train_step", "emloco_locoval_train_workspace_bytes", "emloco_linear_bf16x3_head", "emloco_linear_chain", "emloco_linear_chain_workspace_ints", "emloco_linear_chain_trace", "emloco_sample_actions_parts", "emloco_linear", "emloco_xform", "emloco_ppo_heads", "emloco_disc_heads", "emloco_amp_dropout_mask",
    "emloco_rms_update", "emloco_adam_begin", "emloco_grad_sumsq", "emloco_adam_clip", "emloco_axpy", "emloco_sum_parts", "emloco_player_record", "emloco_dp_reduce_shard", "emloco_dp_adam_shard", "emloco_motion_state", "emloco_amp_obs_demo", "emloco_set_env_models", "emloco_sync", "emloco_last_error", "emloco_version",
]


class EmlocoError(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("device", C.c_int32), ("sim_dt", C.c_float), ("substeps", C.c_int32),
                ("control_freq_inv", C.c_int32), ("gravity_z", C.c_float), ("contact_stiffness", C.c_float),
                ("contact_damping", C.c_float), ("friction_damping", C.c_float), ("friction_mu", C.c_float),
                ("contact_offset", C.c_float), ("max_ang_vel", C.c_float), ("angular_damping", C.c_float),
                ("episode_length", C.c_int32), ("power_coefficient", C.c_float), ("location_coefficient", C.c_float),
                ("fail_dist", C.c_float), ("traj_sample_dt", C.c_float), ("max_effort", C.c_float), ("max_turn", C.c_float),
                ("physics_impl", C.c_int32), ("reserved", C.c_int32 * 5)]


class RolloutCfg(C.Structure):
    _fields_ = [("inversion_penalty_scale", C.c_float), ("reward_scale", C.c_float), ("value_mean", C.c_float),
                ("value_std", C.c_float), ("disc_reward_scale", C.c_float), ("gamma", C.c_float),
                ("step_to_pred", C.c_int32), ("unnorm_value", C.c_int32), ("d_value_stats", C.c_void_p)]


class MotionLib(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("d_gts", "d_grs", "d_lrs", "d_gvs", "d_gavs", "d_dvs", "d_length", "d_dt", "d_bodies", "d_num_frames",
                                          "d_start")] + [("num_motions", C.c_int32), ("reserved", C.c_int32)]


class PostSinks(C.Structure):
    _fields_ = [("obs_copy", C.c_void_p), ("amp_copy", C.c_void_p), ("flip_copy", C.c_void_p), ("obs_mean", C.c_void_p), ("obs_inv_std", C.c_void_p),
                ("self_hi", C.c_void_p), ("self_lo", C.c_void_p), ("ld_self", C.c_int64),
                ("task_hi", C.c_void_p), ("task_lo", C.c_void_p), ("ld_task", C.c_int64),
                ("amp_mean", C.c_void_p), ("amp_inv_std", C.c_void_p),
                ("amp_hi", C.c_void_p), ("amp_lo", C.c_void_p), ("ld_amp", C.c_int64),
                ("rows_only", C.c_int32), ("reserved", C.c_int32),
                ("self_hi2", C.c_void_p), ("self_lo2", C.c_void_p), ("task_hi2", C.c_void_p), ("task_lo2", C.c_void_p)]


class ChainLayer(C.Structure):
    """emloco_chain_layer (include/emloco.h): one dense layer of an emloco_linear_chain launch."""
    _fields_ = [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", C.c_int64), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("ldw", C.c_int64),
                ("d_bias", C.c_void_p), ("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32), ("relu", C.c_int32), ("dep", C.c_int32),
                ("d_y32", C.c_void_p), ("ldy", C.c_int64), ("y_hi", C.c_void_p), ("y_lo", C.c_void_p), ("ldy16", C.c_int64),
                ("d_head_w", C.c_void_p), ("d_head_bias", C.c_void_p), ("d_head_part", C.c_void_p), ("d_head_out", C.c_void_p)]


class TrajCfg(C.Structure):
    """emloco_traj_cfg (include/emloco.h): TrajGenerator.reset parameters, pacer.yaml:45,55-61 defaults."""
    _fields_ = [("dtheta_max", C.c_float), ("speed_min", C.c_float), ("speed_max", C.c_float), ("accel_max", C.c_float),
                ("sharp_turn_prob", C.c_float), ("hybrid_init_prob", C.c_float), ("flags", C.c_int32), ("origin_relative", C.c_int32),
                ("seed", C.c_uint64), ("pool", C.c_void_p), ("pool_count", C.c_int64), ("uniform", C.c_void_p), ("ld_uniform", C.c_int64),
                ("waypoint_traj", C.c_void_p), ("init_pose", C.c_void_p), ("init_vel", C.c_void_p), ("inverted", C.c_void_p),
                ("num_waypoints", C.c_int32), ("reserved", C.c_int32)]


TRAJ_REAL_PATH, TRAJ_ADJUST_ROOT_VEL, TRAJ_INIT_HEADING, TRAJ_HEADING_INVERSION, TRAJ_SLOW, TRAJ_DEFERRED = 1, 2, 4, 8, 16, 32
TRAJ_RAND_COLS = 405


class Model(C.Structure):
    _fields_ = [("parent", C.c_int32 * NB), ("offset", C.c_float * 3 * NB), ("mass", C.c_float * NB),
                ("com", C.c_float * 3 * NB), ("inertia", C.c_float * 6 * NB), ("kp", C.c_float * ND),
                ("kd", C.c_float * ND), ("armature", C.c_float * ND), ("geom_type", C.c_int32 * NB),
                ("geom_a", C.c_float * 3 * NB), ("geom_b", C.c_float * 3 * NB), ("geom_r", C.c_float * NB),
                ("pd_offset", C.c_float * ND), ("pd_scale", C.c_float * ND)]


T_IDS = {n: i for i, n in enumerate([
    "root_state", "dof_state", "rb_state", "contact", "dof_force", "pd_target", "obs", "flip_obs", "rew", "rew_raw",
    "reset", "terminate", "progress", "amp_obs", "traj_verts", "betas", "height", "joint_quat", "actions"])}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmlocoError(f"{LIB_PATH} is missing - run `python -m emloco_b200.build` (or __graft_entry__.build()); "
                          "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.emloco_last_error.restype = C.c_char_p
    lib.emloco_version.restype = C.c_char_p
    lib.emloco_default_cfg.argtypes = [C.POINTER(Cfg)]
    lib.emloco_default_cfg.restype = None
    lib.emloco_create.argtypes = [C.POINTER(Cfg), C.POINTER(Model), C.POINTER(vp)]
    lib.emloco_destroy.argtypes = [vp]
    lib.emloco_tensor.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.emloco_set_height_field.argtypes = [vp, vp, i32, i32]
    lib.emloco_set_pd_targets.argtypes = [vp, vp, vp]
    lib.emloco_simulate.argtypes = [vp, vp]
    lib.emloco_reset_indexed.argtypes = [vp, vp, i32, vp]
    lib.emloco_post_step.argtypes = [vp, i32, vp]
    lib.emloco_step.argtypes = [vp, vp, vp]
    lib.emloco_physics_step.argtypes = [vp, vp, vp]
    lib.emloco_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.emloco_locoval_forward.argtypes = [vp, i32, i32, vp, vp, vp, vp, i64, i32, vp]
    lib.emloco_locoval_backward.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.emloco_locoval_backward_pose.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp]
    lib.emloco_locoval_train_workspace_bytes.argtypes = [i64]
    lib.emloco_locoval_train_workspace_bytes.restype = i64
    lib.emloco_locoval_train_step.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64] + [C.c_float] * 7 + [i32, vp]
    lib.emloco_locoval_forward_host.argtypes = [vp, i32, i32, vp, vp, vp, vp, i64, i32, i32]
    lib.emloco_plausibl_mlp_forward.argtypes = [vp, vp, vp, i64, vp]
    lib.emloco_gae.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, f32, f32, vp]
    lib.emloco_linear.argtypes = [vp, i64, vp, vp, vp, i64, i64, i32, i32, vp, vp, f32, i32, i32, vp]
    lib.emloco_reset_done.argtypes = [vp, vp, vp, vp]
    lib.emloco_sample_actions.argtypes = [vp, i64, vp, vp, vp, vp, i64, i32, vp]
    lib.emloco_sample_actions_parts.argtypes = [vp, i64, i32, i64, vp, i64, vp, vp, vp, vp, i64, i32, vp]
    lib.emloco_disc_reward.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, vp]
    lib.emloco_rollout_record.argtypes = [C.POINTER(RolloutCfg), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.emloco_normalize.argtypes = [vp, i64, vp, i64, i64, i32, vp, vp, f32, vp]
    lib.emloco_split_bf16.argtypes = [vp, i64, i64, i32, vp, vp, f32, vp, vp, i64, vp]
    lib.emloco_linear_bf16x3.argtypes = [vp, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, vp, vp, i64, vp]
    lib.emloco_linear_bf16x3_rows.argtypes = [vp, vp, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, vp, vp, i64, vp]
    lib.emloco_linear_bf16x3_head.argtypes = [vp, vp, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.emloco_linear_chain.argtypes = [C.POINTER(ChainLayer), i32, C.POINTER(i32), i32, vp, i64, vp]
    lib.emloco_linear_chain_workspace_ints.argtypes = [C.POINTER(ChainLayer), i32]
    lib.emloco_linear_chain_workspace_ints.restype = i64
    lib.emloco_linear_chain_trace.argtypes = [vp]
    lib.emloco_timeout_gather.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp]
    lib.emloco_rollout_record_deferred.argtypes = [C.POINTER(RolloutCfg)] + [vp] * 12 + [i64] + [vp] * 6
    lib.emloco_fill_next_values.argtypes = [C.POINTER(RolloutCfg), vp, vp, vp, i64, vp]
    lib.emloco_set_post_sinks.argtypes = [vp, C.POINTER(PostSinks)]
    lib.emloco_xform.argtypes = [vp, i64, vp, vp, i64, vp, vp, f32, vp, i64, vp, f32, f32, vp, i64, vp, i64, vp, vp, i64, vp, vp, i64, vp, vp, i64, i32, vp]
    lib.emloco_ppo_heads.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i64, i32, f32, f32, f32, f32, f32, vp]
    lib.emloco_disc_heads.argtypes = [vp, vp, vp, i64, i64, f32, vp]
    lib.emloco_amp_dropout_mask.argtypes = [vp, vp, i64, f32, vp]
    lib.emloco_rms_update.argtypes = [vp, i64, i64, i32, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    lib.emloco_adam_begin.argtypes = [vp, vp]
    lib.emloco_grad_sumsq.argtypes = [vp, i64, vp, vp, vp]
    lib.emloco_adam_clip.argtypes = [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, vp]
    lib.emloco_axpy.argtypes = [vp, vp, f32, i64, vp]
    lib.emloco_sum_parts.argtypes = [vp, i32, i64, vp, i64, i32, vp]
    lib.emloco_dp_reduce_shard.argtypes = [vp, vp, i64, i64, vp, vp, i32, vp]
    lib.emloco_dp_adam_shard.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp, i32, vp, f32, f32, f32, f32, f32, f32, vp]
    lib.emloco_player_record.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, f32, f32, f32, i32, f32, f32, vp]
    lib.emloco_motion_state.argtypes = [C.POINTER(MotionLib), vp, vp, i64, vp, vp, vp, vp, vp]
    lib.emloco_amp_obs_demo.argtypes = [C.POINTER(MotionLib), vp, vp, i64, i32, f32, vp, vp]
    lib.emloco_set_env_models.argtypes = [vp, vp]
    lib.emloco_sync.argtypes = [vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("emloco_last_error", "emloco_version", "emloco_default_cfg", "emloco_locoval_train_workspace_bytes", "emloco_linear_chain_workspace_ints"):
            fn.restype = C.c_int
    _lib = lib
    return lib


# kernels launched per successful ABI call (the bench's `gpu_launches` claim is counted here, not estimated)
LAUNCHES = {"emloco_step": 2, "emloco_physics_step": 1, "emloco_post_step": 1, "emloco_simulate": 1, "emloco_reset_done": 2,
            "emloco_reset_indexed": 1, "emloco_traj_reset": 1, "emloco_linear": 1, "emloco_normalize": 1, "emloco_sample_actions": 1, "emloco_sample_actions_parts": 1,
            "emloco_disc_reward": 1, "emloco_rollout_record": 1, "emloco_gae": 1, "emloco_locoval_forward": 1,
            "emloco_locoval_backward": 1, "emloco_locoval_backward_pose": 1, "emloco_locoval_train_step": 2, "emloco_plausibl_mlp_forward": 1, "emloco_step_host": 2,
            "emloco_locoval_forward_host": 1, "emloco_split_bf16": 1, "emloco_linear_bf16x3": 1, "emloco_linear_bf16x3_rows": 1, "emloco_linear_bf16x3_head": 2, "emloco_linear_chain": 1,
            "emloco_timeout_gather": 1, "emloco_rollout_record_deferred": 1, "emloco_fill_next_values": 1,
            "emloco_xform": 1, "emloco_ppo_heads": 1, "emloco_disc_heads": 1, "emloco_amp_dropout_mask": 1, "emloco_rms_update": 2,
            "emloco_adam_begin": 1, "emloco_grad_sumsq": 2, "emloco_adam_clip": 1, "emloco_axpy": 1, "emloco_sum_parts": 1, "emloco_player_record": 1, "emloco_dp_reduce_shard": 2, "emloco_dp_adam_shard": 1, "emloco_motion_state": 1, "emloco_amp_obs_demo": 1}
launch_count = 0
mac_count = 0          # multiply-accumulates of the dense-layer launches (M * N * K each), for the benches' FLOP figures


def check(rc, what=""):
    global launch_count
    launch_count += LAUNCHES.get(what, 0)
    if rc != 0:
        msg = load().emloco_last_error().decode()
        raise EmlocoError(f"{what} failed ({rc}): {msg}")


def default_cfg(**over) -> Cfg:
    c = Cfg()
    load().emloco_default_cfg(C.byref(c))
    for k, v in over.items():
        if not hasattr(c, k):
            raise EmlocoError(f"unknown cfg field {k}")
        setattr(c, k, v)
    return c


def make_model(arrs) -> Model:
    m = Model()

    def fill(dst, src, dt):
        a = np.ascontiguousarray(np.asarray(src, dtype=dt))
        assert a.nbytes == C.sizeof(dst), (a.shape, C.sizeof(dst))
        C.memmove(dst, a.ctypes.data, a.nbytes)
    fill(m.parent, arrs["parent"], np.int32); fill(m.offset, arrs["offset"], np.float32)
    fill(m.mass, arrs["mass"], np.float32); fill(m.com, arrs["com"], np.float32)
    fill(m.inertia, arrs["inertia6"], np.float32); fill(m.kp, arrs["kp"], np.float32)
    fill(m.kd, arrs["kd"], np.float32); fill(m.armature, arrs["armature"], np.float32)
    fill(m.geom_type, arrs["geom_type"], np.int32); fill(m.geom_a, arrs["geom_a"], np.float32)
    fill(m.geom_b, arrs["geom_b"], np.float32); fill(m.geom_r, arrs["geom_r"], np.float32)
    fill(m.pd_offset, arrs["pd_offset"], np.float32); fill(m.pd_scale, arrs["pd_scale"], np.float32)
    return m
