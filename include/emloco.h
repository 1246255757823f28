/* emloco_b200 - C ABI of the B200-native EmLoco hot path (libemloco_b200.so).
 *
 * Plain C, plain pointers and sizes; no torch types.  Every entry point names the reference
 * interface it replaces (paths relative to ImIntheMiddle/EmLoco).  All `d_` pointers are CUDA
 * device pointers on the sim's device, all `h_` pointers are host pointers.  Calls are
 * stream-ordered on the `cudaStream_t` passed as `void* stream` (NULL = default stream) and never
 * synchronise unless the name says so (`*_host`, `emloco_sync`).  A sim handle is not re-entrant.
 *
 * Return convention: 0 = ok, negative = error (EMLOCO_E*); text via emloco_last_error().
 * (Isaac Gym itself returns None / prints and the caller quit()s: pacer/pacer/env/tasks/base_task.py:239-241.)
 */
#ifndef EMLOCO_H
#define EMLOCO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMLOCO_OK        0
#define EMLOCO_EINVAL   -1
#define EMLOCO_ECUDA    -2
#define EMLOCO_ENOMEM   -3

#define EMLOCO_NB 24
#define EMLOCO_ND 69

/* Simulation parameters: gymapi.SimParams as filled by parse_sim_params
 * (pacer/pacer/utils/config.py:141-174) from pacer/pacer/data/cfg/pacer.yaml:93-104. */
typedef struct emloco_cfg {
    int32_t num_envs;
    int32_t device;               /* CUDA ordinal (compute_device_id of gym.create_sim) */
    float   sim_dt;               /* 1/60 */
    int32_t substeps;             /* 2: PhysX substeps per gym.simulate */
    int32_t control_freq_inv;     /* 2: gym.simulate calls per env step (base_task.py:792-797) */
    float   gravity_z;            /* -9.81 */
    float   contact_stiffness;    /* N/m per contact point (implicit spring) */
    float   contact_damping;      /* N s/m per contact point, normal */
    float   friction_damping;     /* N s/m per contact point, tangential (sticking regime) */
    float   friction_mu;          /* 1.0 (pacer.yaml:71-72) */
    float   contact_offset;       /* 0.02 */
    float   max_ang_vel;          /* 100 (humanoid.py:685-688) */
    float   angular_damping;      /* 0.01 */
    int32_t episode_length;       /* 168 (pacer.yaml:12) */
    float   power_coefficient;    /* 0.0005 */
    float   location_coefficient; /* 1 */
    float   fail_dist;            /* 4.0 (humanoid_traj.py:31) */
    float   traj_sample_dt;       /* 0.4 */
    float   max_effort;           /* 500: drive torque limit per DOF (MJCF motor gear -> Isaac Gym DOF `effort`); <= 0 off */
    float   max_turn;             /* 0.3 rad: a sub-step is refined until no body turns more than this per piece; <= 0 off */
    int32_t physics_impl;         /* 0: lane-per-env step kernel (default, needs the SMPL tree), 1: warp-per-env kernel */
    int32_t reserved[5];
} emloco_cfg;

/* Articulation + collision model: what gym.load_asset/create_actor build from
 * mjcf/smpl_humanoid.xml (pacer/pacer/env/tasks/humanoid.py:643-835), flattened. */
typedef struct emloco_model {
    int32_t parent[EMLOCO_NB];
    float   offset[EMLOCO_NB][3];
    float   mass[EMLOCO_NB];
    float   com[EMLOCO_NB][3];
    float   inertia[EMLOCO_NB][6];   /* xx xy xz yy yz zz about COM, body frame */
    float   kp[EMLOCO_ND];           /* DOF stiffness after the mass/77*kp_scale scaling, humanoid.py:905-910 */
    float   kd[EMLOCO_ND];
    float   armature[EMLOCO_ND];
    int32_t geom_type[EMLOCO_NB];    /* 0 sphere, 1 capsule, 2 box */
    float   geom_a[EMLOCO_NB][3];
    float   geom_b[EMLOCO_NB][3];
    float   geom_r[EMLOCO_NB];
    float   pd_offset[EMLOCO_ND];    /* _build_pd_action_offset_scale, humanoid.py:950-1025 */
    float   pd_scale[EMLOCO_ND];
} emloco_model;

typedef struct emloco_sim emloco_sim;

/* Tensor ids for emloco_tensor(): the gym.acquire_*_tensor family (humanoid.py:137-150) plus the
 * task buffers of base_task.py:96-112 that the fused post-step kernel writes. */
enum {
    EMLOCO_T_ROOT_STATE = 0,   /* f32 [N,13]     acquire_actor_root_state_tensor */
    EMLOCO_T_DOF_STATE,        /* f32 [N*69,2]   acquire_dof_state_tensor */
    EMLOCO_T_RB_STATE,         /* f32 [N*24,13]  acquire_rigid_body_state_tensor */
    EMLOCO_T_CONTACT,          /* f32 [N*24,3]   acquire_net_contact_force_tensor */
    EMLOCO_T_DOF_FORCE,        /* f32 [N*69]     acquire_dof_force_tensor */
    EMLOCO_T_PD_TARGET,        /* f32 [N,69]     set_dof_position_target_tensor storage */
    EMLOCO_T_OBS,              /* f32 [N,1422]   obs_buf */
    EMLOCO_T_FLIP_OBS,         /* f32 [N,1422]   _flip_obs_buf (humanoid.py:1052-1057) */
    EMLOCO_T_REW,              /* f32 [N]        rew_buf */
    EMLOCO_T_REW_RAW,          /* f32 [N,2]      reward_raw */
    EMLOCO_T_RESET,            /* i64 [N]        reset_buf */
    EMLOCO_T_TERMINATE,        /* i64 [N]        _terminate_buf */
    EMLOCO_T_PROGRESS,         /* i64 [N]        progress_buf */
    EMLOCO_T_AMP_OBS,          /* f32 [N,15,206] _amp_obs_buf (humanoid_amp.py:92-98) */
    EMLOCO_T_TRAJ_VERTS,       /* f32 [N,101,3]  TrajGenerator._verts (traj_generator.py:35-36) */
    EMLOCO_T_BETAS,            /* f32 [N,17]     humanoid_betas */
    EMLOCO_T_HEIGHT,           /* i16 [rows,cols] Terrain.heightsamples (read-only alias: change it with emloco_set_height_field) */
    EMLOCO_T_JOINT_QUAT,       /* f32 [N,23,4]   internal joint rotations (xyzw) */
    EMLOCO_T_ACTIONS,          /* f32 [N,69]     self.actions */
    EMLOCO_T_COUNT
};
enum { EMLOCO_DTYPE_F32 = 0, EMLOCO_DTYPE_I64 = 1, EMLOCO_DTYPE_I16 = 2 };

/* gym.create_sim + load_asset + create_env/create_actor x N + prepare_sim
 * (base_task.py:238, humanoid.py:643-946, base_task.py:128). */
int emloco_create(const emloco_cfg* cfg, const emloco_model* model, emloco_sim** out);
int emloco_destroy(emloco_sim* sim);

/* Per-env body models (reference `has_shape_variation`, pacer/pacer/env/tasks/humanoid.py:597-739: every env
 * simulates the body generated from its own SMPL shape parameters, with PD gains scaled by that body's mass, :905-910).
 * h_env_models: HOST array [num_envs][576] fp32, per env the floats
 *   offset[24][3] | mass[24] | com[24][3] | inertia[24][6] (xx xy xz yy yz zz about the COM) | kp[24] | kd[24] | armature[24]
 *   (per joint, index = body, entry 0 unused) | geom_a[24][3] | geom_b[24][3] | geom_r[24] | geom_bound[24]
 * (same meaning as the fields of emloco_model; geom_bound = largest distance from the body origin to a contact point).
 * Topology, primitive types and the action -> PD-target map stay those of the emloco_model given to emloco_create.
 * NULL restores the shared model.  Synchronises the device.  The shape parameters the observations carry (`betas` tensor) are
 * set by the caller, like `humanoid_betas` in the reference. */
int emloco_set_env_models(emloco_sim* sim, const float* h_env_models);
void emloco_default_cfg(emloco_cfg* cfg);

/* gym.acquire_*_tensor -> gymapi.Tensor{data_address, shape, dtype} (isaacgym/python/isaacgym/gymtorch.py:61-106).
 * shape must hold 4 entries. */
int emloco_tensor(emloco_sim* sim, int which, void** d_ptr, int64_t* shape, int32_t* ndim, int32_t* dtype);

/* Terrain height field: Terrain.heightsamples (humanoid_pedestrain_terrain.py:1166-1171), int16, first dim x. */
int emloco_set_height_field(emloco_sim* sim, const int16_t* h_samples, int32_t rows, int32_t cols);

/* gym.set_dof_position_target_tensor (humanoid.py:1202): copies [N,69] targets into the sim. */
int emloco_set_pd_targets(emloco_sim* sim, const float* d_targets, void* stream);

/* gym.simulate + fetch_results (base_task.py:795,258): `substeps` articulated-body sub-steps of
 * sim_dt/substeps using the stored PD targets; refreshes every state tensor. */
int emloco_simulate(emloco_sim* sim, void* stream);

/* gym.set_actor_root_state_tensor_indexed + set_dof_state_tensor_indexed (humanoid.py:470-475):
 * re-reads root_state / dof_state for the listed envs (the caller has written them through the
 * aliases), rebuilds joint rotations and rigid-body state by forward kinematics, zeroes contact
 * and DOF force.  d_env_ids == NULL means all envs. */
int emloco_reset_indexed(emloco_sim* sim, const int32_t* d_env_ids, int32_t n, void* stream);

/* Optional extra outputs of the fused post-step kernel (and of the post-step part of emloco_reset_done), so that the
 * rollout's copies and operand conversions cost no extra pass over HBM:
 *   obs_copy / amp_copy / flip_copy   experience rows `experience_buffer.update_data('obses' / 'amp_obs' / 'flip_obs', n, ...)`
 *                        of play_steps (amp_continuous_value.py:46,70,74-75; emloco_reset_done writes obs_copy only)
 *   self_* / task_*      clamp((obs-mean)*inv_std, +-5) (utils/running_mean_std.py:82-84) split into bf16 hi/lo: the A operands
 *                        of emloco_linear_bf16x3 for the actor/critic input (cols 0..367) and the task MLP (cols 368..1421)
 *   amp_*                the same for the discriminator input [N,3090]
 * inv_std = 1/sqrt(var + eps) as fp32.  Any group may be NULL.  The struct is copied; pointers must stay valid until replaced.
 * Pitches (ld_*) in elements, multiples of 8; hi/lo base pointers 16-byte aligned. */
typedef struct emloco_post_sinks {
    float*    obs_copy;
    float*    amp_copy;
    float*    flip_copy;      /* experience row of the mirrored observation (motion_sym_loss, amp_continuous_value.py:74-75) */
    const float* obs_mean; const float* obs_inv_std;
    uint16_t* self_hi; uint16_t* self_lo; int64_t ld_self;
    uint16_t* task_hi; uint16_t* task_lo; int64_t ld_task;
    const float* amp_mean; const float* amp_inv_std;
    uint16_t* amp_hi; uint16_t* amp_lo; int64_t ld_amp;
    /* rows_only != 0: the experience rows are the ONLY destination of the mirrored observation (flip_copy) and of the AMP
     * ring (amp_copy) - EMLOCO_T_FLIP_OBS / EMLOCO_T_AMP_OBS are then not refreshed by emloco_post_step (73 MB less written
     * per 4096-env step).  The sim remembers, per env, which row holds its ring (the amp_copy row of its last post-step, or
     * its EMLOCO_T_AMP_OBS row after a reset) and shifts from there: that row must stay intact until the env's next
     * emloco_post_step. */
    int32_t   rows_only; int32_t reserved;
    /* second destination of the self / task operands (same pitches), written by emloco_post_step ONLY - not by the post-step part
     * of emloco_reset_done: the rows keep the observation that FOLLOWED the step (what `_eval_critic(next obs)` reads) while the
     * reset patches the first set for the next policy pass.  Lets both passes run in one emloco_linear_chain launch.  May be NULL. */
    uint16_t* self_hi2; uint16_t* self_lo2; uint16_t* task_hi2; uint16_t* task_lo2;
} emloco_post_sinks;
int emloco_set_post_sinks(emloco_sim* sim, const emloco_post_sinks* sinks /* NULL clears */);

/* post_physics_step (humanoid_amp.py:139-157 -> humanoid.py:1211-1232 -> humanoid_amp_task.py:62-86):
 * progress += advance_progress; obs, flip obs, reward, reset/terminate, AMP obs ring.  One fused kernel. */
int emloco_post_step(emloco_sim* sim, int32_t advance_progress, void* stream);

/* BaseTask.step (base_task.py:245-265): pre_physics_step (actions -> PD targets, humanoid.py:1184-1209),
 * control_freq_inv x simulate, post_physics_step.  d_actions [N,69]. */
int emloco_step(emloco_sim* sim, const float* d_actions, void* stream);

/* The first half of emloco_step on its own - pre_physics_step + control_freq_inv x gym.simulate in one kernel -
 * so that a caller can place its own work (or a timing event) between physics and emloco_post_step(sim, 1, stream). */
int emloco_physics_step(emloco_sim* sim, const float* d_actions, void* stream);

/* Same through HOST buffers (the vec-env call a CPU-side user makes, run.py:148-160): copies
 * actions in, steps, copies obs/rew/reset out, synchronises.  Any output pointer may be NULL. */
int emloco_step_host(emloco_sim* sim, const float* h_actions, float* h_obs, float* h_rew,
                     int64_t* h_reset, float* h_amp_obs);

/* env_reset(done_indices) of play_steps (pacer/pacer/learning/amp_continuous_value.py:45 -> vec_task_wrappers.py:36-39 ->
 * humanoid.py:455-481, humanoid_amp.py:284-293,499-502) entirely on the device: every env whose reset_buf is set takes
 * its root/DOF state from d_init_root [N,13] / d_init_dof [N*69,2], rigid-body state by forward kinematics, progress = 0,
 * reset/terminate cleared, contact forces zeroed, observations recomputed, AMP history filled with the current step
 * (_init_amp_obs_default).  The mocap-sampled initial state of _reset_ref_state_init is the caller's to put in d_init_*. */
int emloco_reset_done(emloco_sim* sim, const float* d_init_root, const float* d_init_dof, void* stream);

/* Device-side `_reset_task` for the envs that reset: TrajGenerator.reset
 * (pacer/pacer/env/util/traj_generator.py:60-237) + the LocoVal inputs captured at reset
 * (humanoid_pedestrain_terrain.py:493-516, exposed by vec_task_wrappers.py:47-63).  One warp regenerates one env:
 * random-walk polyline (:63-118), optional real trajectory from a device-resident pool [P,101,3] (:116-160; the
 * reference re-reads a pickle and loops in python), optional speed matching (--adjust_root_vel :98-104,:149-155),
 * first-segment alignment with the root velocity (--init_heading :176-234) and heading inversion (:195-200).
 * Random draws: `uniform` [N, ld_uniform >= 405] replays explicit U[0,1) draws in the reference's order (cols 0-99
 * dtheta, 100-199 sharp angle, 200-299 bernoulli, 300 heading, 301-400 dspeed, 401 speed, 402 real-data draw,
 * 403 pool pick, 404 inversion); NULL = Philox4x32-10 keyed by (seed, env, per-env reset count).
 * Differences from the reference, both in the random stream only: pool picks are independent (random.sample draws
 * without replacement); --add_noise / --fixed_path / --pred_path are not implemented.
 * Outputs (any may be NULL): waypoint_traj [N,num_waypoints,3] (num_waypoints <= 15; 0 means 15; LocoVal reads the first 13,
 * amp_continuous_value.py:127), init_pose [N,24,3], init_vel [N,2], inverted [N];
 * origin_relative != 0 stores waypoints / pose relative to their first row, as the vec-env getters return them. */
#define EMLOCO_TRAJ_REAL_PATH          1
#define EMLOCO_TRAJ_ADJUST_ROOT_VEL    2
#define EMLOCO_TRAJ_INIT_HEADING       4
#define EMLOCO_TRAJ_HEADING_INVERSION  8
#define EMLOCO_TRAJ_SLOW              16
#define EMLOCO_TRAJ_DEFERRED          32   /* emloco_set_traj_reset only, see there */
#define EMLOCO_TRAJ_RAND_COLS        405
typedef struct emloco_traj_cfg {
    float    dtheta_max, speed_min, speed_max, accel_max, sharp_turn_prob;   /* 2, 0.0005, 3, 2, 0.02 (pacer.yaml:55-61) */
    float    hybrid_init_prob;                                               /* 0.5 (pacer.yaml:45) */
    int32_t  flags;
    int32_t  origin_relative;
    uint64_t seed;
    const float* pool; int64_t pool_count;
    const float* uniform; int64_t ld_uniform;
    float*   waypoint_traj; float* init_pose; float* init_vel; uint8_t* inverted;
    int32_t  num_waypoints; int32_t reserved;
} emloco_traj_cfg;

/* Regenerates the trajectories of the envs whose reset_buf is set RIGHT NOW (flags are left as they are).
 * cfg == NULL runs the stage stored by emloco_set_traj_reset(..EMLOCO_TRAJ_DEFERRED..) and clears reset/terminate. */
int emloco_traj_reset(emloco_sim* sim, const emloco_traj_cfg* cfg, void* stream);

/* Makes emloco_reset_done run the trajectory reset as its last stage - after the observations of the reset envs were
 * recomputed from the OLD polyline, which is the reference's order (humanoid_amp_task.py:54-57: `super()._reset_envs`
 * computes observations, then `_reset_task`).  NULL switches it off.  The struct is copied.
 * With EMLOCO_TRAJ_DEFERRED in cfg->flags emloco_reset_done leaves reset/terminate set and does NOT run the stage: the
 * caller runs it with emloco_traj_reset(sim, NULL, stream) - on another stream if it likes, nothing before the next
 * emloco_post_step reads the polylines - any time before that post-step. */
int emloco_set_traj_reset(emloco_sim* sim, const emloco_traj_cfg* cfg);

/* ---- LocoVal: ValuePoseNet (pacer/pacer/learning/value_pose_net.py:10-159) ----
 * weights: fc1.weight[H1,IN] fc1.bias[H1] fc2.weight[H2,H1] fc2.bias[H2] fc3.weight[1,H2] fc3.bias[1]
 * packed in that order (state-dict order of `_network.fc{1,2,3}.{weight,bias}`).
 * flags: bit0 use_pose, bit1 use_vel, bit2 hide_toe, bit3 hide_spine, bit4 normalize,
 *        bit5 write the rotated/zeroed pose back into d_pose (the reference's in-place side effect, :97,:141-144),
 *        bit6 force the CUDA-core kernel (batches >= 1024 of the full variant otherwise run on the tensor cores).
 * traj [B,T,traj_stride] (only x,y read; T = 13 or 5), pose [B,24,3], vel [B,2], value [B]. */
int emloco_locoval_forward(const float* d_traj, int32_t traj_stride, int32_t num_waypoints, float* d_pose,
                           const float* d_vel, const float* d_weights, float* d_value, int64_t batch,
                           int32_t flags, void* stream);
/* d value / d traj for the EmLoco loss (calc_embodied_motion_loss :151-159; gradient flows through LocoVal
 * into the predictor, social-transmotion/train_jta.py:288-308).  grad_value [B] -> grad_traj [B,T,traj_stride].
 * The pose/vel passed must be the ORIGINAL (un-rotated) inputs of the forward call. */
int emloco_locoval_backward(const float* d_traj, int32_t traj_stride, int32_t num_waypoints, const float* d_pose,
                            const float* d_vel, const float* d_weights, const float* d_grad_value,
                            float* d_grad_traj, int64_t batch, int32_t flags, void* stream);
/* The same with the gradient chain through the mutated pose.  The reference rotates / zeroes init_pose in place WITH autograd
 * history (value_pose_net.py:97,141-144), so in the multi-modal loop of social-transmotion/train_jta.py:294-296 (one call per
 * mode, same init_pose tensor) the loss of mode i also reaches the trajectories of modes < i through the pose.
 * grad_pose_out [B,24,3] (or NULL): gradient w.r.t. the pose as this call left it; grad_pose_in [B,24,3] (or NULL): gradient
 * w.r.t. the pose as it came in (feeds the previous call's grad_pose_out).  d_pose = the pose as it came in. */
int emloco_locoval_backward_pose(const float* d_traj, int32_t traj_stride, int32_t num_waypoints, const float* d_pose,
                                 const float* d_vel, const float* d_weights, const float* d_grad_value, float* d_grad_traj,
                                 const float* d_grad_pose_out, float* d_grad_pose_in, int64_t batch, int32_t flags, void* stream);
/* LocoVal fine-tuning step of the rollout: the `_do_finetune` block of AMPValueAgent.play_steps
 * (pacer/pacer/learning/amp_continuous_value.py:122-146) with the optimiser of common_agent.py:94-96 -
 *   valid = nonzero(game_combined); pred = valuenet(traj, pose, vel)[valid];
 *   target = (game_combined[valid] - r_min) / (r_max - r_min); loss = sum((pred - target)^2); backward; AdamW; game_combined = 0
 * - entirely on the device (the reference syncs with the host for `valid` every control step).  Nothing happens when no env
 * is valid.  d_weights [NW] packed as for emloco_locoval_forward, updated in place; d_m / d_v: AdamW moments [NW]; d_step [1]:
 * optimiser step count (float); d_stats [4] += {loss, sum pred, sum target, count} (vnet_loss / vnet_pred / vnet_gt, :141-144).
 * d_pose is NOT mutated (the reference passes clones here, vec_task_wrappers.py:54-59).  d_workspace:
 * emloco_locoval_train_workspace_bytes(N) bytes.  Deterministic (ordered compaction, fixed reduction order). */
int64_t emloco_locoval_train_workspace_bytes(int64_t num_envs);
int emloco_locoval_train_step(const float* d_traj, int32_t traj_stride, int32_t num_waypoints, const float* d_pose,
                              const float* d_vel, float* d_game_combined, float* d_weights, float* d_m, float* d_v, float* d_step,
                              float* d_stats, void* d_workspace, int64_t num_envs, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float r_min, float r_max, int32_t flags, void* stream);
/* Host-buffer scoring (the batch-of-1 filter loop of social-transmotion/evaluate_jta.py:298-302, batched). */
int emloco_locoval_forward_host(const float* h_traj, int32_t traj_stride, int32_t num_waypoints, const float* h_pose,
                                const float* h_vel, const float* h_weights, float* h_value, int64_t batch,
                                int32_t flags, int32_t device);
/* plausibl/test_value_mlp.py:24-113 MLP.forward: 24 -> 12 -> 6 -> 1, no sigmoid.  x [B,24], packed weights as above. */
int emloco_plausibl_mlp_forward(const float* d_x, const float* d_weights, float* d_value, int64_t batch, void* stream);

/* ---- GAE: discount_values (pacer/pacer/learning/common_agent.py:573-587) ----
 * all [T,N] row-major f32; adv and ret ("mb_returns = mb_advs + mb_values", amp_continuous_value.py:163) written. */
int emloco_gae(const float* d_dones, const float* d_values, const float* d_rewards, const float* d_next_values,
               float* d_adv, float* d_ret, int32_t T, int64_t N, float gamma, float tau, void* stream);

/* ---- dense layers of the actor / critic / discriminator (network builders) ----
 * y[M,N] = act( norm(x)[M,K] @ W[N,K]^T + b ), optional input normalisation
 * clamp((x-mean)/sqrt(var+eps),+-5) (utils/running_mean_std.py:60-84) folded into the operand load.
 * x row stride ldx, y row stride ldy (lets callers concatenate without a copy).  relu: 0/1.
 * mean/var may be NULL.  fp32 in/out; products by bf16x3 tcgen05 MMAs (see below) when `use_tensor_cores`, else fp32 FMA. */
int emloco_linear(const float* d_x, int64_t ldx, const float* d_w, const float* d_b, float* d_y, int64_t ldy,
                  int64_t M, int32_t N, int32_t K, const float* d_mean, const float* d_var, float eps,
                  int32_t relu, int32_t use_tensor_cores, void* stream);

/* ---- per-step rollout arithmetic of play_steps (amp_continuous_value.py:34-178) ---- */
/* Gaussian policy head: a = mu + exp(logstd)*noise, neglogp = 0.5*sum(((a-mu)/sigma)^2) + 0.5*log(2pi)*A + sum(logstd)
 * (rl_games==1.1.4 ModelA2CContinuousLogStd, pinned in pacer/requirements.txt, not vendored; fixed logstd -2.9,
 * amp_humanoid_smpl_sept_task.yaml:19-27).  mu rows have stride ldmu; noise/actions [N,A]; neglogp [N] or NULL. */
int emloco_sample_actions(const float* d_mu, int64_t ldmu, const float* d_logstd, const float* d_noise, float* d_actions,
                          float* d_neglogp, int64_t N, int32_t A, void* stream);
/* The same with mu given as `parts` partial sums (the split-K output of the mu layer: emloco_linear_bf16x3 with a split count in
 * bits 20..23 of `relu` leaves `parts` matrices [N, ldmu], part_stride floats apart); they are added in order, the sum goes to
 * d_mu_out [N, ldout] (may be NULL) and is what the action is sampled around. */
int emloco_sample_actions_parts(const float* d_mu_parts, int64_t ldmu, int32_t parts, int64_t part_stride, float* d_mu_out, int64_t ldout,
                                const float* d_logstd, const float* d_noise, float* d_actions, float* d_neglogp, int64_t N, int32_t A,
                                void* stream);
/* _calc_disc_rewards + _combine_rewards (learning/amp_continuous.py:675-692, 659-664):
 * disc = -log(max(1 - sigmoid(logit), 1e-4)) * scale; combined = w_task*task + w_disc*disc.  Either output may be NULL;
 * d_logit == NULL means d_disc already holds the AMP rewards (combine only). */
int emloco_disc_reward(const float* d_logit, const float* d_task_rew, float* d_disc, float* d_combined, int64_t M,
                       float scale, float w_task, float w_disc, void* stream);
/* Everything play_steps does per env after the nets of step n (:63-118): inversion penalty, reward shaper scale,
 * value un-normalisation (utils/running_mean_std.py:77-79) of the pre-step critic output (d_value_raw -> d_mb_values,
 * may be NULL) and of the next-obs critic output with `next_vals *= 1 - terminated`, AMP reward, and the
 * discounted-return bookkeeping that produces the LocoVal regression target.  d_state is [6,N]:
 * current_rewards, current_lengths, current_combined_rewards, discount_coefs, game_combined_rewards, terminated_flags. */
typedef struct emloco_rollout_cfg {
    float inversion_penalty_scale;   /* 0.3 */
    float reward_scale;              /* reward_shaper.scale_value 1 */
    float value_mean, value_std;     /* value_mean_std running stats: std = sqrt(var + 1e-5) */
    float disc_reward_scale;         /* 2 */
    float gamma;                     /* 0.99 */
    int32_t step_to_pred;            /* 144 */
    int32_t unnorm_value;            /* normalize_value */
    const float* d_value_stats;      /* optional device pointer to {mean, std} (fp32): when non-NULL it replaces value_mean /
                                        value_std and is read by the kernels at run time, so that a value_mean_std update
                                        (every epoch in the reference, common_agent.py:440-442) reaches a captured CUDA graph */
} emloco_rollout_cfg;
int emloco_rollout_record(const emloco_rollout_cfg* cfg, const float* d_rew, const int64_t* d_reset,
                          const int64_t* d_terminate, const float* d_value_raw, const float* d_next_value_raw,
                          const float* d_disc_logit, const uint8_t* d_inverted, float* d_mb_values, float* d_mb_rewards,
                          float* d_mb_dones, float* d_mb_next_values, float* d_mb_amp_rewards, float* d_state, int64_t N,
                          void* stream);

/* ---- tensor-core path of the same dense layers: fp32 operands carried as two bf16 terms (x = hi + lo), three tcgen05
 * bf16 MMAs per k-step into an fp32 TMEM accumulator (A*W ~= Ah*Wh + Ah*Wl + Al*Wh): fp32-grade products at tensor-core
 * rate, which is what keeps the 1e-3 parity bar of north_star (csrc/linear_tc.cu).  bf16 values are raw uint16_t bits.
 *
 * emloco_split_bf16: x[M,K] fp32 (row stride ldx), optional normalisation clamp((x-mean)/sqrt(var+eps),+-5)
 * (utils/running_mean_std.py:82-84) -> hi/lo [M,K] bf16 with row pitch ld16 (ld16 % 8 == 0; pad columns are never read).
 * emloco_linear_bf16x3: y = act(A W^T + b) with A = a_hi+a_lo [M,K] (pitch lda), W = w_hi+w_lo [N,K] (pitch ldw);
 * writes fp32 y32 [M,N] (pitch ldy) and/or the split of y as the next layer's operand y_hi/y_lo (pitch ldy16, N % 32 == 0).
 * All bf16 base pointers 16-byte aligned, pitches multiples of 8 elements.  `relu`: bit 0 = ReLU; bits 8..19 optionally force
 * the N-extent of the output tile (128 or 256; 0 = chosen from the shape; +0x800 = the CTA-pair kernels, 256 rows per
 * pair) - used by the tests and for tuning; bits 20..23 = split-K count S > 1 for skinny layers whose tile count is far below
 * the SM count: y32 then receives S partial matrices [M, ldy], M * ldy floats apart (bias in the first), to be added by the
 * consumer (emloco_sample_actions_parts); needs a plain fp32 output - no ReLU, split output, fused head or row count. */
int emloco_split_bf16(const float* d_x, int64_t ldx, int64_t M, int32_t K, const float* d_mean, const float* d_var, float eps,
                      uint16_t* d_hi, uint16_t* d_lo, int64_t ld16, void* stream);
int emloco_linear_bf16x3(const uint16_t* d_a_hi, const uint16_t* d_a_lo, int64_t lda, const uint16_t* d_w_hi,
                         const uint16_t* d_w_lo, int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K,
                         int32_t relu, float* d_y32, int64_t ldy, uint16_t* d_y_hi, uint16_t* d_y_lo, int64_t ldy16,
                         void* stream);

/* emloco_linear_bf16x3 over a COMPACTED row set: *d_rows (device, <= M) rows are valid, output tiles beyond it are skipped. */
int emloco_linear_bf16x3_rows(const int32_t* d_rows, const uint16_t* d_a_hi, const uint16_t* d_a_lo, int64_t lda,
                              const uint16_t* d_w_hi, const uint16_t* d_w_lo, int64_t ldw, const float* d_bias, int64_t M,
                              int32_t N, int32_t K, int32_t relu, float* d_y32, int64_t ldy, uint16_t* d_y_hi, uint16_t* d_y_lo,
                              int64_t ldy16, void* stream);

/* The same GEMM with a single-output head fused into its epilogue - the `value` / `_disc_logits` layers that follow the last
 * hidden layer (amp_network_sept_builder.py:69-80, amp_network_builder.py:81-84): head_out[m] = head_bias[0] +
 * sum_n act(y[m][n]) * head_w[n], taken from the fp32 accumulator (no rounding of the hidden layer to the split format).
 * The epilogue writes one partial per 64-column group into d_head_part [M, ceil(N/64)]; a second small kernel adds them in
 * column order (deterministic, independent of the tile choice).  d_y32 / y_hi may all be NULL when only the head is wanted.
 * d_rows may be NULL (all M rows).  Not available with the CTA-pair tiles. */
int emloco_linear_bf16x3_head(const int32_t* d_rows, const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* w_hi,
                              const uint16_t* w_lo, int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K, int32_t relu,
                              float* d_y32, int64_t ldy, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy16, const float* d_head_w,
                              const float* d_head_bias, float* d_head_part, float* d_head_out, void* stream);

/* ---- layer chain: all dense layers of ONE network pass in a single persistent launch --------------------------------------
 * Replaces the sequence of nn.Linear / ReLU modules that `AMPSeptValueBuilder.Network.eval_actor / eval_critic / eval_disc`
 * walk per call (amp_network_sept_builder.py:69-111, amp_network_builder.py:81-84; called from get_action_values,
 * _eval_critic and _calc_amp_rewards, amp_continuous_value.py:53,85,93).  Every layer is an emloco_linear_bf16x3 problem
 * (same operand format, same arithmetic, bit-identical outputs); `dep` names the layer whose split output (y_hi / y_lo) is this
 * layer's A operand, -1 when the operand is complete before the launch.  The layers' 128 x 128 output tiles are handed out to the
 * SMs from one queue, so that 4096-row layers of different widths fill the 148 SMs without the wave quantisation and launch gaps
 * of one launch per layer; a tile starts as soon as the 128 rows it reads are complete (per-row-block counters in d_workspace).
 *   order: n_segments triples {layer, first tile, tile count}; tiles of a layer are numbered row block major (tile =
 *          row_block * ceil(N/128) + column block).  Every tile must appear exactly once and after all tiles of `dep` that cover
 *          its row block (checked on the host: EMLOCO_EINVAL otherwise).  NULL = the layers one after the other.
 *   head_w / head_bias / head_part / head_out: the single-output layer that follows (emloco_linear_bf16x3_head semantics);
 *          the CTA that completes a row block adds the partial sums in column order.
 *   d_workspace: emloco_linear_chain_workspace_ints(...) int32 words, zero before the FIRST launch (the kernel leaves it zeroed);
 *          one workspace per chain that may be in flight.
 * At most 12 layers and 32 segments; all layers that are linked by `dep` have the same M. */
typedef struct {
    const uint16_t* a_hi; const uint16_t* a_lo; int64_t lda;
    const uint16_t* w_hi; const uint16_t* w_lo; int64_t ldw;
    const float* d_bias;
    int64_t M; int32_t N, K;
    int32_t relu, dep;
    float* d_y32; int64_t ldy;
    uint16_t* y_hi; uint16_t* y_lo; int64_t ldy16;
    const float* d_head_w; const float* d_head_bias; float* d_head_part; float* d_head_out;
} emloco_chain_layer;
int64_t emloco_linear_chain_workspace_ints(const emloco_chain_layer* layers, int32_t n_layers);
/* Profiling aid (NULL = off, the default): later emloco_linear_chain launches record, per ticket (= position in `order`), eight
 * int64 words {CTA, layer << 24 | tile, ns claimed, ns rows ready, ns accumulator complete, ns stored, ns the MMA thread reached the
 * tile, ns its first k-block landed} (%globaltimer) into d_trace[8 * tiles]; scripts/chain_trace.py turns them into per-layer occupancy figures. */
int emloco_linear_chain_trace(int64_t* d_trace);
int emloco_linear_chain(const emloco_chain_layer* layers, int32_t n_layers, const int32_t* order, int32_t n_segments,
                        int32_t* d_workspace, int64_t workspace_ints, void* stream);

/* ---- value reuse (optional): `_eval_critic(next obs)` of play_steps (amp_continuous_value.py:85-90) without a second full
 * critic pass.  For an env that is not reset, the next observation IS the observation of the following step, whose policy
 * pass evaluates the critic on it anyway; terminated envs get next_value = 0; only envs reset by the episode time-out
 * (reset && !terminate, ~N/168 per step) need their terminal observation evaluated.
 * emloco_timeout_gather compacts those envs: their actor/critic-input and task-MLP operand rows (bf16 hi/lo, as produced
 * by emloco_set_post_sinks / emloco_split_bf16) are copied to rows 0..count-1 of the c_* buffers, d_idx[i] = env, *d_count.
 * emloco_rollout_record_deferred is emloco_rollout_record with next values handled as described: terminated -> 0 now,
 * timed-out -> from the compact critic output d_c_value_raw[i] (env d_c_idx[i]) now, the others one step later
 * (d_prev_dones / d_prev_next_values = rows of step n-1, completed from this step's d_value_raw; NULL at the first step). */
int emloco_timeout_gather(const int64_t* d_reset, const int64_t* d_terminate, int64_t N, const uint16_t* d_self_hi,
                          const uint16_t* d_self_lo, int64_t ld_self, const uint16_t* d_task_hi, const uint16_t* d_task_lo,
                          int64_t ld_task, uint16_t* d_c_self_hi, uint16_t* d_c_self_lo, int64_t ld_cself, uint16_t* d_c_task_hi,
                          uint16_t* d_c_task_lo, int64_t ld_ctask, int32_t* d_idx, int32_t* d_count, void* stream);
int emloco_rollout_record_deferred(const emloco_rollout_cfg* cfg, const float* d_rew, const int64_t* d_reset,
                                   const int64_t* d_terminate, const float* d_value_raw, const float* d_disc_logit,
                                   const uint8_t* d_inverted, float* d_mb_values, float* d_mb_rewards, float* d_mb_dones,
                                   float* d_mb_next_values, float* d_mb_amp_rewards, float* d_state, int64_t N,
                                   const float* d_c_value_raw, const int32_t* d_c_idx, const int32_t* d_c_count,
                                   const float* d_prev_dones, float* d_prev_next_values, void* stream);

/* the deferred completion on its own (used at the last step of a horizon, which runs the full critic pass itself). */
int emloco_fill_next_values(const emloco_rollout_cfg* cfg, const float* d_value_raw, const float* d_prev_dones,
                            float* d_prev_next_values, int64_t N, void* stream);

/* RunningMeanStd.forward, eval branch (pacer/pacer/utils/running_mean_std.py:82-84): y = clamp((x-mean)/sqrt(var+eps), +-5)
 * on a [M,K] slice with row strides ldx/ldy (the self-obs part of the actor/critic input, amp_network_sept_builder.py:75,95). */
int emloco_normalize(const float* d_x, int64_t ldx, float* d_y, int64_t ldy, int64_t M, int32_t K, const float* d_mean,
                     const float* d_var, float eps, void* stream);

/* ---- inference loop: `AMPPlayerContinuousValue.run` (pacer/pacer/learning/amp_value_players.py:123-198), all envs at once ----
 * Per env and control step: the LocoVal score of the episode is taken at its first step (:128-137); the discounted return
 * cr += ((r_loc + r_pow) * 0.5 + r_disc * 0.25) * gamma^(n+1) (plot_val_reward, :144-160) or cr += r * gamma^(n+1) with the
 * inversion penalty (:127,:161-163) is accumulated; cr is snapshotted at n == step_to_pred or when the episode ends earlier
 * (:177-193).  Every episode that ends appends 8 floats {env, pred, cr_to_pred, (cr_to_pred - min) / (max - min), c_loc, c_pow,
 * c_disc, steps} to d_results (slot = atomic increment of *d_count; entries beyond `capacity` are dropped but counted) - the
 * pairs of the value / return correlation (:263-279).  d_state [11,N] f32 (zero-initialised, row 2 = discount coefficients = 1):
 * n, cr, coef, pred, cr_to_pred, c_loc, c_pow, c_disc and the three component snapshots.  d_rew_raw [N,2] = task.reward_raw. */
int emloco_player_record(const float* d_rew, const float* d_rew_raw, const int64_t* d_reset, const float* d_disc_logit,
                         const float* d_locoval_scores, const uint8_t* d_inverted, float* d_state, int64_t N, float* d_results,
                         int32_t* d_count, int32_t capacity, int32_t plot_val_reward, float inversion_penalty_scale,
                         float disc_reward_scale, float gamma, int32_t step_to_pred, float min_reward, float max_reward, void* stream);

/* ---- mocap reset state and AMP demo observations (reference pacer/pacer/utils/motion_lib_smpl.py:485-563,596-614
 * `MotionLibSMPL.get_motion_state_smpl`, env/tasks/humanoid_amp.py:168-220 `fetch_amp_obs_demo` / `build_amp_obs_demo`) ----
 * The motion library as MotionLibSMPL keeps it on the device (:248-330): per FRAME (all clips concatenated) global translations
 * gts [F,24,3], global / local rotations grs / lrs [F,24,4] (xyzw), global linear / angular velocities gvs / gavs [F,24,3], joint
 * velocities dvs [F,23,3]; per MOTION its length (s), frame period, frame count, index of its first frame and shape parameters
 * [17] (gender + betas).  All device pointers, fp32 / int32. */
typedef struct emloco_motion_lib {
    const float *d_gts, *d_grs, *d_lrs, *d_gvs, *d_gavs, *d_dvs;
    const float *d_length, *d_dt, *d_bodies;
    const int32_t *d_num_frames, *d_start;
    int32_t num_motions, reserved;
} emloco_motion_lib;
/* get_motion_state_smpl for n (motion id, time) pairs: frame pair + blend (:596-606), lerp / slerp, local rotations -> exp-map
 * DOFs.  Outputs (any may be NULL) in the layouts the sim consumes: d_root_state [n,13] (pos, rot xyzw, lin vel, ang vel - the
 * rows of emloco_reset_done's d_init_root), d_dof_state [n,69,2] (pos, vel - d_init_dof), d_key_pos [n,4,3] (R/L ankle, R/L
 * wrist), d_rb_state [n,24,13] (rg_pos, rb_rot, body_vel, body_ang_vel). */
int emloco_motion_state(const emloco_motion_lib* lib, const int32_t* d_motion_ids, const float* d_motion_times, int64_t n,
                        float* d_root_state, float* d_dof_state, float* d_key_pos, float* d_rb_state, void* stream);
/* build_amp_obs_demo: for each of n samples `num_steps` AMP observations (206 floats: build_amp_observations_smpl,
 * humanoid_amp.py:917-971) at times t0 - k * dt, newest first -> d_amp_obs [n, num_steps * 206]. */
int emloco_amp_obs_demo(const emloco_motion_lib* lib, const int32_t* d_motion_ids, const float* d_motion_times0, int64_t n,
                        int32_t num_steps, float dt, float* d_amp_obs, void* stream);

/* ---- PPO / AMP update step: `AMPValueAgent.calc_gradients`, pacer/pacer/learning/amp_continuous_value.py:276-428
 * (losses: learning/common_agent.py:594-602,657-683, amp_continuous_value.py:430-444, amp_continuous.py:536-616; optimiser:
 * common_agent.py:84-87 torch.optim.Adam + nn.utils.clip_grad_norm_(grad_norm 50); multi-GPU: Horovod `optimizer.synchronize()`
 * :386-394, replaced by one NCCL all-reduce over a flat gradient buffer).  The dense layers of forward, dgrad, wgrad and of the
 * gradient-penalty double backward run on emloco_linear_bf16x3; the entries below are everything else (csrc/update.cu).
 *
 * emloco_xform: y[m,k] = scale * rowscale[m] * src[m,k] * (gate[m,k] > 0), src = clamp((x-mean)/sqrt(var+eps),+-5) | x | rowvec[k]
 * (x == NULL).  Any subset of outputs: fp32 y [M,K] and y^T [K,M]; bf16 hi/lo split of y and of y^T (the A / W operands of
 * the GEMMs: activations and weight transposes for dgrad, transposed activations and output gradients for wgrad); colsum[k] +=
 * sum_m y (bias gradients); *sumsq += sum y^2 (the gradient penalty).  d_drop_u [M,19] (or NULL): a second gate, the whole-joint
 * dropout of AMP-observation columns (learning/amp_models.py:49-90) evaluated inline from the per-joint uniform draws - column k
 * of joint j passes where d_drop_u[m,j] > drop_rate (K must be a multiple of 206); no mask tensor is materialised.
 * Pointers NULL when unused. */
int emloco_xform(const float* d_x, int64_t ldx, const float* d_rowvec, const float* d_rowscale, int64_t lds, const float* d_mean,
                 const float* d_var, float eps, const float* d_gate, int64_t ldg, const float* d_drop_u, float drop_rate, float scale,
                 float* d_y32, int64_t ldy, float* d_yT32, int64_t ldyT, uint16_t* d_hi, uint16_t* d_lo, int64_t ld16, uint16_t* d_hiT,
                 uint16_t* d_loT, int64_t ldT, float* d_colsum, float* d_sumsq, int64_t M, int32_t K, void* stream);
/* actor (PPO clip), critic, task-value and bound losses -> gradients of the weighted, batch-averaged total loss w.r.t. mu [B,A],
 * value [B], task value [B]; d_stats[7] += sums of {actor loss, critic loss, task-value loss, bound loss, clipped, kl, entropy}. */
int emloco_ppo_heads(const float* d_mu, int64_t ldmu, const float* d_logstd, const float* d_actions, const float* d_old_neglogp,
                     const float* d_adv, const float* d_value, const float* d_task_value, const float* d_returns,
                     const float* d_old_mu, const float* d_old_sigma, float* d_dmu, int64_t lddmu, float* d_dvalue, float* d_dtask,
                     float* d_stats, int64_t B, int32_t A, float e_clip, float actor_coef, float critic_coef, float tv_coef,
                     float bounds_coef, void* stream);
/* discriminator prediction loss 0.5 * (BCE(fake, 0) + BCE(real, 1)) * coef -> d loss / d logit; rows [0, n_agent) are the agent
 * + replay logits, [n_agent, n_agent + n_demo) the demo logits; d_stats[4] += {sum softplus(fake), sum softplus(-real), #fake < 0,
 * #real > 0}. */
int emloco_disc_heads(const float* d_logit, float* d_dlogit, float* d_stats, int64_t n_agent, int64_t n_demo, float coef, void* stream);
/* whole-joint dropout masks of the AMP observations (learning/amp_models.py:49-90): u [rows,19] uniform draws -> mask [rows,3090]. */
int emloco_amp_dropout_mask(const float* d_u, float* d_mask, int64_t rows, float rate, void* stream);
/* RunningMeanStd training-mode update (utils/running_mean_std.py:33-43,86-96) of float64 running_mean / running_var / count from
 * the batch x [M,K]; also refreshes the fp32 copies mean / var / 1/sqrt(var+eps) (may be NULL).  d_scratch: 2*K doubles, zero on
 * entry, left zero. */
int emloco_rms_update(const float* d_x, int64_t ldx, int64_t M, int32_t K, double* d_scratch, double* d_running_mean,
                      double* d_running_var, double* d_count, float* d_mean32, float* d_var32, float* d_inv_std32, float eps,
                      void* stream);
/* clip_grad_norm_ + Adam over flat buffers.  d_state[2] = {step count, sum of squares of the gradient}: emloco_adam_begin
 * increments the step and clears the sum, emloco_grad_sumsq accumulates it (after the all-reduce, on the SUMMED gradient;
 * deterministic two-stage reduction through d_partials [>= 1184] so that every rank computes the same norm to the last bit),
 * emloco_adam_clip scales the gradient by grad_scale (1 / world size) and min(1, max_norm / (norm + 1e-6)) and steps. */
int emloco_adam_begin(float* d_state, void* stream);
int emloco_grad_sumsq(const float* d_grad, int64_t n, float* d_state, float* d_partials, void* stream);
int emloco_adam_clip(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float* d_state, float lr, float beta1,
                     float beta2, float eps, float max_norm, float grad_scale, void* stream);
/* The optimiser step FUSED WITH ITS COLLECTIVE over NVLink / NVSwitch multicast memory (replaces all-reduce + norm + Adam on every
 * rank; reference: Horovod `optimizer.synchronize()` + clip_grad_norm_ + Adam, amp_continuous_value.py:381-398).  mc_* are
 * MULTICAST addresses of buffers every rank allocated symmetrically (flat gradient, flat parameters, exchange [world x 4] floats);
 * rank r owns the slice [lo, lo + count) of the flat vector (multiples of 4 floats).
 * emloco_dp_reduce_shard: `multimem.ld_reduce` adds the ranks' gradients of the slice inside the switch -> d_shard_grad [count];
 *   the slice's sum of squares is broadcast (`multimem.st`) into slot `rank` of every rank's exchange buffer.
 * emloco_dp_adam_shard: sums the `world` exchange slots in rank order, clips (max_norm on the gradient scaled by grad_scale =
 *   1 / world), runs Adam on the slice with THIS rank's moments d_m / d_v [count] (optimiser state sharded over ranks) and
 *   broadcasts the new parameters into every rank's parameter buffer (`multimem.st`).  d_state = {step, total sum of squares}.
 * The caller orders the ranks with symmetric-memory barriers before, between and after the two calls. */
int emloco_dp_reduce_shard(const float* mc_grad, float* d_shard_grad, int64_t lo, int64_t count, float* d_partials, float* mc_exchange,
                           int32_t rank, void* stream);
int emloco_dp_adam_shard(float* mc_param, const float* d_param_local, const float* d_shard_grad, float* d_m, float* d_v, int64_t lo, int64_t count,
                         const float* d_exchange_local, int32_t world, float* d_state, float lr, float beta1, float beta2, float eps,
                         float max_norm, float grad_scale, void* stream);
/* y += a * x (weight-decay / logit-regularisation terms of the discriminator loss, amp_continuous.py:548-550,585-589). */
int emloco_axpy(float* d_y, const float* d_x, float a, int64_t n, void* stream);
/* out[i] = sum_s parts[s * part_stride + i] (+ out[i] when accumulate != 0), in split order: the consumer of a split-K
 * emloco_linear_bf16x3 launch (weight gradients whose output has too few tiles to fill the SMs; the contraction runs over the
 * whole minibatch, so it is split instead). */
int emloco_sum_parts(const float* d_parts, int32_t num_parts, int64_t part_stride, float* d_out, int64_t n, int32_t accumulate, void* stream);

int emloco_sync(emloco_sim* sim);
const char* emloco_last_error(void);
const char* emloco_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EMLOCO_H */
