// Internal (not part of the C ABI): simulation object and kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emloco.h"
#include "common.cuh"

// Per-asset constants, uploaded once to __constant__ memory (shared by all envs: the default
// run has no shape variation because the SMPL model files are not redistributable).
struct EmlModelDev {
    int   parent[EML_NB];
    int   level[EML_NB];
    int   child[EML_NB][3];       // up to 3 children per body, -1 padded
    float offset[EML_NB][3];      // joint anchor in the parent frame
    float mass[EML_NB];
    float com[EML_NB][3];         // body frame
    float inertia[EML_NB][6];     // about COM, body frame: xx xy xz yy yz zz
    float kp[EML_NB];             // per joint (index = body), already scaled by mass/77*kp_scale
    float kd[EML_NB];
    float arm[EML_NB];
    int   geom_type[EML_NB];
    float geom_a[EML_NB][3];
    float geom_b[EML_NB][3];
    float geom_r[EML_NB];
    float pd_offset[EML_ND];
    float pd_scale[EML_ND];
    float geom_bound[EML_NB];     // largest distance from the body origin to a contact point, + the primitive's radius
    int   max_level;
};

// Per-env body models (has_shape_variation - every env simulates the body generated from its own SMPL shape
// parameters, humanoid.py:597-739; PD gains scaled by the body's mass, :905-910).  Optional device array [EM_FLOATS][N]
// (field-major, env-minor: the lane-per-env physics kernel reads it coalesced); when absent all envs share EmlModelDev.
// Topology, geometry TYPES and the action -> PD-target map are shape independent and stay in EmlModelDev.
enum { EM_OFFSET = 0, EM_MASS = 72, EM_COM = 96, EM_INERTIA = 168, EM_KP = 312, EM_KD = 336, EM_ARM = 360, EM_GA = 384, EM_GB = 456,
       EM_GR = 528, EM_BOUND = 552, EM_FLOATS = 576 };

struct emloco_sim {
    emloco_cfg cfg;
    int N;
    int device;
    int physics_impl;       // 0: lane-per-env kernel (physics_soa.cu), 1: warp-per-env kernel (physics.cu)
    float* env_model;       // [EM_FLOATS][N] per-env body models, or NULL (shared model)
    // --- state owned by the sim; addresses are stable for its lifetime (gymtorch.wrap_tensor aliases) ---
    float*   root_state;    // [N,13]   pos3 quat4 lin3 ang3
    float*   dof_state;     // [N*69,2] (exp-map pos, vel) interleaved
    float*   rb_state;      // [N*24,13]
    float*   contact;       // [N*24,3]
    float*   dof_force;     // [N*69]
    float*   pd_target;     // [N,69]
    float*   joint_quat;    // [N,23,4] internal authoritative joint rotation
    float*   actions;       // [N,69]  (copy of the last actions, humanoid.py:1186)
    // --- task buffers (base_task.py:96-112) ---
    float*   obs;           // [N,1422]
    float*   flip_obs;      // [N,1422]
    float*   rew;           // [N]
    float*   rew_raw;       // [N,2]
    int64_t* reset;         // [N]
    int64_t* terminate;     // [N]
    int64_t* progress;      // [N]
    float*   amp_obs;       // [N,15,206]
    float*   verts;         // [N,101,3] trajectory polylines (TrajGenerator._verts)
    float*   betas;         // [N,17]
    int16_t* height;        // [rows,cols]
    int      hf_rows, hf_cols;
    float    hf_max;        // highest terrain sample in metres (bounding test of the contact loops)
    // pinned host staging for the *_host entry points
    float*   h_pin;
    size_t   h_pin_bytes;
    cudaStream_t copy_stream;
    emloco_post_sinks sinks;   // optional extra outputs of the post-step kernel (all NULL by default)
    const float** ring_ptr;    // [N] where each env's AMP ring currently lives (its amp_obs row or an experience row)
    uint32_t* traj_epoch;      // [N] per-env count of device-side trajectory resets (Philox counter word)
    emloco_traj_cfg traj;      // stage appended to emloco_reset_done when traj_on
    int traj_on;
};

struct RecordParams {
    // inputs of this step
    const float* rew;            // [N] env reward (rew_buf)
    const int64_t* reset;        // [N] dones
    const int64_t* terminate;    // [N]
    const float* value_raw;      // [N] critic on the pre-step obs (get_action_values), normalised units, or NULL
    const float* next_value_raw; // [N] critic on next obs, normalised units
    const float* disc_logit;     // [N]
    const uint8_t* inverted;     // [N] or NULL (task.inverted)
    // experience-buffer rows of step n (each [N])
    float* mb_values; float* mb_rewards; float* mb_dones; float* mb_next_values; float* mb_amp_rewards;
    // persistent per-env state
    float* current_rewards; float* current_lengths; float* current_combined; float* discount_coefs; float* game_combined;
    float* terminated_flags;
    long long N;
    float inv_penalty, reward_scale, v_mean, v_std, disc_scale, gamma, step_to_pred;
    int unnorm_value;
    const float* v_stats;        // device {mean, std} overriding v_mean / v_std when non-NULL (graph-safe value_mean_std updates)
    // deferred next-values (value reuse): next_value_raw == NULL.  Terminated envs get 0 now, timed-out envs (reset, not
    // terminated) get the compact critic's value now, the others are filled one step later from that step's value_raw
    const float* c_value_raw; const int32_t* c_idx; const int32_t* c_count;   // compact critic outputs
    const float* prev_dones; float* prev_next_values;                          // step n-1 rows to complete (or NULL)
};

// kernel launchers (defined in the .cu files)
cudaError_t eml_upload_model(const EmlModelDev* m);
cudaError_t eml_launch_post_step(emloco_sim* s, int advance_progress, cudaStream_t st);
cudaError_t eml_launch_physics(emloco_sim* s, const float* d_actions, int n_substeps, int fuse_post, cudaStream_t st);
cudaError_t eml_launch_fk(emloco_sim* s, const int32_t* d_env_ids, int n, cudaStream_t st);
cudaError_t eml_reset_done(emloco_sim* s, const float* d_init_root, const float* d_init_dof, cudaStream_t st);
cudaError_t eml_launch_post_reset(emloco_sim* s, int keep_flags, cudaStream_t st);
cudaError_t eml_traj_reset(emloco_sim* s, const emloco_traj_cfg& c, int clear_flags, cudaStream_t st);
cudaError_t eml_sample_actions(const float* mu, long long ldmu, const float* logstd, const float* noise, float* actions,
                               float* neglogp, long long N, int A, cudaStream_t st);
cudaError_t eml_sample_actions_parts(const float* mu_parts, long long ldmu, int parts, long long part_stride, float* mu_out, long long ldout,
                                     const float* logstd, const float* noise, float* actions, float* neglogp, long long N, int A,
                                     cudaStream_t st);
cudaError_t eml_disc_reward(const float* logit, const float* task_rew, float* disc, float* combined, long long M, float scale,
                            float w_task, float w_disc, cudaStream_t st);
cudaError_t eml_rollout_record(const RecordParams& P, cudaStream_t st);
cudaError_t eml_normalize(const float* x, long long ldx, float* y, long long ldy, long long M, int K, const float* mean,
                          const float* var, float eps, cudaStream_t st);
cudaError_t eml_split_bf16(const float* x, long long ldx, long long M, int K, const float* mean, const float* var, float eps,
                           void* hi, void* lo, long long ld16, cudaStream_t st);
cudaError_t eml_linear_bf16x3(const void* a_hi, const void* a_lo, long long lda, const void* w_hi, const void* w_lo, long long ldw,
                              const float* bias, long long M, int N, int K, int relu, float* y32, long long ldy, void* y_hi,
                              void* y_lo, long long ldy16, int tile_n, const int* m_dev, const float* head_w, float* head_part,
                              int head_ld, cudaStream_t st);
cudaError_t eml_linear_chain(const emloco_chain_layer* layers, int n_layers, const int* order, int n_segments, int* ws, long long ws_ints,
                             cudaStream_t st, const char** why);
void eml_linear_chain_trace(long long* buf);
long long eml_linear_chain_workspace_ints(const emloco_chain_layer* layers, int n_layers);
cudaError_t eml_head_reduce(const float* part, int groups, const float* bias, float* out, long long M, cudaStream_t st);
cudaError_t eml_timeout_gather(const int64_t* reset, const int64_t* terminate, long long N, const uint16_t* self_hi,
                               const uint16_t* self_lo, long long ld_self, const uint16_t* task_hi, const uint16_t* task_lo,
                               long long ld_task, uint16_t* c_self_hi, uint16_t* c_self_lo, long long ld_cself, uint16_t* c_task_hi,
                               uint16_t* c_task_lo, long long ld_ctask, int32_t* idx, int32_t* count, cudaStream_t st);
cudaError_t eml_fill_next_values(const float* value_raw, const float* prev_dones, float* prev_next_values, long long N, float v_mean,
                                 float v_std, const float* v_stats, int unnorm, cudaStream_t st);
