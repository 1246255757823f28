// C ABI of libemloco_b200.so (see include/emloco.h for the reference interface each entry replaces).
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "sim.h"

cudaError_t eml_locoval_forward(const float*, int, int, float*, const float*, const float*, float*, long long, int, cudaStream_t);
cudaError_t eml_locoval_backward(const float*, int, int, const float*, const float*, const float*, const float*, float*, const float*, float*, long long, int, cudaStream_t);
size_t eml_locoval_train_workspace_bytes(long long N);
cudaError_t eml_locoval_train_step(const float* traj, int stride, int T, const float* pose, const float* vel, float* gc, float* w,
                                   float* m, float* v, float* step, float* stats, void* workspace, long long N, float lr, float beta1,
                                   float beta2, float eps, float wd, float r_min, float r_max, int flags, cudaStream_t st);

cudaError_t eml_plausibl_forward(const float*, const float*, float*, long long, cudaStream_t);
cudaError_t eml_gae(const float*, const float*, const float*, const float*, float*, float*, int, long long, float, float, cudaStream_t);
cudaError_t eml_linear_fma(const float*, long long, const float*, const float*, float*, long long, long long, int, int,
                           const float*, const float*, float, int, cudaStream_t);
cudaError_t eml_linear_tc(const float*, long long, const float*, const float*, float*, long long, long long, int, int,
                          const float*, const float*, float, int, cudaStream_t);

cudaError_t eml_xform(const float* x, long long ldx, const float* rowvec, const float* rowscale, long long lds, const float* mean,
                      const float* var, float eps, const float* gate, long long ldg, const float* drop_u, float drop_rate, float scale,
                      float* y32, long long ldy, float* yT32, long long ldyT, void* hi, void* lo, long long ld16, void* hiT, void* loT,
                      long long ldT, float* colsum, float* sumsq, long long M, int K, cudaStream_t st);
cudaError_t eml_ppo_heads(const float* mu, long long ldmu, const float* logstd, const float* actions, const float* old_neglogp,
                          const float* adv, const float* value, const float* task_value, const float* returns, const float* old_mu,
                          const float* old_sigma, float* dmu, long long lddmu, float* dvalue, float* dtask, float* stats, long long B,
                          int A, float e_clip, float actor_coef, float critic_coef, float tv_coef, float bounds_coef, cudaStream_t st);
cudaError_t eml_disc_heads(const float* logit, float* dlogit, float* stats, long long n_agent, long long n_demo, float coef, cudaStream_t st);
cudaError_t eml_amp_dropout_mask(const float* u, float* mask, long long rows, float rate, cudaStream_t st);
cudaError_t eml_rms_update(const float* x, long long ldx, long long M, int K, double* scratch, double* rmean, double* rvar, double* count,
                           float* mean32, float* var32, float* inv32, float eps, cudaStream_t st);
cudaError_t eml_grad_sumsq(const float* g, long long n, float* state, float* partials, cudaStream_t st);
cudaError_t eml_adam_clip(float* p, const float* g, float* m, float* v, long long n, float* state, float lr, float beta1, float beta2,
                          float eps, float max_norm, float grad_scale, cudaStream_t st);
cudaError_t eml_adam_begin(float* state, cudaStream_t st);
cudaError_t eml_dp_reduce_shard(const float* mc_grad, float* shard, long long lo, long long count, float* partials, float* mc_exchange,
                                int rank, cudaStream_t st);
cudaError_t eml_dp_adam_shard(float* mc_param, const float* p_local, const float* shard, float* m, float* v, long long lo, long long count,
                              const float* exchange, int world, float* state, float lr, float beta1, float beta2, float eps, float max_norm,
                              float grad_scale, cudaStream_t st);
cudaError_t eml_axpy(float* y, const float* x, float a, long long n, cudaStream_t st);
cudaError_t eml_sum_parts(const float* parts, int S, long long stride, float* out, long long n, int accumulate, cudaStream_t st);

cudaError_t eml_player_record(const float* rew, const float* rew_raw, const int64_t* reset, const float* logit, const float* scores,
                              const uint8_t* inverted, float* st, long long N, float* results, int* count, int capacity,
                              int plot_val_reward, float inv_penalty, float disc_scale, float gamma, int step_to_pred, float min_reward,
                              float max_reward, cudaStream_t stream);

struct EmlMotionLibDev {
    const float *gts, *grs, *lrs, *gvs, *gavs, *dvs;
    const float *length, *dt, *bodies;
    const int *num_frames, *start;
    int num_motions;
};
cudaError_t eml_motion_state(const EmlMotionLibDev& L, const int* ids, const float* times, long long n, float* root, float* dof,
                             float* key_pos, float* rb, cudaStream_t st);
cudaError_t eml_amp_obs_demo(const EmlMotionLibDev& L, const int* ids, const float* times0, long long n, int steps, float dt, float* out,
                             cudaStream_t st);

static thread_local std::string g_err;
static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}
#define CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(EMLOCO_ECUDA, what, _e); } while (0)

extern "C" {

const char* emloco_last_error(void) { return g_err.c_str(); }
const char* emloco_version(void) { return "emloco_b200 0.1 (sm_100a)"; }

void emloco_default_cfg(emloco_cfg* c) {
    memset(c, 0, sizeof *c);
    c->num_envs = 4096; c->device = 0;
    c->sim_dt = 1.0f / 60.0f; c->substeps = 2; c->control_freq_inv = 2;      // config.py:24, pacer.yaml:94,42
    c->gravity_z = -9.81f;                                                  // base_task.py:225-231
    c->contact_stiffness = 5.0e4f; c->contact_damping = 1.0e3f; c->friction_damping = 2.0e3f;
    c->friction_mu = 1.0f;                                                  // pacer.yaml:71-72
    c->contact_offset = 0.02f; c->max_ang_vel = 100.0f; c->angular_damping = 0.01f;
    c->episode_length = 168; c->power_coefficient = 0.0005f; c->location_coefficient = 1.0f;
    c->fail_dist = 4.0f; c->traj_sample_dt = 0.4f;
    c->max_effort = 500.0f;                                                 // smpl_humanoid.xml:171-233 motor gear
    c->max_turn = 0.3f;
}

int emloco_create(const emloco_cfg* cfg, const emloco_model* model, emloco_sim** out) {
    if (!cfg || !model || !out) return fail(EMLOCO_EINVAL, "emloco_create: null argument");
    if (cfg->num_envs <= 0) return fail(EMLOCO_EINVAL, "emloco_create: num_envs must be positive");
    if (cfg->substeps <= 0 || cfg->control_freq_inv <= 0 || cfg->sim_dt <= 0) return fail(EMLOCO_EINVAL, "emloco_create: bad time stepping");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(EMLOCO_EINVAL, "emloco_create: no such CUDA device");
    CK(cudaSetDevice(cfg->device), "cudaSetDevice");

    EmlModelDev md;
    memset(&md, 0, sizeof md);
    int nchild[EML_NB] = {0};
    for (int i = 0; i < EML_NB; ++i) for (int s = 0; s < 3; ++s) md.child[i][s] = -1;
    md.max_level = 0;
    for (int i = 0; i < EML_NB; ++i) {
        int p = model->parent[i];
        if (i == 0 ? p != -1 : (p < 0 || p >= i)) return fail(EMLOCO_EINVAL, "emloco_create: bodies must be in DFS order with parent < child");
        md.parent[i] = p;
        md.level[i] = i == 0 ? 0 : md.level[p] + 1;
        if (md.level[i] > md.max_level) md.max_level = md.level[i];
        if (i > 0) {
            if (nchild[p] >= 3) return fail(EMLOCO_EINVAL, "emloco_create: more than 3 children per body");
            md.child[p][nchild[p]++] = i;
            int d = 3 * (i - 1);
            md.kp[i] = model->kp[d]; md.kd[i] = model->kd[d]; md.arm[i] = model->armature[d];
            for (int k = 1; k < 3; ++k)
                if (model->kp[d + k] != model->kp[d] || model->kd[d + k] != model->kd[d] || model->armature[d + k] != model->armature[d])
                    return fail(EMLOCO_EINVAL, "emloco_create: x/y/z gains of a joint must be equal (spherical-joint formulation)");
        }
        md.mass[i] = model->mass[i];
        md.geom_type[i] = model->geom_type[i]; md.geom_r[i] = model->geom_r[i];
        {   // bound of the body's contact points (physics_soa.cu skips the contact loop of bodies that cannot reach the ground)
            const float* a = model->geom_a[i]; const float* b = model->geom_b[i];
            const int gt = model->geom_type[i];
            float ext;
            if (gt == 0) ext = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) + model->geom_r[i];
            else if (gt == 1) ext = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2])) + model->geom_r[i];
            else { const float cx = fabsf(a[0]) + fabsf(b[0]), cy = fabsf(a[1]) + fabsf(b[1]), cz = fabsf(a[2]) + fabsf(b[2]); ext = sqrtf(cx * cx + cy * cy + cz * cz); }
            md.geom_bound[i] = ext * 1.0001f + 1e-5f;
        }
        for (int k = 0; k < 3; ++k) {
            md.offset[i][k] = model->offset[i][k]; md.com[i][k] = model->com[i][k];
            md.geom_a[i][k] = model->geom_a[i][k]; md.geom_b[i][k] = model->geom_b[i][k];
        }
        for (int k = 0; k < 6; ++k) md.inertia[i][k] = model->inertia[i][k];
    }
    // the level-synchronous child->parent reduction assumes multi-child bodies sit at levels 0 and 3 only
    for (int i = 0; i < EML_NB; ++i)
        if (nchild[i] > 1 && md.level[i] != 0 && md.level[i] != 3)
            return fail(EMLOCO_EINVAL, "emloco_create: unsupported tree shape (branching body not at level 0 or 3)");
    for (int d = 0; d < EML_ND; ++d) { md.pd_offset[d] = model->pd_offset[d]; md.pd_scale[d] = model->pd_scale[d]; }
    CK(eml_upload_model(&md), "upload model");

    emloco_sim* s = (emloco_sim*)calloc(1, sizeof(emloco_sim));
    if (!s) return fail(EMLOCO_ENOMEM, "emloco_create: out of host memory");
    s->cfg = *cfg; s->N = cfg->num_envs; s->device = cfg->device;
    {   // the lane-per-env kernel hard-codes the SMPL kinematic chains (legs, spine + head, arms)
        static const int smpl_parent[EML_NB] = {-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22};
        bool smpl = true;
        for (int i = 0; i < EML_NB; ++i) smpl &= model->parent[i] == smpl_parent[i];
        if (cfg->physics_impl != 0 && cfg->physics_impl != 1) { free(s); return fail(EMLOCO_EINVAL, "emloco_create: physics_impl must be 0 or 1"); }
        s->physics_impl = smpl ? cfg->physics_impl : 1;
    }
    const size_t N = (size_t)s->N;
#define ALLOC(ptr, count, type) do { CK(cudaMalloc((void**)&(ptr), (count) * sizeof(type)), "cudaMalloc " #ptr); \
                                     CK(cudaMemset((ptr), 0, (count) * sizeof(type)), "cudaMemset " #ptr); } while (0)
    ALLOC(s->root_state, N * 13, float); ALLOC(s->dof_state, N * EML_ND * 2, float); ALLOC(s->rb_state, N * EML_NB * 13, float);
    ALLOC(s->contact, N * EML_NB * 3, float); ALLOC(s->dof_force, N * EML_ND, float); ALLOC(s->pd_target, N * EML_ND, float);
    ALLOC(s->joint_quat, N * EML_NJ * 4, float); ALLOC(s->actions, N * EML_ND, float);
    ALLOC(s->obs, N * EML_OBS, float); ALLOC(s->flip_obs, N * EML_OBS, float); ALLOC(s->rew, N, float); ALLOC(s->rew_raw, N * 2, float);
    ALLOC(s->reset, N, int64_t); ALLOC(s->terminate, N, int64_t); ALLOC(s->progress, N, int64_t);
    ALLOC(s->amp_obs, N * EML_AMP_OBS, float); ALLOC(s->verts, N * EML_NUM_VERTS * 3, float); ALLOC(s->betas, N * 17, float);
    ALLOC(s->traj_epoch, N, uint32_t); ALLOC(s->ring_ptr, N, const float*);
    {
        std::vector<const float*> h(N);
        for (size_t i = 0; i < N; ++i) h[i] = s->amp_obs + i * EML_AMP_OBS;
        CK(cudaMemcpy(s->ring_ptr, h.data(), N * sizeof(const float*), cudaMemcpyHostToDevice), "init ring_ptr");
    }
    // default terrain: flat 1080 x 1080 (8 m map + 50 m border at 0.1 m, humanoid_pedestrain_terrain.py:1142-1165)
    s->hf_rows = 1080; s->hf_cols = 1080;
    ALLOC(s->height, (size_t)s->hf_rows * s->hf_cols, int16_t);
#undef ALLOC
    s->h_pin_bytes = N * (EML_ND + EML_OBS + EML_AMP_OBS + 1) * sizeof(float) + N * sizeof(int64_t);
    CK(cudaMallocHost((void**)&s->h_pin, s->h_pin_bytes), "cudaMallocHost");
    // identity pose: unit quaternions
    {
        float* h = (float*)malloc(N * 13 * sizeof(float));
        memset(h, 0, N * 13 * sizeof(float));
        for (size_t e = 0; e < N; ++e) h[e * 13 + 6] = 1.f;
        CK(cudaMemcpy(s->root_state, h, N * 13 * sizeof(float), cudaMemcpyHostToDevice), "init root");
        free(h);
    }
    CK(eml_launch_fk(s, nullptr, 0, 0), "initial forward kinematics");
    CK(cudaDeviceSynchronize(), "emloco_create sync");
    *out = s;
    return EMLOCO_OK;
}

int emloco_destroy(emloco_sim* s) {
    if (!s) return EMLOCO_OK;
    cudaSetDevice(s->device);
    void* ptrs[] = {s->root_state, s->dof_state, s->rb_state, s->contact, s->dof_force, s->pd_target, s->joint_quat, s->actions,
                    s->obs, s->flip_obs, s->rew, s->rew_raw, s->reset, s->terminate, s->progress, s->amp_obs, s->verts, s->betas,
                    s->height, s->traj_epoch, (void*)s->ring_ptr, s->env_model};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    free(s);
    return EMLOCO_OK;
}

int emloco_set_env_models(emloco_sim* s, const float* h_env_models) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_set_env_models: null sim");
    CK(cudaSetDevice(s->device), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "sync before replacing the body models");
    if (!h_env_models) {                                      // back to the shared model
        if (s->env_model) cudaFree(s->env_model);
        s->env_model = nullptr;
        return EMLOCO_OK;
    }
    const size_t N = (size_t)s->N;
    std::vector<float> t((size_t)EM_FLOATS * N);              // env-major [N][576] -> field-major [576][N]
    for (size_t e = 0; e < N; ++e) {
        const float* m = h_env_models + e * EM_FLOATS;
        for (int b = 0; b < EML_NB; ++b)
            if (!(m[EM_MASS + b] > 0.f)) return fail(EMLOCO_EINVAL, "emloco_set_env_models: body mass must be positive");
        for (int f = 0; f < EM_FLOATS; ++f) t[(size_t)f * N + e] = m[f];
    }
    if (!s->env_model) CK(cudaMalloc(&s->env_model, t.size() * sizeof(float)), "cudaMalloc env models");
    CK(cudaMemcpy(s->env_model, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice), "upload env models");
    return EMLOCO_OK;
}

int emloco_tensor(emloco_sim* s, int which, void** d_ptr, int64_t* shape, int32_t* ndim, int32_t* dtype) {
    if (!s || !d_ptr || !shape || !ndim || !dtype) return fail(EMLOCO_EINVAL, "emloco_tensor: null argument");
    const int64_t N = s->N;
    *dtype = EMLOCO_DTYPE_F32;
    shape[0] = shape[1] = shape[2] = shape[3] = 1;
    switch (which) {
    case EMLOCO_T_ROOT_STATE: *d_ptr = s->root_state; *ndim = 2; shape[0] = N; shape[1] = 13; break;
    case EMLOCO_T_DOF_STATE: *d_ptr = s->dof_state; *ndim = 2; shape[0] = N * EML_ND; shape[1] = 2; break;
    case EMLOCO_T_RB_STATE: *d_ptr = s->rb_state; *ndim = 2; shape[0] = N * EML_NB; shape[1] = 13; break;
    case EMLOCO_T_CONTACT: *d_ptr = s->contact; *ndim = 2; shape[0] = N * EML_NB; shape[1] = 3; break;
    case EMLOCO_T_DOF_FORCE: *d_ptr = s->dof_force; *ndim = 1; shape[0] = N * EML_ND; break;
    case EMLOCO_T_PD_TARGET: *d_ptr = s->pd_target; *ndim = 2; shape[0] = N; shape[1] = EML_ND; break;
    case EMLOCO_T_OBS: *d_ptr = s->obs; *ndim = 2; shape[0] = N; shape[1] = EML_OBS; break;
    case EMLOCO_T_FLIP_OBS: *d_ptr = s->flip_obs; *ndim = 2; shape[0] = N; shape[1] = EML_OBS; break;
    case EMLOCO_T_REW: *d_ptr = s->rew; *ndim = 1; shape[0] = N; break;
    case EMLOCO_T_REW_RAW: *d_ptr = s->rew_raw; *ndim = 2; shape[0] = N; shape[1] = 2; break;
    case EMLOCO_T_RESET: *d_ptr = s->reset; *ndim = 1; shape[0] = N; *dtype = EMLOCO_DTYPE_I64; break;
    case EMLOCO_T_TERMINATE: *d_ptr = s->terminate; *ndim = 1; shape[0] = N; *dtype = EMLOCO_DTYPE_I64; break;
    case EMLOCO_T_PROGRESS: *d_ptr = s->progress; *ndim = 1; shape[0] = N; *dtype = EMLOCO_DTYPE_I64; break;
    case EMLOCO_T_AMP_OBS: *d_ptr = s->amp_obs; *ndim = 3; shape[0] = N; shape[1] = EML_AMP_STEPS; shape[2] = EML_AMP_STEP; break;
    case EMLOCO_T_TRAJ_VERTS: *d_ptr = s->verts; *ndim = 3; shape[0] = N; shape[1] = EML_NUM_VERTS; shape[2] = 3; break;
    case EMLOCO_T_BETAS: *d_ptr = s->betas; *ndim = 2; shape[0] = N; shape[1] = 17; break;
    case EMLOCO_T_HEIGHT: *d_ptr = s->height; *ndim = 2; shape[0] = s->hf_rows; shape[1] = s->hf_cols; *dtype = EMLOCO_DTYPE_I16; break;
    case EMLOCO_T_JOINT_QUAT: *d_ptr = s->joint_quat; *ndim = 3; shape[0] = N; shape[1] = EML_NJ; shape[2] = 4; break;
    case EMLOCO_T_ACTIONS: *d_ptr = s->actions; *ndim = 2; shape[0] = N; shape[1] = EML_ND; break;
    default: return fail(EMLOCO_EINVAL, "emloco_tensor: unknown tensor id");
    }
    return EMLOCO_OK;
}

int emloco_set_height_field(emloco_sim* s, const int16_t* h, int32_t rows, int32_t cols) {
    if (!s || !h || rows < 2 || cols < 2) return fail(EMLOCO_EINVAL, "emloco_set_height_field: bad argument");
    CK(cudaSetDevice(s->device), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "sync before terrain swap");
    if ((size_t)rows * cols != (size_t)s->hf_rows * s->hf_cols) {
        int16_t* n = nullptr;
        CK(cudaMalloc((void**)&n, (size_t)rows * cols * sizeof(int16_t)), "cudaMalloc height");
        cudaFree(s->height);
        s->height = n;
    }
    s->hf_rows = rows; s->hf_cols = cols;
    CK(cudaMemcpy(s->height, h, (size_t)rows * cols * sizeof(int16_t), cudaMemcpyHostToDevice), "copy height field");
    int16_t mx = h[0];
    for (size_t i = 1; i < (size_t)rows * cols; ++i) mx = h[i] > mx ? h[i] : mx;
    s->hf_max = (float)mx * 0.005f;
    return EMLOCO_OK;
}

int emloco_set_pd_targets(emloco_sim* s, const float* d_targets, void* stream) {
    if (!s || !d_targets) return fail(EMLOCO_EINVAL, "emloco_set_pd_targets: null argument");
    if (d_targets != s->pd_target)
        CK(cudaMemcpyAsync(s->pd_target, d_targets, (size_t)s->N * EML_ND * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "copy pd targets");
    return EMLOCO_OK;
}

int emloco_simulate(emloco_sim* s, void* stream) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_simulate: null sim");
    CK(eml_launch_physics(s, nullptr, s->cfg.substeps, 0, (cudaStream_t)stream), "physics kernel");
    return EMLOCO_OK;
}

int emloco_reset_indexed(emloco_sim* s, const int32_t* d_env_ids, int32_t n, void* stream) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_reset_indexed: null sim");
    if (d_env_ids && n < 0) return fail(EMLOCO_EINVAL, "emloco_reset_indexed: negative count");
    if (d_env_ids && n == 0) return EMLOCO_OK;
    CK(eml_launch_fk(s, d_env_ids, n, (cudaStream_t)stream), "fk kernel");
    return EMLOCO_OK;
}

int emloco_set_post_sinks(emloco_sim* s, const emloco_post_sinks* k) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_set_post_sinks: null sim");
    if (!k) { memset(&s->sinks, 0, sizeof s->sinks); return EMLOCO_OK; }
    if (k->self_hi && (!k->self_lo || !k->task_hi || !k->task_lo || !k->obs_mean || !k->obs_inv_std || (k->ld_self & 7) || (k->ld_task & 7) ||
                       k->ld_self < EML_SELF_OBS || k->ld_task < EML_TASK_OBS))
        return fail(EMLOCO_EINVAL, "emloco_set_post_sinks: incomplete obs operand sink");
    if (k->amp_hi && (!k->amp_lo || !k->amp_mean || !k->amp_inv_std || (k->ld_amp & 7) || k->ld_amp < EML_AMP_OBS))
        return fail(EMLOCO_EINVAL, "emloco_set_post_sinks: incomplete AMP operand sink");
    if (k->self_hi2 && (!k->self_hi || !k->self_lo2 || !k->task_hi2 || !k->task_lo2))
        return fail(EMLOCO_EINVAL, "emloco_set_post_sinks: the second operand set needs the first and all four of its pointers");
    s->sinks = *k;
    return EMLOCO_OK;
}

int emloco_post_step(emloco_sim* s, int32_t advance_progress, void* stream) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_post_step: null sim");
    CK(eml_launch_post_step(s, advance_progress ? 1 : 0, (cudaStream_t)stream), "post-step kernel");
    return EMLOCO_OK;
}

int emloco_step(emloco_sim* s, const float* d_actions, void* stream) {
    if (!s || !d_actions) return fail(EMLOCO_EINVAL, "emloco_step: null argument");
    CK(eml_launch_physics(s, d_actions, s->cfg.substeps * s->cfg.control_freq_inv, 1, (cudaStream_t)stream), "physics kernel");
    CK(eml_launch_post_step(s, 1, (cudaStream_t)stream), "post-step kernel");
    return EMLOCO_OK;
}

int emloco_physics_step(emloco_sim* s, const float* d_actions, void* stream) {
    if (!s || !d_actions) return fail(EMLOCO_EINVAL, "emloco_physics_step: null argument");
    CK(eml_launch_physics(s, d_actions, s->cfg.substeps * s->cfg.control_freq_inv, 1, (cudaStream_t)stream), "physics kernel");
    return EMLOCO_OK;
}

int emloco_step_host(emloco_sim* s, const float* h_actions, float* h_obs, float* h_rew, int64_t* h_reset, float* h_amp_obs) {
    if (!s || !h_actions) return fail(EMLOCO_EINVAL, "emloco_step_host: null argument");
    CK(cudaSetDevice(s->device), "cudaSetDevice");
    const size_t N = (size_t)s->N;
    float* p_act = s->h_pin;
    float* p_obs = p_act + N * EML_ND;
    float* p_rew = p_obs + N * EML_OBS;
    float* p_amp = p_rew + N;
    int64_t* p_reset = (int64_t*)(p_amp + N * EML_AMP_OBS);
    memcpy(p_act, h_actions, N * EML_ND * sizeof(float));
    cudaStream_t st = 0;
    float* d_act = s->actions;   // staging: the physics kernel reads P.actions and re-writes the same values
    CK(cudaMemcpyAsync(d_act, p_act, N * EML_ND * sizeof(float), cudaMemcpyHostToDevice, st), "H2D actions");
    int rc = emloco_step(s, d_act, st);
    if (rc) return rc;
    if (h_obs) CK(cudaMemcpyAsync(p_obs, s->obs, N * EML_OBS * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H obs");
    if (h_rew) CK(cudaMemcpyAsync(p_rew, s->rew, N * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H rew");
    if (h_reset) CK(cudaMemcpyAsync(p_reset, s->reset, N * sizeof(int64_t), cudaMemcpyDeviceToHost, st), "D2H reset");
    if (h_amp_obs) CK(cudaMemcpyAsync(p_amp, s->amp_obs, N * EML_AMP_OBS * sizeof(float), cudaMemcpyDeviceToHost, st), "D2H amp");
    CK(cudaStreamSynchronize(st), "step_host sync");
    if (h_obs) memcpy(h_obs, p_obs, N * EML_OBS * sizeof(float));
    if (h_rew) memcpy(h_rew, p_rew, N * sizeof(float));
    if (h_reset) memcpy(h_reset, p_reset, N * sizeof(int64_t));
    if (h_amp_obs) memcpy(h_amp_obs, p_amp, N * EML_AMP_OBS * sizeof(float));
    return EMLOCO_OK;
}

int emloco_locoval_forward(const float* d_traj, int32_t traj_stride, int32_t T, float* d_pose, const float* d_vel,
                           const float* d_weights, float* d_value, int64_t batch, int32_t flags, void* stream) {
    if (batch == 0) return EMLOCO_OK;
    if (!d_traj || !d_weights || !d_value) return fail(EMLOCO_EINVAL, "emloco_locoval_forward: null argument");
    if ((flags & 1) && !d_pose) return fail(EMLOCO_EINVAL, "emloco_locoval_forward: init_pose should be included");   // value_pose_net.py:114,136
    if ((flags & 2) && !d_vel) return fail(EMLOCO_EINVAL, "emloco_locoval_forward: init_vel should be included");
    if (traj_stride < 2 || (T != 13 && T != 5) || batch < 0) return fail(EMLOCO_EINVAL, "emloco_locoval_forward: bad shape");
    CK(eml_locoval_forward(d_traj, traj_stride, T, d_pose, d_vel, d_weights, d_value, batch, flags, (cudaStream_t)stream), "locoval forward");
    return EMLOCO_OK;
}

int emloco_locoval_backward_pose(const float* d_traj, int32_t traj_stride, int32_t T, const float* d_pose, const float* d_vel,
                                 const float* d_weights, const float* d_grad_value, float* d_grad_traj, const float* d_grad_pose_out,
                                 float* d_grad_pose_in, int64_t batch, int32_t flags, void* stream) {
    if (batch == 0) return EMLOCO_OK;
    if (!d_traj || !d_weights || !d_grad_value || !d_grad_traj) return fail(EMLOCO_EINVAL, "emloco_locoval_backward: null argument");
    if ((flags & 1) && !d_pose) return fail(EMLOCO_EINVAL, "emloco_locoval_backward: init_pose should be included");
    if ((flags & 2) && !d_vel) return fail(EMLOCO_EINVAL, "emloco_locoval_backward: init_vel should be included");
    if (!(flags & 1) && (d_grad_pose_out || d_grad_pose_in)) return fail(EMLOCO_EINVAL, "emloco_locoval_backward: pose gradients without use_pose");
    if (traj_stride < 2 || (T != 13 && T != 5) || batch < 0) return fail(EMLOCO_EINVAL, "emloco_locoval_backward: bad shape");
    CK(eml_locoval_backward(d_traj, traj_stride, T, d_pose, d_vel, d_weights, d_grad_value, d_grad_traj, d_grad_pose_out, d_grad_pose_in,
                            batch, flags, (cudaStream_t)stream), "locoval backward");
    return EMLOCO_OK;
}

int emloco_locoval_backward(const float* d_traj, int32_t traj_stride, int32_t T, const float* d_pose, const float* d_vel,
                            const float* d_weights, const float* d_grad_value, float* d_grad_traj, int64_t batch, int32_t flags,
                            void* stream) {
    return emloco_locoval_backward_pose(d_traj, traj_stride, T, d_pose, d_vel, d_weights, d_grad_value, d_grad_traj, nullptr, nullptr,
                                        batch, flags, stream);
}

int64_t emloco_locoval_train_workspace_bytes(int64_t N) { return N < 0 ? 0 : (int64_t)eml_locoval_train_workspace_bytes(N); }

int emloco_locoval_train_step(const float* d_traj, int32_t traj_stride, int32_t T, const float* d_pose, const float* d_vel,
                              float* d_gc, float* d_w, float* d_m, float* d_v, float* d_step, float* d_stats, void* d_ws, int64_t N,
                              float lr, float beta1, float beta2, float eps, float wd, float r_min, float r_max, int32_t flags,
                              void* stream) {
    if (N == 0) return EMLOCO_OK;
    if (!d_traj || !d_gc || !d_w || !d_m || !d_v || !d_step || !d_stats || !d_ws) return fail(EMLOCO_EINVAL, "emloco_locoval_train_step: null argument");
    if ((flags & 1) && !d_pose) return fail(EMLOCO_EINVAL, "emloco_locoval_train_step: init_pose should be included");
    if ((flags & 2) && !d_vel) return fail(EMLOCO_EINVAL, "emloco_locoval_train_step: init_vel should be included");
    if (traj_stride < 2 || (T != 13 && T != 5) || N < 0 || !(r_max > r_min)) return fail(EMLOCO_EINVAL, "emloco_locoval_train_step: bad shape or reward range");
    CK(eml_locoval_train_step(d_traj, traj_stride, T, d_pose, d_vel, d_gc, d_w, d_m, d_v, d_step, d_stats, d_ws, N, lr, beta1, beta2, eps,
                              wd, r_min, r_max, flags, (cudaStream_t)stream), "locoval train step");
    return EMLOCO_OK;
}

static int locoval_nweights(int T, int flags) {
    int in = 2 * T + ((flags & 1) ? 72 : 0) + ((flags & 2) ? 2 : 0);
    int h1 = in / 2 - 1, h2 = h1 / 2;
    return in * h1 + h1 + h1 * h2 + h2 + h2 + 1;
}

int emloco_locoval_forward_host(const float* h_traj, int32_t traj_stride, int32_t T, const float* h_pose, const float* h_vel,
                                const float* h_weights, float* h_value, int64_t B, int32_t flags, int32_t device) {
    if (!h_traj || !h_weights || !h_value) return fail(EMLOCO_EINVAL, "emloco_locoval_forward_host: null argument");
    if (B <= 0) return B == 0 ? EMLOCO_OK : fail(EMLOCO_EINVAL, "emloco_locoval_forward_host: negative batch");
    CK(cudaSetDevice(device), "cudaSetDevice");
    float *dt = nullptr, *dp = nullptr, *dv = nullptr, *dw = nullptr, *dval = nullptr;
    size_t nt = (size_t)B * T * traj_stride, nw = locoval_nweights(T, flags);
    CK(cudaMalloc((void**)&dt, nt * 4), "malloc traj"); CK(cudaMalloc((void**)&dw, nw * 4), "malloc w"); CK(cudaMalloc((void**)&dval, (size_t)B * 4), "malloc value");
    CK(cudaMemcpyAsync(dt, h_traj, nt * 4, cudaMemcpyHostToDevice, 0), "H2D traj");
    CK(cudaMemcpyAsync(dw, h_weights, nw * 4, cudaMemcpyHostToDevice, 0), "H2D w");
    if (flags & 1) { CK(cudaMalloc((void**)&dp, (size_t)B * 72 * 4), "malloc pose"); CK(cudaMemcpyAsync(dp, h_pose, (size_t)B * 72 * 4, cudaMemcpyHostToDevice, 0), "H2D pose"); }
    if (flags & 2) { CK(cudaMalloc((void**)&dv, (size_t)B * 2 * 4), "malloc vel"); CK(cudaMemcpyAsync(dv, h_vel, (size_t)B * 2 * 4, cudaMemcpyHostToDevice, 0), "H2D vel"); }
    int rc = emloco_locoval_forward(dt, traj_stride, T, dp, dv, dw, dval, B, flags & ~32, 0);
    if (!rc) { cudaError_t e = cudaMemcpy(h_value, dval, (size_t)B * 4, cudaMemcpyDeviceToHost); if (e != cudaSuccess) rc = fail(EMLOCO_ECUDA, "D2H value", e); }
    cudaFree(dt); cudaFree(dw); cudaFree(dval); if (dp) cudaFree(dp); if (dv) cudaFree(dv);
    return rc;
}

int emloco_plausibl_mlp_forward(const float* d_x, const float* d_weights, float* d_value, int64_t batch, void* stream) {
    if (!d_x || !d_weights || !d_value || batch < 0) return fail(EMLOCO_EINVAL, "emloco_plausibl_mlp_forward: bad argument");
    CK(eml_plausibl_forward(d_x, d_weights, d_value, batch, (cudaStream_t)stream), "plausibl mlp");
    return EMLOCO_OK;
}

int emloco_gae(const float* d_dones, const float* d_values, const float* d_rewards, const float* d_next_values, float* d_adv,
               float* d_ret, int32_t T, int64_t N, float gamma, float tau, void* stream) {
    if (!d_dones || !d_values || !d_rewards || !d_next_values || !d_adv || T < 0 || N < 0) return fail(EMLOCO_EINVAL, "emloco_gae: bad argument");
    CK(eml_gae(d_dones, d_values, d_rewards, d_next_values, d_adv, d_ret, T, N, gamma, tau, (cudaStream_t)stream), "gae kernel");
    return EMLOCO_OK;
}

int emloco_linear(const float* d_x, int64_t ldx, const float* d_w, const float* d_b, float* d_y, int64_t ldy, int64_t M, int32_t N,
                  int32_t K, const float* d_mean, const float* d_var, float eps, int32_t relu, int32_t use_tc, void* stream) {
    if (!d_x || !d_w || !d_y || M < 0 || N <= 0 || K <= 0 || ldx < K || ldy < N) return fail(EMLOCO_EINVAL, "emloco_linear: bad argument");
    if ((d_mean == nullptr) != (d_var == nullptr)) return fail(EMLOCO_EINVAL, "emloco_linear: mean and var must come together");
    if (use_tc) CK(eml_linear_tc(d_x, ldx, d_w, d_b, d_y, ldy, M, N, K, d_mean, d_var, eps, relu, (cudaStream_t)stream), "linear (tensor core)");
    else CK(eml_linear_fma(d_x, ldx, d_w, d_b, d_y, ldy, M, N, K, d_mean, d_var, eps, relu, (cudaStream_t)stream), "linear (fma)");
    return EMLOCO_OK;
}

int emloco_reset_done(emloco_sim* s, const float* d_init_root, const float* d_init_dof, void* stream) {
    if (!s || !d_init_root || !d_init_dof) return fail(EMLOCO_EINVAL, "emloco_reset_done: null argument");
    CK(eml_reset_done(s, d_init_root, d_init_dof, (cudaStream_t)stream), "reset-done kernels");
    return EMLOCO_OK;
}

static int check_traj_cfg(const emloco_traj_cfg* c) {
    if (c->uniform && c->ld_uniform < EMLOCO_TRAJ_RAND_COLS) return fail(EMLOCO_EINVAL, "trajectory reset: ld_uniform < 405");
    if ((c->flags & EMLOCO_TRAJ_REAL_PATH) && (!c->pool || c->pool_count <= 0)) return fail(EMLOCO_EINVAL, "trajectory reset: real_path needs a trajectory pool");
    if (c->num_waypoints < 0 || c->num_waypoints > EML_TRAJ_SAMPLES) return fail(EMLOCO_EINVAL, "trajectory reset: num_waypoints must be 0..15");
    if (!(c->speed_max >= c->speed_min) || !(c->speed_min > 0.0f)) return fail(EMLOCO_EINVAL, "trajectory reset: need 0 < speed_min <= speed_max");
    return EMLOCO_OK;
}

int emloco_traj_reset(emloco_sim* s, const emloco_traj_cfg* c, void* stream) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_traj_reset: null sim");
    if (!c) {
        if (s->traj_on != 2) return fail(EMLOCO_EINVAL, "emloco_traj_reset: no deferred stage stored (emloco_set_traj_reset with EMLOCO_TRAJ_DEFERRED)");
        CK(eml_traj_reset(s, s->traj, 1, (cudaStream_t)stream), "trajectory reset kernel");
        return EMLOCO_OK;
    }
    if (int e = check_traj_cfg(c)) return e;
    CK(eml_traj_reset(s, *c, 0, (cudaStream_t)stream), "trajectory reset kernel");
    return EMLOCO_OK;
}

int emloco_set_traj_reset(emloco_sim* s, const emloco_traj_cfg* c) {
    if (!s) return fail(EMLOCO_EINVAL, "emloco_set_traj_reset: null sim");
    if (!c) { s->traj_on = 0; return EMLOCO_OK; }
    if (int e = check_traj_cfg(c)) return e;
    s->traj = *c; s->traj_on = (c->flags & EMLOCO_TRAJ_DEFERRED) ? 2 : 1;
    return EMLOCO_OK;
}

int emloco_sample_actions(const float* d_mu, int64_t ldmu, const float* d_logstd, const float* d_noise, float* d_actions,
                          float* d_neglogp, int64_t N, int32_t A, void* stream) {
    if (!d_mu || !d_logstd || !d_noise || !d_actions || N < 0 || A <= 0 || ldmu < A) return fail(EMLOCO_EINVAL, "emloco_sample_actions: bad argument");
    CK(eml_sample_actions(d_mu, ldmu, d_logstd, d_noise, d_actions, d_neglogp, N, A, (cudaStream_t)stream), "sample actions");
    return EMLOCO_OK;
}

int emloco_sample_actions_parts(const float* d_mu_parts, int64_t ldmu, int32_t parts, int64_t part_stride, float* d_mu_out, int64_t ldout,
                                const float* d_logstd, const float* d_noise, float* d_actions, float* d_neglogp, int64_t N, int32_t A,
                                void* stream) {
    if (!d_mu_parts || !d_logstd || !d_noise || !d_actions || N < 0 || A <= 0 || ldmu < A || parts < 1 || (d_mu_out && ldout < A))
        return fail(EMLOCO_EINVAL, "emloco_sample_actions_parts: bad argument");
    CK(eml_sample_actions_parts(d_mu_parts, ldmu, parts, part_stride, d_mu_out, ldout, d_logstd, d_noise, d_actions, d_neglogp, N, A,
                                (cudaStream_t)stream), "sample actions");
    return EMLOCO_OK;
}

int emloco_disc_reward(const float* d_logit, const float* d_task_rew, float* d_disc, float* d_combined, int64_t M, float scale,
                       float w_task, float w_disc, void* stream) {
    if ((!d_logit && !d_disc) || M < 0 || (d_combined && !d_task_rew)) return fail(EMLOCO_EINVAL, "emloco_disc_reward: bad argument");
    CK(eml_disc_reward(d_logit, d_task_rew, d_disc, d_combined, M, scale, w_task, w_disc, (cudaStream_t)stream), "disc reward");
    return EMLOCO_OK;
}

int emloco_rollout_record(const emloco_rollout_cfg* c, const float* d_rew, const int64_t* d_reset, const int64_t* d_terminate,
                          const float* d_value_raw, const float* d_next_value_raw, const float* d_disc_logit,
                          const uint8_t* d_inverted, float* d_mb_values, float* d_mb_rewards,
                          float* d_mb_dones, float* d_mb_next_values, float* d_mb_amp_rewards, float* d_state, int64_t N,
                          void* stream) {
    if (!c || !d_rew || !d_reset || !d_terminate || !d_next_value_raw || !d_disc_logit || !d_mb_rewards || !d_mb_dones ||
        !d_mb_next_values || !d_state || N < 0 || (d_value_raw && !d_mb_values))
        return fail(EMLOCO_EINVAL, "emloco_rollout_record: bad argument");
    RecordParams P;
    P.rew = d_rew; P.reset = d_reset; P.terminate = d_terminate; P.value_raw = d_value_raw; P.mb_values = d_mb_values;
    P.next_value_raw = d_next_value_raw; P.disc_logit = d_disc_logit;
    P.inverted = d_inverted; P.mb_rewards = d_mb_rewards; P.mb_dones = d_mb_dones; P.mb_next_values = d_mb_next_values;
    P.mb_amp_rewards = d_mb_amp_rewards;
    P.current_rewards = d_state; P.current_lengths = d_state + N; P.current_combined = d_state + 2 * N;
    P.discount_coefs = d_state + 3 * N; P.game_combined = d_state + 4 * N; P.terminated_flags = d_state + 5 * N;
    P.N = N; P.inv_penalty = c->inversion_penalty_scale; P.reward_scale = c->reward_scale; P.v_mean = c->value_mean;
    P.v_std = c->value_std; P.v_stats = c->d_value_stats; P.disc_scale = c->disc_reward_scale; P.gamma = c->gamma; P.step_to_pred = (float)c->step_to_pred;
    P.unnorm_value = c->unnorm_value;
    P.c_value_raw = nullptr; P.c_idx = nullptr; P.c_count = nullptr; P.prev_dones = nullptr; P.prev_next_values = nullptr;
    CK(eml_rollout_record(P, (cudaStream_t)stream), "rollout record");
    return EMLOCO_OK;
}

int emloco_rollout_record_deferred(const emloco_rollout_cfg* c, const float* d_rew, const int64_t* d_reset, const int64_t* d_terminate,
                                   const float* d_value_raw, const float* d_disc_logit, const uint8_t* d_inverted, float* d_mb_values,
                                   float* d_mb_rewards, float* d_mb_dones, float* d_mb_next_values, float* d_mb_amp_rewards,
                                   float* d_state, int64_t N, const float* d_c_value_raw, const int32_t* d_c_idx,
                                   const int32_t* d_c_count, const float* d_prev_dones, float* d_prev_next_values, void* stream) {
    if (!c || !d_rew || !d_reset || !d_terminate || !d_value_raw || !d_disc_logit || !d_mb_values || !d_mb_rewards || !d_mb_dones ||
        !d_mb_next_values || !d_state || N < 0 || !d_c_value_raw || !d_c_idx || !d_c_count || ((d_prev_dones == nullptr) != (d_prev_next_values == nullptr)))
        return fail(EMLOCO_EINVAL, "emloco_rollout_record_deferred: bad argument");
    RecordParams P;
    P.rew = d_rew; P.reset = d_reset; P.terminate = d_terminate; P.value_raw = d_value_raw; P.mb_values = d_mb_values;
    P.next_value_raw = nullptr; P.disc_logit = d_disc_logit; P.inverted = d_inverted;
    P.mb_rewards = d_mb_rewards; P.mb_dones = d_mb_dones; P.mb_next_values = d_mb_next_values; P.mb_amp_rewards = d_mb_amp_rewards;
    P.current_rewards = d_state; P.current_lengths = d_state + N; P.current_combined = d_state + 2 * N;
    P.discount_coefs = d_state + 3 * N; P.game_combined = d_state + 4 * N; P.terminated_flags = d_state + 5 * N;
    P.N = N; P.inv_penalty = c->inversion_penalty_scale; P.reward_scale = c->reward_scale; P.v_mean = c->value_mean;
    P.v_std = c->value_std; P.v_stats = c->d_value_stats; P.disc_scale = c->disc_reward_scale; P.gamma = c->gamma; P.step_to_pred = (float)c->step_to_pred;
    P.unnorm_value = c->unnorm_value;
    P.c_value_raw = d_c_value_raw; P.c_idx = d_c_idx; P.c_count = d_c_count; P.prev_dones = d_prev_dones; P.prev_next_values = d_prev_next_values;
    CK(eml_rollout_record(P, (cudaStream_t)stream), "rollout record (deferred next values)");
    return EMLOCO_OK;
}

int emloco_fill_next_values(const emloco_rollout_cfg* c, const float* d_value_raw, const float* d_prev_dones, float* d_prev_next_values,
                            int64_t N, void* stream) {
    if (!c || !d_value_raw || !d_prev_dones || !d_prev_next_values || N < 0) return fail(EMLOCO_EINVAL, "emloco_fill_next_values: bad argument");
    CK(eml_fill_next_values(d_value_raw, d_prev_dones, d_prev_next_values, N, c->value_mean, c->value_std, c->d_value_stats, c->unnorm_value,
                            (cudaStream_t)stream), "fill next values");
    return EMLOCO_OK;
}

int emloco_timeout_gather(const int64_t* d_reset, const int64_t* d_terminate, int64_t N, const uint16_t* self_hi, const uint16_t* self_lo,
                          int64_t ld_self, const uint16_t* task_hi, const uint16_t* task_lo, int64_t ld_task, uint16_t* c_self_hi,
                          uint16_t* c_self_lo, int64_t ld_cself, uint16_t* c_task_hi, uint16_t* c_task_lo, int64_t ld_ctask, int32_t* d_idx,
                          int32_t* d_count, void* stream) {
    if (!d_reset || !d_terminate || !self_hi || !self_lo || !task_hi || !task_lo || !c_self_hi || !c_self_lo || !c_task_hi || !c_task_lo ||
        !d_idx || !d_count || N < 0 || (ld_self & 7) || (ld_task & 7) || (ld_cself & 7) || (ld_ctask & 7) || ld_ctask < ld_task)
        return fail(EMLOCO_EINVAL, "emloco_timeout_gather: bad argument");
    CK(eml_timeout_gather(d_reset, d_terminate, N, self_hi, self_lo, ld_self, task_hi, task_lo, ld_task, c_self_hi, c_self_lo, ld_cself,
                          c_task_hi, c_task_lo, ld_ctask, d_idx, d_count, (cudaStream_t)stream), "timeout gather");
    return EMLOCO_OK;
}

int emloco_split_bf16(const float* d_x, int64_t ldx, int64_t M, int32_t K, const float* d_mean, const float* d_var, float eps,
                      uint16_t* d_hi, uint16_t* d_lo, int64_t ld16, void* stream) {
    if (!d_x || !d_hi || !d_lo || M < 0 || K <= 0 || ldx < K || ld16 < K || (ld16 & 7) || ((uintptr_t)d_hi & 15) || ((uintptr_t)d_lo & 15))
        return fail(EMLOCO_EINVAL, "emloco_split_bf16: bad argument (pitch must be a multiple of 8, pointers 16-byte aligned)");
    if ((d_mean == nullptr) != (d_var == nullptr)) return fail(EMLOCO_EINVAL, "emloco_split_bf16: mean and var must come together");
    CK(eml_split_bf16(d_x, ldx, M, K, d_mean, d_var, eps, d_hi, d_lo, ld16, (cudaStream_t)stream), "split bf16");
    return EMLOCO_OK;
}

static int linear_bf16x3_impl(const int32_t* d_rows, const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* w_hi, const uint16_t* w_lo,
                         int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K, int32_t relu, float* d_y32, int64_t ldy,
                         uint16_t* y_hi, uint16_t* y_lo, int64_t ldy16, void* stream, const float* head_w = nullptr,
                         float* head_part = nullptr) {
    if (!a_hi || !a_lo || !w_hi || !w_lo || M < 0 || N <= 0 || K <= 0 || lda < K || ldw < K || (lda & 7) || (ldw & 7))
        return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: bad operand");
    if (((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)w_hi | (uintptr_t)w_lo) & 15)
        return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: operand pointers must be 16-byte aligned");
    if (!d_y32 && !y_hi && !head_w) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: no output");
    if (d_y32 && ldy < N) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: ldy < N");
    if ((y_hi == nullptr) != (y_lo == nullptr)) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: y_hi and y_lo must come together");
    if (y_hi && ((N & 31) || ldy16 < N || (ldy16 & 7) || (((uintptr_t)y_hi | (uintptr_t)y_lo) & 15)))
        return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: split output needs N % 32 == 0, pitch % 8 == 0, 16-byte aligned pointers");
    int tile_n = (relu >> 8) & 0xfff;           // bits 8..19 of `relu`: 0 = automatic tile choice, 128 / 256 = forced (tests, tuning)
    if (tile_n != 0 && (tile_n & 0x7ff) != 128 && (tile_n & 0x7ff) != 256)
        return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: tile must be 0, 128 or 256 (+0x800 for the CTA-pair kernels)");
    const int splits = (relu >> 20) & 0xf;      // bits 20..23: split-K count (0 / 1 = off)
    if (splits > 1) {
        if ((relu & 1) || y_hi || head_w || !d_y32 || (tile_n & 0x800) || d_rows)
            return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3: split-K needs a plain fp32 output (no ReLU, split output, fused head, row count or CTA-pair tile)");
        tile_n |= splits << 12;
    }
    if (head_w && (tile_n & 0x800)) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3_head: not available with the CTA-pair kernels");
    CK(eml_linear_bf16x3(a_hi, a_lo, lda, w_hi, w_lo, ldw, d_bias, M, N, K, relu & 1, d_y32, ldy, y_hi, y_lo, ldy16, tile_n,
                         d_rows, head_w, head_part, (N + 63) / 64, (cudaStream_t)stream), "linear bf16x3 (tcgen05)");
    return EMLOCO_OK;
}

int emloco_linear_bf16x3_head(const int32_t* d_rows, const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* w_hi,
                              const uint16_t* w_lo, int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K, int32_t relu,
                              float* d_y32, int64_t ldy, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy16, const float* d_head_w,
                              const float* d_head_bias, float* d_head_part, float* d_head_out, void* stream) {
    if (!d_head_w || !d_head_part || !d_head_out) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3_head: null head argument");
    if (int e = linear_bf16x3_impl(d_rows, a_hi, a_lo, lda, w_hi, w_lo, ldw, d_bias, M, N, K, relu, d_y32, ldy, y_hi, y_lo, ldy16, stream,
                                   d_head_w, d_head_part)) return e;
    CK(eml_head_reduce(d_head_part, (N + 63) / 64, d_head_bias, d_head_out, M, (cudaStream_t)stream), "head reduce");
    return EMLOCO_OK;
}

int64_t emloco_linear_chain_workspace_ints(const emloco_chain_layer* layers, int32_t n_layers) {
    if (!layers || n_layers <= 0) return 0;
    return eml_linear_chain_workspace_ints(layers, n_layers);
}

int emloco_linear_chain_trace(int64_t* d_trace) {
    eml_linear_chain_trace((long long*)d_trace);
    return EMLOCO_OK;
}

int emloco_linear_chain(const emloco_chain_layer* layers, int32_t n_layers, const int32_t* order, int32_t n_segments,
                        int32_t* d_workspace, int64_t workspace_ints, void* stream) {
    const char* why = nullptr;
    cudaError_t e = eml_linear_chain(layers, n_layers, order, n_segments, d_workspace, workspace_ints, (cudaStream_t)stream, &why);
    if (e == cudaErrorInvalidValue && why) return fail(EMLOCO_EINVAL, why);
    CK(e, "linear chain (tcgen05)");
    return EMLOCO_OK;
}

int emloco_linear_bf16x3(const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* w_hi, const uint16_t* w_lo,
                         int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K, int32_t relu, float* d_y32, int64_t ldy,
                         uint16_t* y_hi, uint16_t* y_lo, int64_t ldy16, void* stream) {
    return linear_bf16x3_impl(nullptr, a_hi, a_lo, lda, w_hi, w_lo, ldw, d_bias, M, N, K, relu, d_y32, ldy, y_hi, y_lo, ldy16, stream);
}

int emloco_linear_bf16x3_rows(const int32_t* d_rows, const uint16_t* a_hi, const uint16_t* a_lo, int64_t lda, const uint16_t* w_hi,
                              const uint16_t* w_lo, int64_t ldw, const float* d_bias, int64_t M, int32_t N, int32_t K, int32_t relu,
                              float* d_y32, int64_t ldy, uint16_t* y_hi, uint16_t* y_lo, int64_t ldy16, void* stream) {
    if (!d_rows) return fail(EMLOCO_EINVAL, "emloco_linear_bf16x3_rows: null row count");
    return linear_bf16x3_impl(d_rows, a_hi, a_lo, lda, w_hi, w_lo, ldw, d_bias, M, N, K, relu, d_y32, ldy, y_hi, y_lo, ldy16, stream);
}

int emloco_normalize(const float* d_x, int64_t ldx, float* d_y, int64_t ldy, int64_t M, int32_t K, const float* d_mean,
                     const float* d_var, float eps, void* stream) {
    if (!d_x || !d_y || !d_mean || !d_var || M < 0 || K <= 0 || ldx < K || ldy < K) return fail(EMLOCO_EINVAL, "emloco_normalize: bad argument");
    CK(eml_normalize(d_x, ldx, d_y, ldy, M, K, d_mean, d_var, eps, (cudaStream_t)stream), "normalize");
    return EMLOCO_OK;
}

int emloco_player_record(const float* d_rew, const float* d_rew_raw, const int64_t* d_reset, const float* d_disc_logit,
                         const float* d_locoval_scores, const uint8_t* d_inverted, float* d_state, int64_t N, float* d_results,
                         int32_t* d_count, int32_t capacity, int32_t plot_val_reward, float inversion_penalty_scale,
                         float disc_reward_scale, float gamma, int32_t step_to_pred, float min_reward, float max_reward, void* stream) {
    if (!d_rew || !d_rew_raw || !d_reset || !d_disc_logit || !d_locoval_scores || !d_state || !d_results || !d_count || N < 0 || capacity < 0 ||
        max_reward == min_reward)
        return fail(EMLOCO_EINVAL, "emloco_player_record: bad argument");
    CK(eml_player_record(d_rew, d_rew_raw, d_reset, d_disc_logit, d_locoval_scores, d_inverted, d_state, N, d_results, d_count, capacity,
                         plot_val_reward, inversion_penalty_scale, disc_reward_scale, gamma, step_to_pred, min_reward, max_reward,
                         (cudaStream_t)stream), "player record");
    return EMLOCO_OK;
}

// ---- mocap reset state / AMP demo observations (csrc/motion.cu) ----
static bool motion_lib_ok(const emloco_motion_lib* L) {
    return L && L->d_gts && L->d_grs && L->d_lrs && L->d_gvs && L->d_gavs && L->d_dvs && L->d_length && L->d_dt && L->d_bodies &&
           L->d_num_frames && L->d_start && L->num_motions > 0;
}
static EmlMotionLibDev motion_lib_dev(const emloco_motion_lib* L) {
    EmlMotionLibDev D;
    D.gts = L->d_gts; D.grs = L->d_grs; D.lrs = L->d_lrs; D.gvs = L->d_gvs; D.gavs = L->d_gavs; D.dvs = L->d_dvs;
    D.length = L->d_length; D.dt = L->d_dt; D.bodies = L->d_bodies; D.num_frames = L->d_num_frames; D.start = L->d_start;
    D.num_motions = L->num_motions;
    return D;
}
int emloco_motion_state(const emloco_motion_lib* lib, const int32_t* d_motion_ids, const float* d_motion_times, int64_t n,
                        float* d_root_state, float* d_dof_state, float* d_key_pos, float* d_rb_state, void* stream) {
    if (!motion_lib_ok(lib) || !d_motion_ids || !d_motion_times || n < 0) return fail(EMLOCO_EINVAL, "emloco_motion_state: bad argument");
    CK(eml_motion_state(motion_lib_dev(lib), d_motion_ids, d_motion_times, n, d_root_state, d_dof_state, d_key_pos, d_rb_state,
                        (cudaStream_t)stream), "motion state");
    return EMLOCO_OK;
}
int emloco_amp_obs_demo(const emloco_motion_lib* lib, const int32_t* d_motion_ids, const float* d_motion_times0, int64_t n,
                        int32_t num_steps, float dt, float* d_amp_obs, void* stream) {
    if (!motion_lib_ok(lib) || !d_motion_ids || !d_motion_times0 || !d_amp_obs || n < 0 || num_steps <= 0 || !(dt > 0))
        return fail(EMLOCO_EINVAL, "emloco_amp_obs_demo: bad argument");
    CK(eml_amp_obs_demo(motion_lib_dev(lib), d_motion_ids, d_motion_times0, n, num_steps, dt, d_amp_obs, (cudaStream_t)stream), "amp obs demo");
    return EMLOCO_OK;
}

// ---- PPO / AMP update step (csrc/update.cu) ----
int emloco_xform(const float* d_x, int64_t ldx, const float* d_rowvec, const float* d_rowscale, int64_t lds, const float* d_mean,
                 const float* d_var, float eps, const float* d_gate, int64_t ldg, const float* d_drop_u, float drop_rate, float scale,
                 float* d_y32, int64_t ldy, float* d_yT32, int64_t ldyT, uint16_t* d_hi, uint16_t* d_lo, int64_t ld16, uint16_t* d_hiT,
                 uint16_t* d_loT, int64_t ldT, float* d_colsum, float* d_sumsq, int64_t M, int32_t K, void* stream) {
    if ((!d_x && !d_rowvec) || M < 0 || K < 0 || ((d_mean == nullptr) != (d_var == nullptr)) || (d_mean && !d_x) ||
        ((d_hi == nullptr) != (d_lo == nullptr)) || ((d_hiT == nullptr) != (d_loT == nullptr)))
        return fail(EMLOCO_EINVAL, "emloco_xform: bad argument");
    if ((d_hi && (ld16 < K)) || (d_hiT && (ldT < M)) || (d_y32 && ldy < K) || (d_yT32 && ldyT < M))
        return fail(EMLOCO_EINVAL, "emloco_xform: output pitch smaller than the row length");
    if (d_drop_u && K % EML_AMP_STEP) return fail(EMLOCO_EINVAL, "emloco_xform: the joint-dropout gate applies to AMP observation rows (K a multiple of 206)");
    CK(eml_xform(d_x, ldx, d_rowvec, d_rowscale, lds, d_mean, d_var, eps, d_gate, ldg, d_drop_u, drop_rate, scale, d_y32, ldy, d_yT32, ldyT, d_hi, d_lo, ld16,
                 d_hiT, d_loT, ldT, d_colsum, d_sumsq, M, K, (cudaStream_t)stream), "xform");
    return EMLOCO_OK;
}

int emloco_ppo_heads(const float* d_mu, int64_t ldmu, const float* d_logstd, const float* d_actions, const float* d_old_neglogp,
                     const float* d_adv, const float* d_value, const float* d_task_value, const float* d_returns,
                     const float* d_old_mu, const float* d_old_sigma, float* d_dmu, int64_t lddmu, float* d_dvalue, float* d_dtask,
                     float* d_stats, int64_t B, int32_t A, float e_clip, float actor_coef, float critic_coef, float tv_coef,
                     float bounds_coef, void* stream) {
    if (!d_mu || !d_logstd || !d_actions || !d_old_neglogp || !d_adv || !d_value || !d_task_value || !d_returns || !d_dmu || !d_dvalue ||
        !d_dtask || !d_stats || B < 0 || A <= 0 || ((d_old_mu == nullptr) != (d_old_sigma == nullptr)))
        return fail(EMLOCO_EINVAL, "emloco_ppo_heads: bad argument");
    CK(eml_ppo_heads(d_mu, ldmu, d_logstd, d_actions, d_old_neglogp, d_adv, d_value, d_task_value, d_returns, d_old_mu, d_old_sigma, d_dmu,
                     lddmu, d_dvalue, d_dtask, d_stats, B, A, e_clip, actor_coef, critic_coef, tv_coef, bounds_coef, (cudaStream_t)stream),
       "ppo heads");
    return EMLOCO_OK;
}

int emloco_disc_heads(const float* d_logit, float* d_dlogit, float* d_stats, int64_t n_agent, int64_t n_demo, float coef, void* stream) {
    if (!d_logit || !d_dlogit || !d_stats || n_agent < 0 || n_demo < 0) return fail(EMLOCO_EINVAL, "emloco_disc_heads: bad argument");
    CK(eml_disc_heads(d_logit, d_dlogit, d_stats, n_agent, n_demo, coef, (cudaStream_t)stream), "disc heads");
    return EMLOCO_OK;
}

int emloco_amp_dropout_mask(const float* d_u, float* d_mask, int64_t rows, float rate, void* stream) {
    if (!d_u || !d_mask || rows < 0) return fail(EMLOCO_EINVAL, "emloco_amp_dropout_mask: bad argument");
    CK(eml_amp_dropout_mask(d_u, d_mask, rows, rate, (cudaStream_t)stream), "amp dropout mask");
    return EMLOCO_OK;
}

int emloco_rms_update(const float* d_x, int64_t ldx, int64_t M, int32_t K, double* d_scratch, double* d_running_mean,
                      double* d_running_var, double* d_count, float* d_mean32, float* d_var32, float* d_inv_std32, float eps,
                      void* stream) {
    if (!d_x || !d_scratch || !d_running_mean || !d_running_var || !d_count || M < 0 || K < 0 ||
        (d_mean32 && (!d_var32 || !d_inv_std32)))
        return fail(EMLOCO_EINVAL, "emloco_rms_update: bad argument");
    CK(eml_rms_update(d_x, ldx, M, K, d_scratch, d_running_mean, d_running_var, d_count, d_mean32, d_var32, d_inv_std32, eps,
                      (cudaStream_t)stream), "rms update");
    return EMLOCO_OK;
}

int emloco_adam_begin(float* d_state, void* stream) {
    if (!d_state) return fail(EMLOCO_EINVAL, "emloco_adam_begin: null argument");
    CK(eml_adam_begin(d_state, (cudaStream_t)stream), "adam begin");
    return EMLOCO_OK;
}

int emloco_grad_sumsq(const float* d_grad, int64_t n, float* d_state, float* d_partials, void* stream) {
    if (!d_grad || !d_state || !d_partials || n < 0) return fail(EMLOCO_EINVAL, "emloco_grad_sumsq: bad argument");
    CK(eml_grad_sumsq(d_grad, n, d_state, d_partials, (cudaStream_t)stream), "grad sumsq");
    return EMLOCO_OK;
}

int emloco_adam_clip(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float* d_state, float lr, float beta1,
                     float beta2, float eps, float max_norm, float grad_scale, void* stream) {
    if (!d_param || !d_grad || !d_m || !d_v || !d_state || n < 0) return fail(EMLOCO_EINVAL, "emloco_adam_clip: bad argument");
    CK(eml_adam_clip(d_param, d_grad, d_m, d_v, n, d_state, lr, beta1, beta2, eps, max_norm, grad_scale, (cudaStream_t)stream), "adam");
    return EMLOCO_OK;
}

int emloco_dp_reduce_shard(const float* mc_grad, float* d_shard_grad, int64_t lo, int64_t count, float* d_partials, float* mc_exchange,
                           int32_t rank, void* stream) {
    if (!mc_grad || !d_shard_grad || !d_partials || !mc_exchange || lo < 0 || count < 0 || (lo & 3) || (count & 3) || rank < 0)
        return fail(EMLOCO_EINVAL, "emloco_dp_reduce_shard: bad argument (slices are multiples of 4 floats)");
    CK(eml_dp_reduce_shard(mc_grad, d_shard_grad, lo, count, d_partials, mc_exchange, rank, (cudaStream_t)stream), "dp reduce shard");
    return EMLOCO_OK;
}
int emloco_dp_adam_shard(float* mc_param, const float* d_param_local, const float* d_shard_grad, float* d_m, float* d_v, int64_t lo, int64_t count,
                         const float* d_exchange_local, int32_t world, float* d_state, float lr, float beta1, float beta2, float eps,
                         float max_norm, float grad_scale, void* stream) {
    if (!mc_param || !d_param_local || !d_shard_grad || !d_m || !d_v || !d_exchange_local || !d_state || lo < 0 || count < 0 || (lo & 3) ||
        (count & 3) || world < 1)
        return fail(EMLOCO_EINVAL, "emloco_dp_adam_shard: bad argument");
    CK(eml_dp_adam_shard(mc_param, d_param_local, d_shard_grad, d_m, d_v, lo, count, d_exchange_local, world, d_state, lr, beta1, beta2, eps,
                         max_norm, grad_scale, (cudaStream_t)stream), "dp adam shard");
    return EMLOCO_OK;
}

int emloco_sum_parts(const float* d_parts, int32_t num_parts, int64_t part_stride, float* d_out, int64_t n, int32_t accumulate, void* stream) {
    if (!d_parts || !d_out || num_parts <= 0 || n < 0 || part_stride < n) return fail(EMLOCO_EINVAL, "emloco_sum_parts: bad argument");
    CK(eml_sum_parts(d_parts, num_parts, part_stride, d_out, n, accumulate, (cudaStream_t)stream), "sum parts");
    return EMLOCO_OK;
}

int emloco_axpy(float* d_y, const float* d_x, float a, int64_t n, void* stream) {
    if (!d_y || !d_x || n < 0) return fail(EMLOCO_EINVAL, "emloco_axpy: bad argument");
    CK(eml_axpy(d_y, d_x, a, n, (cudaStream_t)stream), "axpy");
    return EMLOCO_OK;
}

int emloco_sync(emloco_sim* s) {
    if (s) CK(cudaSetDevice(s->device), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    return EMLOCO_OK;
}

}  // extern "C"
