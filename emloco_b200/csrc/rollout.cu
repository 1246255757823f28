// Small per-env kernels of the rollout loop (reference pacer/pacer/learning/amp_continuous_value.py:34-178
// `play_steps`), each replacing a run of eager torch launches:
//   sample_actions   a = mu + exp(logstd)*eps, neglogp            rl_games 1.1.4 ModelA2CContinuousLogStd (not in tree,
//                                                                pinned by pacer/requirements.txt); formula restated SURVEY 8c(5)
//   disc_reward      -log(max(1-sigmoid(logit),1e-4))*scale        learning/amp_continuous.py:675-692, combine :659-664
//   rollout_record   everything play_steps does per env after the nets of one step (:63-118): inversion penalty,
//                    value un-normalisation + terminated mask, AMP reward, LocoVal-target bookkeeping
#include "sim.h"

// ---- a = mu + sigma * eps ; neglogp = 0.5*sum(((a-mu)/sigma)^2) + 0.5*log(2pi)*A + sum(logstd) ----
// one warp per env row, lanes stride over the A action dims
// parts > 1: mu arrives as `parts` partial sums (split-K output of the mu layer, part p at mu + p * part_stride); they are added
// in order and the sum is written to mu_out
__global__ void __launch_bounds__(128) sample_actions_kernel(const float* __restrict__ mu, long long ldmu, int parts, long long part_stride,
                                                             float* __restrict__ mu_out, long long ldout,
                                                             const float* __restrict__ logstd, const float* __restrict__ noise,
                                                             float* __restrict__ actions, float* __restrict__ neglogp,
                                                             long long N, int A) {
    const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    float acc = 0.f, ls = 0.f;
    for (int j = lane; j < A; j += 32) {
        float l = logstd[j];
        float s = expf(l);
        float m = mu[row * ldmu + j];
        for (int p = 1; p < parts; ++p) m += mu[p * part_stride + row * ldmu + j];
        if (mu_out) mu_out[row * ldout + j] = m;
        float a = m + s * noise[row * A + j];
        actions[row * A + j] = a;
        float z = (a - m) / s;
        acc += z * z;
        ls += l;
    }
    acc = warp_sum(acc); ls = warp_sum(ls);
    if (lane == 0 && neglogp) neglogp[row] = 0.5f * acc + 0.91893853320467274178f * (float)A + ls;
}

cudaError_t eml_sample_actions(const float* mu, long long ldmu, const float* logstd, const float* noise, float* actions,
                               float* neglogp, long long N, int A, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    sample_actions_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(mu, ldmu, 1, 0, nullptr, 0, logstd, noise, actions, neglogp, N, A);
    return cudaGetLastError();
}

cudaError_t eml_sample_actions_parts(const float* mu_parts, long long ldmu, int parts, long long part_stride, float* mu_out, long long ldout,
                                     const float* logstd, const float* noise, float* actions, float* neglogp, long long N, int A,
                                     cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    sample_actions_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(mu_parts, ldmu, parts, part_stride, mu_out, ldout, logstd, noise, actions,
                                                                  neglogp, N, A);
    return cudaGetLastError();
}

__device__ __forceinline__ float disc_r(float logit, float scale) {
    float prob = 1.0f / (1.0f + expf(-logit));
    return -logf(fmaxf(1.0f - prob, 0.0001f)) * scale;
}

__global__ void disc_reward_kernel(const float* __restrict__ logit, const float* __restrict__ task_rew,
                                   float* __restrict__ disc, float* __restrict__ combined, long long M, float scale,
                                   float w_task, float w_disc) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float r = logit ? disc_r(logit[i], scale) : disc[i];     // logit == NULL: disc already holds the AMP rewards
    if (disc && logit) disc[i] = r;
    if (combined) combined[i] = w_task * task_rew[i] + w_disc * r;
}

cudaError_t eml_disc_reward(const float* logit, const float* task_rew, float* disc, float* combined, long long M, float scale,
                            float w_task, float w_disc, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    disc_reward_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(logit, task_rew, disc, combined, M, scale, w_task, w_disc);
    return cudaGetLastError();
}


__global__ void rollout_record_kernel(RecordParams P) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    float r = P.rew[i];
    if (P.v_stats) { P.v_mean = __ldg(P.v_stats); P.v_std = __ldg(P.v_stats + 1); }
    if (P.inverted && P.inverted[i]) r *= -P.inv_penalty;                 // :63-64
    float shaped = r * P.reward_scale;                                    // rewards_shaper (scale_value 1)
    float done = P.reset[i] != 0 ? 1.0f : 0.0f;
    float term = (float)P.terminate[i];
    float v_now = 0.f;
    if (P.value_raw) {                                                    // res_dict['values'] of get_action_values
        v_now = P.value_raw[i];
        if (P.unnorm_value) v_now = P.v_std * fminf(fmaxf(v_now, -5.0f), 5.0f) + P.v_mean;   // running_mean_std.py:77-79
        P.mb_values[i] = v_now;
    }
    const bool deferred = P.next_value_raw == nullptr;
    float nv = 0.f;
    if (!deferred) {
        nv = P.next_value_raw[i];
        if (P.unnorm_value) nv = P.v_std * fminf(fmaxf(nv, -5.0f), 5.0f) + P.v_mean;
        nv *= (1.0f - term);                                              // :86-90
    } else {
        // value reuse: critic(next obs) of an env that is NOT reset equals the critic output of the next step's policy pass
        // on the very same observation row - it is filled then.  Complete step n-1 now:
        if (P.prev_next_values && P.prev_dones[i] == 0.0f) P.prev_next_values[i] = v_now;
        // timed-out envs (reset without termination) were evaluated by the compact critic pass on their terminal observation
        if (P.c_count && i < (long long)*P.c_count) {
            float cv = P.c_value_raw[i];
            if (P.unnorm_value) cv = P.v_std * fminf(fmaxf(cv, -5.0f), 5.0f) + P.v_mean;
            P.mb_next_values[P.c_idx[i]] = cv;                           // these envs have term == 0
        }
    }
    float amp = disc_r(P.disc_logit[i], P.disc_scale);                    // :93
    P.mb_rewards[i] = shaped; P.mb_dones[i] = done;
    if (!deferred) P.mb_next_values[i] = nv;
    else if (term != 0.0f) P.mb_next_values[i] = 0.f;                     // next_vals *= 1 - terminated
    if (P.mb_amp_rewards) P.mb_amp_rewards[i] = amp;
    P.terminated_flags[i] += term;
    // ---- LocoVal target bookkeeping (:94-118) ----
    float cr = P.current_rewards[i] + r;
    float len = P.current_lengths[i] + 1.0f;
    float coef = P.discount_coefs[i];
    float cc = P.current_combined[i] + (shaped + amp) * coef;
    float not_done = 1.0f - done;
    bool done_early = (len <= P.step_to_pred) && done != 0.0f;
    bool over_pred = (len == P.step_to_pred) && not_done != 0.0f;
    P.game_combined[i] += cc * ((done_early || over_pred) ? 1.0f : 0.0f);
    P.current_combined[i] = cc * not_done;
    P.discount_coefs[i] = done != 0.0f ? 1.0f : coef * P.gamma;
    P.current_rewards[i] = cr * not_done;
    P.current_lengths[i] = len * not_done;
}

cudaError_t eml_rollout_record(const RecordParams& P, cudaStream_t st) {
    if (P.N <= 0) return cudaSuccess;
    rollout_record_kernel<<<(unsigned)((P.N + 127) / 128), 128, 0, st>>>(P);
    return cudaGetLastError();
}


// ---- RunningMeanStd.forward eval branch (utils/running_mean_std.py:82-84) for the self-obs slice that is concatenated
// with the task-MLP output (amp_network_sept_builder.py:75,95); the other slices are normalised inside the GEMM operand load
__global__ void normalize_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ y, long long ldy, long long M,
                                 int K, const float* __restrict__ mean, const float* __restrict__ var, float eps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * K) return;
    long long r = i / K; int k = (int)(i - r * K);
    float v = (x[r * ldx + k] - mean[k]) / sqrtf(var[k] + eps);
    y[r * ldy + k] = fminf(fmaxf(v, -5.0f), 5.0f);
}

cudaError_t eml_normalize(const float* x, long long ldx, float* y, long long ldy, long long M, int K, const float* mean,
                          const float* var, float eps, cudaStream_t st) {
    if (M <= 0 || K <= 0) return cudaSuccess;
    normalize_kernel<<<(unsigned)((M * K + 255) / 256), 256, 0, st>>>(x, ldx, y, ldy, M, K, mean, var, eps);
    return cudaGetLastError();
}

// ---- value reuse: compact the envs that were reset WITHOUT terminating (episode time-out) and gather their critic operands ----
// One warp per env; rows keep their 16-byte chunks.  `count` is zeroed by the launcher (memset node) before the kernel runs.
__global__ void timeout_gather_kernel(const int64_t* __restrict__ reset, const int64_t* __restrict__ terminate, long long N,
                                      const uint4* __restrict__ self_hi, const uint4* __restrict__ self_lo, int self_ld4,
                                      const uint4* __restrict__ task_hi, const uint4* __restrict__ task_lo, int task_ld4,
                                      uint4* c_self_hi, uint4* c_self_lo, int cself_ld4, uint4* c_task_hi, uint4* c_task_lo, int ctask_ld4,
                                      int32_t* idx, int32_t* count) {
    const long long env = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (env >= N) return;
    if (reset[env] == 0 || terminate[env] != 0) return;                  // warp-uniform
    int slot = 0;
    if (lane == 0) { slot = atomicAdd(count, 1); idx[slot] = (int32_t)env; }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    for (int i = lane; i < EML_SELF_OBS / 8; i += 32) {                   // 368 bf16 = 46 chunks of 16 bytes
        c_self_hi[(long long)slot * cself_ld4 + i] = self_hi[env * self_ld4 + i];
        c_self_lo[(long long)slot * cself_ld4 + i] = self_lo[env * self_ld4 + i];
    }
    for (int i = lane; i < task_ld4; i += 32) {                           // whole padded pitch of the task row
        c_task_hi[(long long)slot * ctask_ld4 + i] = task_hi[env * task_ld4 + i];
        c_task_lo[(long long)slot * ctask_ld4 + i] = task_lo[env * task_ld4 + i];
    }
}

cudaError_t eml_timeout_gather(const int64_t* reset, const int64_t* terminate, long long N, const uint16_t* self_hi,
                               const uint16_t* self_lo, long long ld_self, const uint16_t* task_hi, const uint16_t* task_lo,
                               long long ld_task, uint16_t* c_self_hi, uint16_t* c_self_lo, long long ld_cself, uint16_t* c_task_hi,
                               uint16_t* c_task_lo, long long ld_ctask, int32_t* idx, int32_t* count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess || N <= 0) return e;
    timeout_gather_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(
        reset, terminate, N, (const uint4*)self_hi, (const uint4*)self_lo, (int)(ld_self / 8), (const uint4*)task_hi, (const uint4*)task_lo,
        (int)(ld_task / 8), (uint4*)c_self_hi, (uint4*)c_self_lo, (int)(ld_cself / 8), (uint4*)c_task_hi, (uint4*)c_task_lo, (int)(ld_ctask / 8),
        idx, count);
    return cudaGetLastError();
}

// next_values of step n-1 for the envs that were not reset: this step's (un-normalised) critic output
__global__ void fill_next_values_kernel(const float* __restrict__ value_raw, const float* __restrict__ prev_dones,
                                        float* __restrict__ prev_next_values, long long N, float v_mean, float v_std,
                                        const float* __restrict__ v_stats, int unnorm) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || prev_dones[i] != 0.0f) return;
    if (v_stats) { v_mean = __ldg(v_stats); v_std = __ldg(v_stats + 1); }
    float v = value_raw[i];
    if (unnorm) v = v_std * fminf(fmaxf(v, -5.0f), 5.0f) + v_mean;
    prev_next_values[i] = v;
}
cudaError_t eml_fill_next_values(const float* value_raw, const float* prev_dones, float* prev_next_values, long long N, float v_mean,
                                 float v_std, const float* v_stats, int unnorm, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    fill_next_values_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(value_raw, prev_dones, prev_next_values, N, v_mean, v_std, v_stats, unnorm);
    return cudaGetLastError();
}

// value / logit heads fused into the producing GEMM (emloco_linear_bf16x3_head): the epilogue leaves one partial dot product per
// 64-column group; they are added here in column order (deterministic), plus the head's bias
__global__ void head_reduce_kernel(const float* __restrict__ part, int groups, const float* __restrict__ bias, float* __restrict__ out,
                                   long long M) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float a = 0.f;
    for (int g = 0; g < groups; ++g) a += part[i * groups + g];
    out[i] = a + (bias ? __ldg(bias) : 0.f);
}

cudaError_t eml_head_reduce(const float* part, int groups, const float* bias, float* out, long long M, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    head_reduce_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(part, groups, bias, out, M);
    return cudaGetLastError();
}


// ---- inference loop bookkeeping: AMPPlayerContinuousValue.run (pacer/pacer/learning/amp_value_players.py:123-198), all envs at once.
// Per env: LocoVal score kept from the episode's first step, discounted return ((r_loc + r_pow) * 0.5 + r_disc * 0.25) * gamma^(n+1)
// (or the penalised task reward when plot_val_reward is off), snapshot at n == step_to_pred or at an earlier end; a finished
// episode appends {env, pred, cr_to_pred, normalised return, c_loc, c_pow, c_disc, steps} to the result list.
// state [11,N]: n, cr, coef, pred, cr_to_pred, c_loc, c_pow, c_disc, loc_to_pred, pow_to_pred, disc_to_pred
__global__ void player_record_kernel(const float* __restrict__ rew, const float* __restrict__ rew_raw, const int64_t* __restrict__ reset,
                                     const float* __restrict__ logit, const float* __restrict__ scores, const uint8_t* __restrict__ inverted,
                                     float* __restrict__ st, long long N, float* __restrict__ results, int* __restrict__ count, int capacity,
                                     int plot_val_reward, float inv_penalty, float disc_scale, float gamma, float step_to_pred,
                                     float min_reward, float max_reward) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float* S[11];
#pragma unroll
    for (int q = 0; q < 11; ++q) S[q] = st + (long long)q * N + i;
    const float n = *S[0];
    float pred = *S[3];
    if (n == 0.f) pred = scores[i];                                       // :128-137
    const float disc = disc_r(logit[i], disc_scale);
    const float coef = *S[2] * gamma;                                     // :144, before use
    float cr = *S[1], c_loc = *S[5], c_pow = *S[6], c_disc = *S[7];
    if (plot_val_reward) {                                                // :145-160
        const float r_loc = rew_raw[2 * i], r_pow = rew_raw[2 * i + 1];
        c_disc += disc * 0.25f * coef; c_loc += r_loc * 0.5f * coef; c_pow += r_pow * 0.5f * coef;
        cr += ((r_loc + r_pow) * 0.5f + disc * 0.25f) * coef;
    } else {                                                              // :161-163 with the penalty of :127
        float r = rew[i];
        if (inverted && inverted[i]) r *= -inv_penalty;
        cr += r * coef;
    }
    const bool done = reset[i] != 0;
    float ctp = *S[4], ltp = *S[8], ptp = *S[9], dtp = *S[10];
    if (n == step_to_pred || (done && n < step_to_pred)) { ctp = cr; ltp = c_loc; ptp = c_pow; dtp = c_disc; }   // :177-193
    if (done) {
        const int slot = atomicAdd(count, 1);
        if (slot < capacity) {
            float* o = results + (long long)slot * 8;
            o[0] = (float)i; o[1] = pred; o[2] = ctp; o[3] = (ctp - min_reward) / (max_reward - min_reward);    // :195
            o[4] = ltp; o[5] = ptp; o[6] = dtp; o[7] = n + 1.0f;
        }
        cr = 0.f; c_loc = 0.f; c_pow = 0.f; c_disc = 0.f;                  // the next game starts from scratch (:76-100)
    }
    *S[0] = done ? 0.f : n + 1.0f; *S[1] = cr; *S[2] = done ? 1.0f : coef; *S[3] = pred; *S[4] = ctp;
    *S[5] = c_loc; *S[6] = c_pow; *S[7] = c_disc; *S[8] = ltp; *S[9] = ptp; *S[10] = dtp;
}

cudaError_t eml_player_record(const float* rew, const float* rew_raw, const int64_t* reset, const float* logit, const float* scores,
                              const uint8_t* inverted, float* st, long long N, float* results, int* count, int capacity,
                              int plot_val_reward, float inv_penalty, float disc_scale, float gamma, int step_to_pred, float min_reward,
                              float max_reward, cudaStream_t stream) {
    if (N <= 0) return cudaSuccess;
    player_record_kernel<<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(rew, rew_raw, reset, logit, scores, inverted, st, N, results, count,
                                                                         capacity, plot_val_reward, inv_penalty, disc_scale, gamma,
                                                                         (float)step_to_pred, min_reward, max_reward);
    return cudaGetLastError();
}
