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
__global__ void __launch_bounds__(128) sample_actions_kernel(const float* __restrict__ mu, long long ldmu,
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
    sample_actions_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(mu, ldmu, logstd, noise, actions, neglogp, N, A);
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
    if (P.inverted && P.inverted[i]) r *= -P.inv_penalty;                 // :63-64
    float shaped = r * P.reward_scale;                                    // rewards_shaper (scale_value 1)
    float done = P.reset[i] != 0 ? 1.0f : 0.0f;
    float term = (float)P.terminate[i];
    float nv = P.next_value_raw[i];
    if (P.unnorm_value) nv = P.v_std * fminf(fmaxf(nv, -5.0f), 5.0f) + P.v_mean;   // running_mean_std.py:77-79
    if (P.value_raw) {                                                    // res_dict['values'] of get_action_values
        float v = P.value_raw[i];
        if (P.unnorm_value) v = P.v_std * fminf(fmaxf(v, -5.0f), 5.0f) + P.v_mean;
        P.mb_values[i] = v;
    }
    nv *= (1.0f - term);                                                  // :86-90
    float amp = disc_r(P.disc_logit[i], P.disc_scale);                    // :93
    P.mb_rewards[i] = shaped; P.mb_dones[i] = done; P.mb_next_values[i] = nv;
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
