// GAE reverse scan: discount_values (reference pacer/pacer/learning/common_agent.py:573-587) plus
// "mb_returns = mb_advs + mb_values" (amp_continuous_value.py:163) in one launch instead of a
// 32-iteration Python loop of ~6 eager kernels.  One thread per env walks t = T-1..0; at a fixed t
// consecutive threads touch consecutive addresses, so every access is coalesced.  24 B per (t, env).
#include "sim.h"

__global__ void gae_kernel(const float* __restrict__ dones, const float* __restrict__ values,
                           const float* __restrict__ rewards, const float* __restrict__ next_values,
                           float* __restrict__ adv, float* __restrict__ ret, int T, long long N, float gamma, float tau) {
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float last = 0.f;
    const float gt = gamma * tau;
    for (int t = T - 1; t >= 0; --t) {
        long long i = (long long)t * N + n;
        float v = values[i];
        float not_done = 1.0f - dones[i];
        float delta = rewards[i] + gamma * next_values[i] - v;
        last = delta + gt * not_done * last;
        adv[i] = last;
        if (ret) ret[i] = last + v;
    }
}

cudaError_t eml_gae(const float* dones, const float* values, const float* rewards, const float* next_values,
                    float* adv, float* ret, int T, long long N, float gamma, float tau, cudaStream_t st) {
    if (N <= 0 || T <= 0) return cudaSuccess;
    int threads = 64;   // 4096 envs -> 64 CTAs; small CTAs spread the scan over more SMs
    long long blocks = (N + threads - 1) / threads;
    gae_kernel<<<(unsigned)blocks, threads, 0, st>>>(dones, values, rewards, next_values, adv, ret, T, N, gamma, tau);
    return cudaGetLastError();
}
