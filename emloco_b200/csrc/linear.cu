// Dense layer y = act(norm(x) W^T + b) for the actor / critic / discriminator MLPs:
// reference = nn.Linear stacks built by pacer/pacer/learning/network_builder.py / amp_network_*builder.py
// preceded by RunningMeanStd.forward (utils/running_mean_std.py:60-84).
//
// This file holds the portable fp32 FMA path (exact fp32 products, used for parity and for the tiny
// layers); linear_tc.cu holds the tcgen05 TF32 tensor-core path for the wide layers.
// The input normalisation clamp((x-mean)/sqrt(var+eps),+-5) is fused into the A-operand load, bias and ReLU
// into the epilogue, and the row strides ldx/ldy let callers read/write slices of wider buffers so the
// `torch.cat([self_obs, task_out])` of amp_network_sept_builder.py:75,95 never materialises.
#include "sim.h"

#define LT_BM 64
#define LT_BN 64
#define LT_BK 16

__global__ void __launch_bounds__(256) linear_fma_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, long long ldy,
                                                         long long M, int N, int K, const float* __restrict__ mean,
                                                         const float* __restrict__ var, float eps, int relu) {
    __shared__ float sa[LT_BK][LT_BM + 4];
    __shared__ float sb[LT_BK][LT_BN + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.y * LT_BM;
    const int n0 = blockIdx.x * LT_BN;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += LT_BK) {
        // load A tile (64 rows x 16 k) and B tile (64 cols x 16 k): 1024 elements each, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + 256 * i;
            int r = e >> 4, kk = e & 15;
            int k = k0 + kk;
            float a = 0.f, bv = 0.f;
            if (k < K) {
                if (m0 + r < M) {
                    a = x[(m0 + r) * ldx + k];
                    if (mean) {
                        a = (a - mean[k]) / sqrtf(var[k] + eps);
                        a = fminf(fmaxf(a, -5.0f), 5.0f);
                    }
                }
                if (n0 + r < N) bv = w[(long long)(n0 + r) * K + k];
            }
            sa[kk][r] = a;
            sb[kk][r] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LT_BK; ++kk) {
            float4 a4 = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
            float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            y[m * ldy + n] = v;
        }
    }
}

cudaError_t eml_linear_fma(const float* x, long long ldx, const float* w, const float* b, float* y, long long ldy, long long M,
                           int N, int K, const float* mean, const float* var, float eps, int relu, cudaStream_t st) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    dim3 grid((N + LT_BN - 1) / LT_BN, (unsigned)((M + LT_BM - 1) / LT_BM));
    linear_fma_kernel<<<grid, 256, 0, st>>>(x, ldx, w, b, y, ldy, M, N, K, mean, var, eps, relu);
    return cudaGetLastError();
}
