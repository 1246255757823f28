// Element-wise / reduction kernels of the PPO / AMP update step: everything of
// `AMPValueAgent.calc_gradients` (pacer/pacer/learning/amp_continuous_value.py:276-428) that is not a dense-layer GEMM.
// The GEMMs (forward, dgrad, wgrad, and the double-backward chain of the discriminator's gradient penalty) run on the
// tcgen05 bf16x3 kernel of linear_tc.cu; the kernels here produce its operands (fp32 -> bf16 hi/lo in both orientations),
// apply ReLU masks, seed the backward pass from the losses, reduce bias gradients and apply clip-norm + Adam.
//
//   xform_kernel          y = scale * rowscale[m] * src[m,k] * [gate[m,k] > 0],  src = norm(x) | x | rowvec[k]
//                         -> any of: fp32 y, fp32 y^T, split y (hi/lo), split y^T, column sums, sum of squares
//   ppo_head_kernel       actor / critic / task-value / bound losses and their gradients w.r.t. mu, value, task value
//                         (common_agent.py:594-602,657-683, amp_continuous_value.py:430-444; neglogp / entropy / kl:
//                         rl_games 1.1.4 ModelA2CContinuousLogStd + torch_ext.policy_kl, restated)
//   disc_head_kernel      BCE-with-logits prediction loss of the discriminator and d loss / d logit (amp_continuous.py:536-616)
//   amp_dropout_kernel    whole-joint dropout masks of the AMP observations (amp_models.py:49-90)
//   column_moments / rms_update   RunningMeanStd training-mode update (utils/running_mean_std.py:33-43,86-96)
//   sumsq / adam_clip     nn.utils.clip_grad_norm_(50) + torch.optim.Adam step over flat parameter / gradient buffers
#include "sim.h"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

struct XformParams {
    const float* x; long long ldx;          // [M,K] or NULL
    const float* rowvec;                    // [K] (used when x == NULL)
    const float* rowscale; long long lds;   // [M] (stride lds) or NULL
    const float* mean; const float* var; float eps;   // optional normalisation of x: clamp((x-mean)/sqrt(var+eps), +-5)
    const float* gate; long long ldg;       // [M,K] or NULL: pass where gate > 0
    const float* drop_u; float drop_rate;   // [M,19] or NULL: whole-joint dropout of AMP observation columns (amp_models.py:49-90):
                                            // column k of joint j passes where drop_u[m,j] > drop_rate (computed inline, no mask tensor)
    float scale;
    float* y32; long long ldy;              // [M,K]
    float* yT32; long long ldyT;            // [K,M]
    __nv_bfloat16* hi; __nv_bfloat16* lo; long long ld16;       // [M,K]
    __nv_bfloat16* hiT; __nv_bfloat16* loT; long long ldT;      // [K,M]
    float* colsum;                          // [K] += sum over rows
    float* sumsq;                           // [1] += sum of squares
    long long M; int K;
    int xg;                                 // m-tiles per block group (see xform_tile)
};

constexpr int XT = 64;     // tile edge

// Block -> tile.  Blocks that run at the same time should touch neighbouring memory in BOTH orientations: xg consecutive blocks
// take xg consecutive m-tiles of one k-tile (their transposed stores fill xg x 128 contiguous bytes of every k-row), the next
// xg blocks take the next k-tile of the same rows (their row-major accesses continue where the previous group stopped).
__device__ __forceinline__ bool xform_tile(const XformParams& P, int& kt, int& mt) {
    const int gx = (P.K + XT - 1) / XT;
    const long long gy = (P.M + XT - 1) / XT;
    const long long g = blockIdx.x / P.xg;
    kt = (int)(g % gx);
    mt = (int)((g / gx) * P.xg + blockIdx.x % P.xg);
    return mt < gy;
}

__global__ void __launch_bounds__(256) xform_kernel(XformParams P) {
    __shared__ float tile[XT][XT + 1];
    __shared__ float red[8];
    const int tid = threadIdx.x;
    int kt, mt;
    if (!xform_tile(P, kt, mt)) return;
    const long long m0 = (long long)mt * XT;
    const int k0 = kt * XT;
    const int c = tid & 63, r0 = tid >> 6;              // column within the tile, first row; rows r0, r0+4, ...
    const int k = k0 + c;
    float mu = 0.f, is = 1.f, rv = 0.f;
    const bool kin = k < P.K;
    int jd = -1;
    if (kin) {
        if (P.mean) { mu = P.mean[k]; is = 1.0f / sqrtf(P.var[k] + P.eps); }
        if (!P.x) rv = P.rowvec[k];
        if (P.drop_u) {
            const int f = k % EML_AMP_STEP;
            if (f >= 12 && f < 12 + 19 * 6) jd = (f - 12) / 6;
            else if (f >= 126 && f < 126 + 19 * 3) jd = (f - 126) / 3;
        }
    }
    float ss = 0.f;
#pragma unroll 4
    for (int i = 0; i < XT / 4; ++i) {
        const int r = r0 + 4 * i;
        const long long m = m0 + r;
        float v = 0.f;
        if (kin && m < P.M) {
            if (P.x) {
                v = P.x[m * P.ldx + k];
                if (P.mean) v = fminf(fmaxf((v - mu) * is, -5.0f), 5.0f);
            } else v = rv;
            if (P.rowscale) v *= P.rowscale[m * P.lds];
            v *= P.scale;
            if (P.gate && !(P.gate[m * P.ldg + k] > 0.f)) v = 0.f;
            if (jd >= 0 && !(P.drop_u[m * 19 + jd] > P.drop_rate)) v = 0.f;
            if (P.y32) P.y32[m * P.ldy + k] = v;
            if (P.hi) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                P.hi[m * P.ld16 + k] = h;
                P.lo[m * P.ld16 + k] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
            ss += v * v;
        }
        tile[r][c] = v;
    }
    __syncthreads();
    if (P.yT32 || P.hiT) {
        // transposed write: thread (c, r0) now owns row index c of the tile (an m) and columns r0, r0+4, ... (k's)
        const long long m = m0 + c;
        if (m < P.M) {
#pragma unroll 4
            for (int i = 0; i < XT / 4; ++i) {
                const int kk = r0 + 4 * i;
                if (k0 + kk >= P.K) break;
                const float v = tile[c][kk];
                if (P.yT32) P.yT32[(long long)(k0 + kk) * P.ldyT + m] = v;
                if (P.hiT) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(v);
                    P.hiT[(long long)(k0 + kk) * P.ldT + m] = h;
                    P.loT[(long long)(k0 + kk) * P.ldT + m] = __float2bfloat16_rn(v - __bfloat162float(h));
                }
            }
        }
    }
    if (P.colsum && tid < XT && k0 + tid < P.K) {
        float a = 0.f;
#pragma unroll 8
        for (int r = 0; r < XT; ++r) a += tile[r][tid];
        atomicAdd(P.colsum + k0 + tid, a);
    }
    if (P.sumsq) {
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        if (tid == 0) { float a = 0.f; for (int w = 0; w < 8; ++w) a += red[w]; atomicAdd(P.sumsq, a); }
    }
}

// joint of AMP-observation column k (206 features per history step: 12 root features, 19 x 6 joint rotations, 19 x 3 joint
// velocities, 12 key-body + 11 shape features), or -1 when the column is never dropped
__device__ __forceinline__ int amp_joint_of(int k) {
    const int f = k % EML_AMP_STEP;
    if (f >= 12 && f < 12 + 19 * 6) return (f - 12) / 6;
    if (f >= 126 && f < 126 + 19 * 3) return (f - 126) / 3;
    return -1;
}

// Same operation, two adjacent columns per thread: 8-byte loads, packed bf16x2 stores in both orientations.  Launched when
// K, the pitches and the base pointers allow it (eml_xform checks); the scalar kernel above is the general path.
__global__ void __launch_bounds__(256) xform2_kernel(XformParams P) {
    __shared__ float tile[XT][XT + 1];
    __shared__ float red[8];
    const int tid = threadIdx.x;
    int kt, mt;
    if (!xform_tile(P, kt, mt)) return;
    const long long m0 = (long long)mt * XT;
    const int k0 = kt * XT;
    const int c = tid & 31, r0 = tid >> 5;               // column pair within the tile, first row; rows r0, r0+8, ...
    const int k = k0 + 2 * c;
    const bool kin = k < P.K;                            // K is even: the pair is in or out together
    float2 mu = make_float2(0.f, 0.f), is = make_float2(1.f, 1.f), rv = make_float2(0.f, 0.f);
    int j0 = -1, j1 = -1;
    if (kin) {
        if (P.mean) {
            mu = *reinterpret_cast<const float2*>(P.mean + k);
            const float2 vr = *reinterpret_cast<const float2*>(P.var + k);
            is = make_float2(1.0f / sqrtf(vr.x + P.eps), 1.0f / sqrtf(vr.y + P.eps));
        }
        if (!P.x) rv = *reinterpret_cast<const float2*>(P.rowvec + k);
        if (P.drop_u) { j0 = amp_joint_of(k); j1 = amp_joint_of(k + 1); }
    }
    float ss = 0.f;
#pragma unroll 4
    for (int i = 0; i < XT / 8; ++i) {
        const int r = r0 + 8 * i;
        const long long m = m0 + r;
        float2 v = make_float2(0.f, 0.f);
        if (kin && m < P.M) {
            if (P.x) {
                v = *reinterpret_cast<const float2*>(P.x + m * P.ldx + k);
                if (P.mean) {
                    v.x = fminf(fmaxf((v.x - mu.x) * is.x, -5.0f), 5.0f);
                    v.y = fminf(fmaxf((v.y - mu.y) * is.y, -5.0f), 5.0f);
                }
            } else v = rv;
            float sc = P.scale;
            if (P.rowscale) sc *= P.rowscale[m * P.lds];
            v.x *= sc; v.y *= sc;
            if (P.gate) {
                const float2 g = *reinterpret_cast<const float2*>(P.gate + m * P.ldg + k);
                if (!(g.x > 0.f)) v.x = 0.f;
                if (!(g.y > 0.f)) v.y = 0.f;
            }
            if (P.drop_u) {
                if (j0 >= 0 && !(P.drop_u[m * 19 + j0] > P.drop_rate)) v.x = 0.f;
                if (j1 >= 0 && !(P.drop_u[m * 19 + j1] > P.drop_rate)) v.y = 0.f;
            }
            if (P.y32) *reinterpret_cast<float2*>(P.y32 + m * P.ldy + k) = v;
            if (P.hi) {
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v.x - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v.y - __bfloat162float(h1));
                *reinterpret_cast<uint32_t*>(P.hi + m * P.ld16 + k) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                *reinterpret_cast<uint32_t*>(P.lo + m * P.ld16 + k) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            ss += v.x * v.x + v.y * v.y;
        }
        tile[r][2 * c] = v.x; tile[r][2 * c + 1] = v.y;
    }
    __syncthreads();
    if (P.yT32 || P.hiT) {
        // transposed write: thread (c, r0) owns rows 2c, 2c+1 of the tile (two adjacent m) and columns r0, r0+8, ... (k's)
        const long long m = m0 + 2 * c;
        if (m < P.M) {                                  // M is even: the pair is in or out together
#pragma unroll 4
            for (int i = 0; i < XT / 8; ++i) {
                const int kk = r0 + 8 * i;
                if (k0 + kk >= P.K) break;
                const float a = tile[2 * c][kk], b = tile[2 * c + 1][kk];
                if (P.yT32) *reinterpret_cast<float2*>(P.yT32 + (long long)(k0 + kk) * P.ldyT + m) = make_float2(a, b);
                if (P.hiT) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(a), h1 = __float2bfloat16_rn(b);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(a - __bfloat162float(h0)), l1 = __float2bfloat16_rn(b - __bfloat162float(h1));
                    *reinterpret_cast<uint32_t*>(P.hiT + (long long)(k0 + kk) * P.ldT + m) = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    *reinterpret_cast<uint32_t*>(P.loT + (long long)(k0 + kk) * P.ldT + m) = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                }
            }
        }
    }
    if (P.colsum && tid < XT && k0 + tid < P.K) {
        float a = 0.f;
#pragma unroll 8
        for (int r = 0; r < XT; ++r) a += tile[r][tid];
        atomicAdd(P.colsum + k0 + tid, a);
    }
    if (P.sumsq) {
        ss = warp_sum(ss);
        if ((tid & 31) == 0) red[tid >> 5] = ss;
        __syncthreads();
        if (tid == 0) { float a = 0.f; for (int w = 0; w < 8; ++w) a += red[w]; atomicAdd(P.sumsq, a); }
    }
}

// ---- PPO heads: one warp per sample --------------------------------------------------------------------------------
struct PpoHeadParams {
    const float* mu; long long ldmu;        // [B,A] new mu
    const float* logstd;                    // [A]
    const float* actions;                   // [B,A] prev_actions
    const float* old_neglogp;               // [B]
    const float* adv;                       // [B]
    const float* value; const float* task_value; const float* returns;     // [B] (normalised value space)
    const float* old_mu; const float* old_sigma;                           // [B,A] (kl, info only) or NULL
    float* dmu; long long lddmu; float* dvalue; float* dtask;              // gradients of the TOTAL loss
    float* stats;                           // [8] += a_loss, c_loss, tv_loss, b_loss, clipped, kl, entropy, (unused)
    long long B; int A;
    float e_clip, actor_coef, critic_coef, tv_coef, bounds_coef, inv_B;
};

__global__ void __launch_bounds__(128) ppo_head_kernel(PpoHeadParams P) {
    __shared__ float red[4][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * 4 + w;
    float st[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (b < P.B) {
        float q = 0.f, ls = 0.f, bl = 0.f, kl = 0.f;
        for (int j = lane; j < P.A; j += 32) {
            const float l = P.logstd[j], s = expf(l);
            const float m = P.mu[b * P.ldmu + j];
            const float z = (P.actions[b * P.A + j] - m) / s;
            q += z * z; ls += l;
            const float hi_ = fmaxf(m - 1.0f, 0.f), lo_ = fminf(m + 1.0f, 0.f);          // bound_loss, soft_bound 1 (:594-602)
            bl += hi_ * hi_ + lo_ * lo_;
            if (P.old_mu) {                                                                // torch_ext.policy_kl(p0 = new, p1 = old)
                const float s1 = P.old_sigma[b * P.A + j], m1 = P.old_mu[b * P.A + j];
                kl += logf(s1 / s + 1e-5f) + (s * s + (m1 - m) * (m1 - m)) / (2.0f * (s1 * s1 + 1e-5f)) - 0.5f;
            }
        }
        q = warp_sum(q); ls = warp_sum(ls); bl = warp_sum(bl); kl = warp_sum(kl);
        const float neglogp = 0.5f * q + 0.91893853320467274178f * (float)P.A + ls;
        const float A_ = P.adv[b];
        const float ratio = expf(P.old_neglogp[b] - neglogp);                              // _actor_loss (:657-669)
        const float rc = fminf(fmaxf(ratio, 1.0f - P.e_clip), 1.0f + P.e_clip);
        const float s1 = -A_ * ratio, s2 = -A_ * rc;
        const float a_loss = fmaxf(s1, s2);
        // d max(s1, s2) / d ratio: inside the clip range both branches are the same function of ratio (torch.maximum halves the
        // gradient on the tie, the clamp passes the other half); outside, s2 is constant and s1 carries it only when it wins
        const bool inside = ratio >= 1.0f - P.e_clip && ratio <= 1.0f + P.e_clip;
        const float dl_dratio = (inside || s1 > s2) ? -A_ : (s1 == s2 ? -0.5f * A_ : 0.f);
        const float dl_dneglogp = -ratio * dl_dratio * P.actor_coef * P.inv_B;
        for (int j = lane; j < P.A; j += 32) {
            const float l = P.logstd[j], s = expf(l);
            const float m = P.mu[b * P.ldmu + j];
            const float dn_dm = -(P.actions[b * P.A + j] - m) / (s * s);
            const float hi_ = fmaxf(m - 1.0f, 0.f), lo_ = fminf(m + 1.0f, 0.f);
            P.dmu[b * P.lddmu + j] = dl_dneglogp * dn_dm + P.bounds_coef * P.inv_B * 2.0f * (hi_ + lo_);
        }
        if (lane == 0) {
            const float dv = P.returns[b] - P.value[b], dt = P.returns[b] - P.task_value[b];
            P.dvalue[b] = -2.0f * dv * P.critic_coef * P.inv_B;                            // _critic_loss, clip_value False (:671-683)
            P.dtask[b] = -2.0f * dt * P.tv_coef * P.inv_B;                                 // _task_value_loss (:430-444)
            st[0] = a_loss; st[1] = dv * dv; st[2] = dt * dt; st[3] = bl;
            st[4] = fabsf(ratio - 1.0f) > P.e_clip ? 1.f : 0.f; st[5] = kl;
            st[6] = ls + (float)P.A * (0.5f + 0.91893853320467274178f);                    // Normal.entropy().sum(-1)
        }
    }
    if (lane == 0) for (int i = 0; i < 8; ++i) red[w][i] = st[i];
    __syncthreads();
    if (threadIdx.x < 7) atomicAdd(P.stats + threadIdx.x, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// ---- discriminator prediction loss: rows [0, n_agent) are fakes (agent + replay), rows [n_agent, n_agent + n_demo) real ----
__global__ void disc_head_kernel(const float* __restrict__ logit, float* __restrict__ dlogit, float* __restrict__ stats,
                                 long long n_agent, long long n_demo, float coef) {
    __shared__ float red[8][4];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < n_agent + n_demo) {
        const float l = logit[i];
        const float sp_pos = fmaxf(l, 0.f) + log1pf(expf(-fabsf(l)));       // softplus(l)  = BCEWithLogits(l, 0)
        const float sig = 1.0f / (1.0f + expf(-l));
        if (i < n_agent) {
            st[0] = sp_pos; st[2] = l < 0.f ? 1.f : 0.f;
            dlogit[i] = coef * 0.5f * sig / (float)n_agent;
        } else {
            st[1] = sp_pos - l; st[3] = l > 0.f ? 1.f : 0.f;                 // softplus(-l) = BCEWithLogits(l, 1)
            dlogit[i] = -coef * 0.5f * (1.0f - sig) / (float)n_demo;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) st[q] = warp_sum(st[q]);
    if ((threadIdx.x & 31) == 0) for (int q = 0; q < 4; ++q) red[threadIdx.x >> 5][q] = st[q];
    __syncthreads();
    if (threadIdx.x < 4) {
        float a = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += red[w][threadIdx.x];
        atomicAdd(stats + threadIdx.x, a);
    }
}

// ---- amp_models.py:49-90 get_dropout_mask: per sample and per joint (19), one keep / drop decision shared by the joint's
// 6 rotation features and 3 velocity features in every one of the 15 history steps.  u [rows, 19] uniform draws.
__global__ void amp_dropout_kernel(const float* __restrict__ u, float* __restrict__ mask, long long rows, float rate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * EML_AMP_OBS) return;
    const long long r = i / EML_AMP_OBS;
    const int f = (int)(i - r * EML_AMP_OBS) % EML_AMP_STEP;
    float keep = 1.f;
    int j = -1;
    if (f >= 12 && f < 12 + 19 * 6) j = (f - 12) / 6;
    else if (f >= 126 && f < 126 + 19 * 3) j = (f - 126) / 3;
    if (j >= 0) keep = u[r * 19 + j] > rate ? 1.f : 0.f;
    mask[i] = keep;
}

// ---- RunningMeanStd training update ----
__global__ void __launch_bounds__(256) column_moments_kernel(const float* __restrict__ x, long long ldx, long long M, int K,
                                                             double* __restrict__ sum, double* __restrict__ sumsq) {
    // block = 32 columns x 8 row groups; grid.y strides over rows
    __shared__ double s1[8][33], s2[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + c;
    double a = 0.0, b = 0.0;
    if (k < K)
        for (long long m = (long long)blockIdx.y * 8 + g; m < M; m += (long long)gridDim.y * 8) {
            const double v = (double)x[m * ldx + k];
            a += v; b += v * v;
        }
    s1[g][c] = a; s2[g][c] = b;
    __syncthreads();
    if (g == 0 && k < K) {
        for (int q = 1; q < 8; ++q) { a += s1[q][c]; b += s2[q][c]; }
        atomicAdd(sum + k, a); atomicAdd(sumsq + k, b);
    }
}

// one block: _update_mean_var_count_from_moments (running_mean_std.py:33-43); batch mean / var enter as fp32 like the reference's
// `input.mean(0)` / `input.var(0)` (unbiased); also refreshes the fp32 copies the kernels read and clears the accumulators
__global__ void rms_update_kernel(double* __restrict__ sum, double* __restrict__ sumsq, long long M, int K, double* __restrict__ rmean,
                                  double* __restrict__ rvar, double* __restrict__ count, float* __restrict__ mean32,
                                  float* __restrict__ var32, float* __restrict__ inv32, float eps) {
    __shared__ double s_count;
    if (threadIdx.x == 0) s_count = *count;
    __syncthreads();
    const double cnt = s_count, bc = (double)M, tot = cnt + bc;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double mean_d = sum[k] / bc;
        const double var_d = M > 1 ? (sumsq[k] - sum[k] * mean_d) / (bc - 1.0) : 0.0;
        const float bm = (float)mean_d, bv = (float)fmax(var_d, 0.0);
        const double delta = (double)bm - rmean[k];
        const double nm = rmean[k] + delta * bc / tot;
        const double m2 = rvar[k] * cnt + (double)(bv * (float)bc) + delta * delta * cnt * bc / tot;
        const double nv = m2 / tot;
        rmean[k] = nm; rvar[k] = nv;
        if (mean32) { mean32[k] = (float)nm; var32[k] = (float)nv; inv32[k] = 1.0f / sqrtf((float)nv + eps); }
        sum[k] = 0.0; sumsq[k] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) *count = tot;
}

// ---- clip_grad_norm_ + Adam ----
// Deterministic (fixed grid, fixed summation order): every rank of a data-parallel run computes bit-identical norms from the
// bit-identical all-reduced gradient, so the parameters never drift apart (float atomics would differ in the last bit).
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
    __shared__ float red[8];
    float a = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { const float v = g[i]; a += v * v; }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; partials[blockIdx.x] = s; }
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int count, float* __restrict__ out) {
    __shared__ float red[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < count; i += 256) a += partials[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *out += red[0];
}

// state: [0] step count (float), [1] sum of squares of the gradient (input, cleared by the last block... no: cleared by the caller)
__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const float* __restrict__ state, float lr,
                                                        float beta1, float beta2, float eps, float max_norm, float grad_scale) {
    const float t = state[0];
    // nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1; grad_scale = 1 / world size (gradient average)
    float coef = grad_scale;
    if (max_norm > 0.f) coef *= fminf(max_norm / (sqrtf(state[1]) * grad_scale + 1e-6f), 1.0f);
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, rs_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * rs_bc2 + eps);
    }
}

__global__ void bump_step_kernel(float* state) { state[0] += 1.0f; state[1] = 0.f; }

// ---- data-parallel optimiser step fused with its collective over NVLink / NVSwitch multicast memory (NVLS) --------------------
// The flat gradient and parameter buffers of all ranks are mapped behind ONE multicast address each (torch symmetric memory).
// Rank r owns the slice [lo, lo + count) of the flat vector:
//   dp_reduce_shard   multimem.ld_reduce: the switch adds the W ranks' gradients of the slice (one load per 16 bytes, no
//                     reduce-scatter round trips); the slice's sum of squares goes, by multimem.st, into slot r of every
//                     rank's exchange buffer
//   dp_adam_shard     every rank adds the W slots in rank order (same bits everywhere), clips, runs Adam on ITS slice only
//                     (moments are sharded: 1/W of the optimiser state per rank) and broadcasts the new parameters with
//                     multimem.st - the all-gather happens inside the store
// Replaces: NCCL all-reduce (44.8 MB) + norm + Adam over the whole vector on every rank.  Cross-rank ordering: symmetric-memory
// barriers issued on the stream around the two kernels (update.py).
__device__ __forceinline__ float4 mm_ld_reduce_f4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_f4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256) dp_reduce_shard_kernel(const float* __restrict__ mc_grad, float* __restrict__ shard, long long lo,
                                                              long long count4, float* __restrict__ partials) {
    __shared__ float red[8];
    float a = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = mm_ld_reduce_f4(mc_grad + lo + 4 * i);
        reinterpret_cast<float4*>(shard)[i] = v;
        a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; partials[blockIdx.x] = s; }
}
__global__ void __launch_bounds__(256) dp_publish_sumsq_kernel(const float* __restrict__ partials, int count, float* __restrict__ mc_exchange, int rank) {
    __shared__ float red[256];
    float a = 0.f;
    for (int i = threadIdx.x; i < count; i += 256) a += partials[i];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { mm_st_f4(mc_exchange + 4 * rank, make_float4(red[0], 0.f, 0.f, 0.f)); __threadfence_system(); }
}
__global__ void __launch_bounds__(256) dp_adam_shard_kernel(float* __restrict__ mc_param, const float* __restrict__ p_local,
                                                            const float* __restrict__ shard, float* __restrict__ m, float* __restrict__ v,
                                                            long long lo, long long count4, const float* __restrict__ exchange, int world,
                                                            float* __restrict__ state, float lr, float beta1, float beta2, float eps,
                                                            float max_norm, float grad_scale) {
    float total = 0.f;
    for (int r = 0; r < world; ++r) total += exchange[4 * r];                    // rank order: identical bits on every rank
    const float t = state[0];
    float coef = grad_scale;
    if (max_norm > 0.f) coef *= fminf(max_norm / (sqrtf(total) * grad_scale + 1e-6f), 1.0f);
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, rs_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g4 = reinterpret_cast<const float4*>(shard)[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        float4 p4 = *reinterpret_cast<const float4*>(p_local + lo + 4 * i);
        const float g[4] = {g4.x * coef, g4.y * coef, g4.z * coef, g4.w * coef};
        float* mm = &m4.x; float* vv = &v4.x; float* pp = &p4.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            mm[q] = beta1 * mm[q] + (1.0f - beta1) * g[q];
            vv[q] = beta2 * vv[q] + (1.0f - beta2) * g[q] * g[q];
            pp[q] -= step_size * mm[q] / (sqrtf(vv[q]) * rs_bc2 + eps);
        }
        reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4;
        mm_st_f4(mc_param + lo + 4 * i, p4);                                      // lands in every rank's parameter buffer
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = total;
    __threadfence_system();
}

// out[i] = sum_s parts[s * stride + i] (+ out[i] when accumulate): the partial matrices of a split-K GEMM, added in split order
__global__ void sum_parts_kernel(const float* __restrict__ parts, int S, long long stride, float* __restrict__ out, long long n, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a = accumulate ? out[i] : 0.f;
        for (int s = 0; s < S; ++s) a += parts[(long long)s * stride + i];
        out[i] = a;
    }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}

}  // namespace

cudaError_t eml_xform(const float* x, long long ldx, const float* rowvec, const float* rowscale, long long lds, const float* mean,
                      const float* var, float eps, const float* gate, long long ldg, const float* drop_u, float drop_rate, float scale,
                      float* y32, long long ldy, float* yT32, long long ldyT, void* hi, void* lo, long long ld16, void* hiT, void* loT,
                      long long ldT, float* colsum, float* sumsq, long long M, int K, cudaStream_t st) {
    if (M <= 0 || K <= 0) return cudaSuccess;
    XformParams P;
    P.x = x; P.ldx = ldx; P.rowvec = rowvec; P.rowscale = rowscale; P.lds = lds; P.mean = mean; P.var = var; P.eps = eps;
    P.gate = gate; P.ldg = ldg; P.drop_u = drop_u; P.drop_rate = drop_rate; P.scale = scale; P.y32 = y32; P.ldy = ldy; P.yT32 = yT32; P.ldyT = ldyT;
    P.hi = (__nv_bfloat16*)hi; P.lo = (__nv_bfloat16*)lo; P.ld16 = ld16; P.hiT = (__nv_bfloat16*)hiT; P.loT = (__nv_bfloat16*)loT; P.ldT = ldT;
    P.colsum = colsum; P.sumsq = sumsq; P.M = M; P.K = K;
    const long long gx = (K + XT - 1) / XT, gy = (M + XT - 1) / XT;
    static const int xg_env = [] { const char* e = getenv("EMLOCO_XG"); int v = e ? atoi(e) : 0; return v > 0 ? v : 8; }();
    P.xg = xg_env;
    const unsigned grid = (unsigned)(gx * ((gy + P.xg - 1) / P.xg) * P.xg);
    // two-column kernel: needs even K (and even M for the transposed outputs), even pitches, 8-byte (fp32) / 4-byte (bf16) aligned bases
    auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    bool v2 = (K % 2 == 0) && (!x || (al(x, 8) && ldx % 2 == 0)) && (!rowvec || al(rowvec, 8)) && (!mean || (al(mean, 8) && al(var, 8))) &&
              (!gate || (al(gate, 8) && ldg % 2 == 0)) && (!y32 || (al(y32, 8) && ldy % 2 == 0)) && (!hi || (al(hi, 4) && al(lo, 4) && ld16 % 2 == 0));
    if (yT32 || hiT) v2 = v2 && (M % 2 == 0) && (!yT32 || (al(yT32, 8) && ldyT % 2 == 0)) && (!hiT || (al(hiT, 4) && al(loT, 4) && ldT % 2 == 0));
    if (v2) xform2_kernel<<<grid, 256, 0, st>>>(P);
    else xform_kernel<<<grid, 256, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t eml_ppo_heads(const float* mu, long long ldmu, const float* logstd, const float* actions, const float* old_neglogp,
                          const float* adv, const float* value, const float* task_value, const float* returns, const float* old_mu,
                          const float* old_sigma, float* dmu, long long lddmu, float* dvalue, float* dtask, float* stats, long long B,
                          int A, float e_clip, float actor_coef, float critic_coef, float tv_coef, float bounds_coef, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    PpoHeadParams P;
    P.mu = mu; P.ldmu = ldmu; P.logstd = logstd; P.actions = actions; P.old_neglogp = old_neglogp; P.adv = adv; P.value = value;
    P.task_value = task_value; P.returns = returns; P.old_mu = old_mu; P.old_sigma = old_sigma; P.dmu = dmu; P.lddmu = lddmu;
    P.dvalue = dvalue; P.dtask = dtask; P.stats = stats; P.B = B; P.A = A; P.e_clip = e_clip; P.actor_coef = actor_coef;
    P.critic_coef = critic_coef; P.tv_coef = tv_coef; P.bounds_coef = bounds_coef; P.inv_B = 1.0f / (float)B;
    ppo_head_kernel<<<(unsigned)((B + 3) / 4), 128, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t eml_disc_heads(const float* logit, float* dlogit, float* stats, long long n_agent, long long n_demo, float coef, cudaStream_t st) {
    const long long n = n_agent + n_demo;
    if (n <= 0) return cudaSuccess;
    disc_head_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logit, dlogit, stats, n_agent, n_demo, coef);
    return cudaGetLastError();
}

cudaError_t eml_amp_dropout_mask(const float* u, float* mask, long long rows, float rate, cudaStream_t st) {
    if (rows <= 0) return cudaSuccess;
    const long long n = rows * EML_AMP_OBS;
    amp_dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(u, mask, rows, rate);
    return cudaGetLastError();
}

cudaError_t eml_rms_update(const float* x, long long ldx, long long M, int K, double* scratch, double* rmean, double* rvar, double* count,
                           float* mean32, float* var32, float* inv32, float eps, cudaStream_t st) {
    if (M <= 0 || K <= 0) return cudaSuccess;
    long long gy = (M + 7) / 8; if (gy > 64) gy = 64;
    dim3 grid((K + 31) / 32, (unsigned)gy);
    column_moments_kernel<<<grid, 256, 0, st>>>(x, ldx, M, K, scratch, scratch + K);
    rms_update_kernel<<<1, 1024, 0, st>>>(scratch, scratch + K, M, K, rmean, rvar, count, mean32, var32, inv32, eps);
    return cudaGetLastError();
}

cudaError_t eml_grad_sumsq(const float* g, long long n, float* state, float* partials, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n, partials);
    sumsq_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, state + 1);
    return cudaGetLastError();
}

cudaError_t eml_adam_clip(float* p, const float* g, float* m, float* v, long long n, float* state, float lr, float beta1, float beta2,
                          float eps, float max_norm, float grad_scale, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    adam_clip_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, state, lr, beta1, beta2, eps, max_norm, grad_scale);
    return cudaGetLastError();
}

cudaError_t eml_dp_reduce_shard(const float* mc_grad, float* shard, long long lo, long long count, float* partials, float* mc_exchange,
                                int rank, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const long long c4 = count / 4;
    long long blocks = (c4 + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    dp_reduce_shard_kernel<<<(unsigned)blocks, 256, 0, st>>>(mc_grad, shard, lo, c4, partials);
    dp_publish_sumsq_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, mc_exchange, rank);
    return cudaGetLastError();
}
cudaError_t eml_dp_adam_shard(float* mc_param, const float* p_local, const float* shard, float* m, float* v, long long lo, long long count,
                              const float* exchange, int world, float* state, float lr, float beta1, float beta2, float eps, float max_norm,
                              float grad_scale, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    const long long c4 = count / 4;
    long long blocks = (c4 + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    dp_adam_shard_kernel<<<(unsigned)blocks, 256, 0, st>>>(mc_param, p_local, shard, m, v, lo, c4, exchange, world, state, lr, beta1, beta2, eps,
                                                          max_norm, grad_scale);
    return cudaGetLastError();
}

cudaError_t eml_adam_begin(float* state, cudaStream_t st) {
    bump_step_kernel<<<1, 1, 0, st>>>(state);
    return cudaGetLastError();
}

cudaError_t eml_sum_parts(const float* parts, int S, long long stride, float* out, long long n, int accumulate, cudaStream_t st) {
    if (n <= 0 || S <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    sum_parts_kernel<<<(unsigned)blocks, 256, 0, st>>>(parts, S, stride, out, n, accumulate);
    return cudaGetLastError();
}

cudaError_t eml_axpy(float* y, const float* x, float a, long long n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    long long blocks = (n + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    axpy_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, x, a, n);
    return cudaGetLastError();
}
