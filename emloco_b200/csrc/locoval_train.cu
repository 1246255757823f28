// LocoVal fine-tuning step inside the rollout: the `if self._do_finetune and len(valid_rewards_idx[0]) > 0` block of
// AMPValueAgent.play_steps (reference pacer/pacer/learning/amp_continuous_value.py:122-146) with the optimiser of
// common_agent.py:94-96 (AdamW lr 1e-3 weight_decay 1e-4, MSELoss(reduction='sum')):
//     valid   = nonzero(game_combined_rewards)
//     pred    = valuenet(waypoint_traj[:, :13], init_pose, init_vel)[valid]
//     target  = (game_combined_rewards[valid] - min_cum_rewards) / (max_cum_rewards - min_cum_rewards)
//     loss    = sum((pred - target)^2);  loss.backward();  AdamW step;  game_combined_rewards = 0
// The reference synchronises with the host for `valid` every control step; here the whole block is two launches driven by the
// device-side flags.  Deterministic: valid envs are compacted in env order, CTA c owns tiles c, c+G, ... and the per-CTA
// gradient partials are reduced in CTA order by the last CTA to finish, which also applies AdamW.
//   kernel 1  lv_valid_kernel    ordered compaction of the envs with game_combined != 0
//   kernel 2  lv_train_kernel    thread = sample: forward + backward to the pre-activations (weights transposed in shared
//                                memory, as locoval.cu); then the CTA turns the tile's (activation, delta) pairs into weight
//                                gradients with every thread owning a strided set of parameters
#include "locoval_common.cuh"

namespace {

constexpr int TR_TILE = 64;        // samples per tile
constexpr int TR_THREADS = 128;
constexpr int TR_GRID = 64;        // CTAs (fixed: the reduction order is part of the result)

struct LvTrainParams {
    const float* traj; int stride; const float* pose; const float* vel;
    float* gc; float* w; float* m; float* v; float* step; float* stats; float* partial;
    int* idx; int* count; unsigned int* ticket;
    float lr, beta1, beta2, eps, wd, r_min, r_max;
    int flags; long long N;
};

__global__ void __launch_bounds__(1024) lv_valid_kernel(const float* __restrict__ gc, long long N, int* __restrict__ idx,
                                                        int* __restrict__ count, unsigned int* __restrict__ ticket) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < N; i0 += 1024) {
        const long long i = i0 + tid;
        const bool valid = i < N && gc[i] != 0.0f;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) { const int c = s_warp[w]; if (w < warp) before += c; total += c; }
        if (valid) idx[s_base + before + __popc(bal & ((1u << lane) - 1u))] = (int)i;
        __syncthreads();
        if (tid == 0) s_base += total;
        __syncthreads();
    }
    if (tid == 0) { *count = s_base; *ticket = 0u; }
}

template <int T, bool POSE, bool VEL>
__global__ void __launch_bounds__(TR_THREADS) lv_train_kernel(LvTrainParams P) {
    using D = LvDims<T, POSE, VEL>;
    constexpr int IN = D::IN, H1 = D::H1, H2 = D::H2, NW = D::NW;
    extern __shared__ __align__(16) float smem[];
    float* s_w1t = smem;                                   // staged weights, transposed (lv_stage_weights)
    float* s_b1 = s_w1t + IN * D::H1P;
    float* s_w2t = s_b1 + D::H1P;
    float* s_b2 = s_w2t + H1 * D::H2P;
    float* s_w3 = s_b2 + D::H2P;
    float* s_b3 = s_w3 + D::H2P;
    float* s_g = s_b3 + 4;                                 // this CTA's gradient accumulator, packed like the weights
    float* s_x = s_g + NW;                                 // tile: inputs [TILE][IN]
    float* s_a1 = s_x + TR_TILE * IN;                      //       relu(h1) [TILE][H1]
    float* s_d1 = s_a1 + TR_TILE * H1;                     //       dL/dh1   [TILE][H1]
    float* s_a2 = s_d1 + TR_TILE * H1;                     //       relu(h2) [TILE][H2]
    float* s_d2 = s_a2 + TR_TILE * H2;                     //       dL/dh2   [TILE][H2]
    float* s_dz = s_d2 + TR_TILE * H2;                     //       dL/dz    [TILE]
    __shared__ float s_red[3][TR_THREADS / 32];
    __shared__ unsigned int s_last;
    __shared__ float s_step;
    const int tid = threadIdx.x;
    const int count = *P.count;
    if (count == 0) return;                                // no valid env: no optimiser step (:124)
    lv_stage_weights<T, POSE, VEL>(P.w, s_w1t, s_b1, s_w2t, s_b2, s_w3, s_b3);
    for (int i = tid; i < NW; i += TR_THREADS) s_g[i] = 0.f;
    __syncthreads();
    const bool hide_toe = P.flags & 4, hide_spine = P.flags & 8, normalize = P.flags & 16;
    const float inv_range = 1.0f / (P.r_max - P.r_min);
    float loss = 0.f, sum_pred = 0.f, sum_gt = 0.f;

    for (int t0 = blockIdx.x * TR_TILE; t0 < count; t0 += gridDim.x * TR_TILE) {
        const int ns = min(TR_TILE, count - t0);
        if (tid < ns) {
            const long long e = P.idx[t0 + tid];
            float* x = s_x + tid * IN;
            const float* tr = P.traj + e * T * P.stride;
            float c, s, xe; bool near0;
            lv_angle(tr[P.stride], tr[P.stride + 1], normalize, c, s, xe, near0);
            for (int n = 0; n < T; ++n) {
                const float px = tr[n * P.stride], py = tr[n * P.stride + 1];
                x[2 * n] = px * c + py * s; x[2 * n + 1] = py * c - px * s;
            }
            if (POSE) {
                const float* pp = P.pose + e * 72;
                for (int j = 0; j < 24; ++j) {
                    float xr = pp[3 * j] * c + pp[3 * j + 1] * s, yr = pp[3 * j + 1] * c - pp[3 * j] * s, z = pp[3 * j + 2];
                    if ((hide_toe && (j == 4 || j == 8)) || (hide_spine && j >= 9 && j <= 11)) { xr = 0.f; yr = 0.f; z = 0.f; }
                    x[2 * T + 3 * j] = xr; x[2 * T + 3 * j + 1] = yr; x[2 * T + 3 * j + 2] = z;
                }
            }
            if (VEL) {
                const float vx = P.vel[e * 2], vy = P.vel[e * 2 + 1];
                x[IN - 2] = vx * c + vy * s; x[IN - 1] = vy * c - vx * s;
            }
            // forward
            float* a1 = s_a1 + tid * H1; float* a2 = s_a2 + tid * H2;
            for (int j = 0; j < H1; ++j) {
                float h = s_b1[j];
                for (int k = 0; k < IN; ++k) h += s_w1t[k * D::H1P + j] * x[k];
                a1[j] = fmaxf(h, 0.f);
            }
            float z = s_b3[0];
            for (int o = 0; o < H2; ++o) {
                float h = s_b2[o];
                for (int j = 0; j < H1; ++j) h += s_w2t[j * D::H2P + o] * a1[j];
                a2[o] = fmaxf(h, 0.f);
                z += s_w3[o] * a2[o];
            }
            const float v = 1.0f / (1.0f + expf(-z));
            const float target = (P.gc[e] - P.r_min) * inv_range;                 // :135
            P.gc[e] = 0.f;                                                        // :145 (the valid envs are the non-zero ones)
            const float diff = v - target;
            loss += diff * diff; sum_pred += v; sum_gt += target;
            // backward to the pre-activations: d(sum sq)/dz = 2 diff v (1 - v)
            const float dz = 2.0f * diff * v * (1.0f - v);
            s_dz[tid] = dz;
            float* d2 = s_d2 + tid * H2; float* d1 = s_d1 + tid * H1;
            for (int o = 0; o < H2; ++o) d2[o] = a2[o] > 0.f ? s_w3[o] * dz : 0.f;
            for (int j = 0; j < H1; ++j) {
                float d = 0.f;
                for (int o = 0; o < H2; ++o) d += s_w2t[j * D::H2P + o] * d2[o];
                d1[j] = a1[j] > 0.f ? d : 0.f;
            }
        }
        __syncthreads();
        // weight gradients of the tile; parameter p of the packed layout [w1 | b1 | w2 | b2 | w3 | b3] is owned by one thread
        for (int p = tid; p < H1 * IN; p += TR_THREADS) {
            const int j = p / IN, k = p - j * IN;
            float a = 0.f;
            for (int s2 = 0; s2 < ns; ++s2) a += s_d1[s2 * H1 + j] * s_x[s2 * IN + k];
            s_g[p] += a;
        }
        for (int p = tid; p < H1; p += TR_THREADS) {
            float a = 0.f;
            for (int s2 = 0; s2 < ns; ++s2) a += s_d1[s2 * H1 + p];
            s_g[H1 * IN + p] += a;
        }
        for (int p = tid; p < H2 * H1; p += TR_THREADS) {
            const int o = p / H1, j = p - o * H1;
            float a = 0.f;
            for (int s2 = 0; s2 < ns; ++s2) a += s_d2[s2 * H2 + o] * s_a1[s2 * H1 + j];
            s_g[H1 * IN + H1 + p] += a;
        }
        for (int p = tid; p < H2; p += TR_THREADS) {
            float a = 0.f, bq = 0.f;
            for (int s2 = 0; s2 < ns; ++s2) { a += s_d2[s2 * H2 + p]; bq += s_dz[s2] * s_a2[s2 * H2 + p]; }
            s_g[H1 * IN + H1 + H2 * H1 + p] += a;                                  // b2
            s_g[H1 * IN + H1 + H2 * H1 + H2 + p] += bq;                            // w3
        }
        if (tid == 0) {
            float a = 0.f;
            for (int s2 = 0; s2 < ns; ++s2) a += s_dz[s2];
            s_g[NW - 1] += a;                                                     // b3
        }
        __syncthreads();
    }
    // per-CTA partials (gradient + loss / prediction / target sums), then the last CTA reduces in CTA order and steps AdamW
    float* part = P.partial + (size_t)blockIdx.x * (NW + 4);
    for (int i = tid; i < NW; i += TR_THREADS) part[i] = s_g[i];
    {
        float r[3] = {loss, sum_pred, sum_gt};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v = r[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) s_red[q][tid >> 5] = v;
        }
        __syncthreads();
        if (tid < 3) { float v = 0.f; for (int w = 0; w < TR_THREADS / 32; ++w) v += s_red[tid][w]; part[NW + tid] = v; }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(P.ticket, 1u);
    __syncthreads();
    if (s_last != gridDim.x - 1) return;
    __threadfence();
    if (tid == 0) s_step = *P.step + 1.0f;                                        // one reader, then a barrier: no thread may see
    __syncthreads();                                                              // the count tid 4 stores below
    const float t = s_step;                                                       // optimiser step count (torch AdamW)
    const float bc1 = 1.0f - powf(P.beta1, t), bc2 = 1.0f - powf(P.beta2, t);
    const float step_size = P.lr / bc1, rs_bc2 = rsqrtf(bc2);
    for (int i = tid; i < NW; i += TR_THREADS) {
        float g = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) g += __ldcg(P.partial + (size_t)c * (NW + 4) + i);
        float w = P.w[i] * (1.0f - P.lr * P.wd);                                  // decoupled weight decay
        const float m = P.beta1 * P.m[i] + (1.0f - P.beta1) * g;
        const float v = P.beta2 * P.v[i] + (1.0f - P.beta2) * g * g;
        w -= step_size * m / (sqrtf(v) * rs_bc2 + P.eps);
        P.w[i] = w; P.m[i] = m; P.v[i] = v;
    }
    if (tid < 3) {
        float a = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) a += __ldcg(P.partial + (size_t)c * (NW + 4) + NW + tid);
        P.stats[tid] += a;                                                        // vnet_loss, sum(vnet_pred), sum(vnet_gt) (:141-144)
    }
    if (tid == 3) P.stats[3] += (float)count;
    if (tid == 4) *P.step = t;
}

template <int T, bool POSE, bool VEL>
cudaError_t train_launch(const LvTrainParams& P, cudaStream_t st) {
    using D = LvDims<T, POSE, VEL>;
    const size_t smem = D::SMEM + (size_t)(D::NW + TR_TILE * (D::IN + 2 * D::H1 + 2 * D::H2 + 1)) * 4;
    auto k = lv_train_kernel<T, POSE, VEL>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    lv_valid_kernel<<<1, 1024, 0, st>>>(P.gc, P.N, P.idx, P.count, P.ticket);
    k<<<TR_GRID, TR_THREADS, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

size_t eml_locoval_train_workspace_bytes(long long N) {
    // idx [N] | count, ticket | partial [GRID][NW_max + 4]
    return ((size_t)N + 4) * 4 + (size_t)TR_GRID * (6174 + 4) * 4;
}

cudaError_t eml_locoval_train_step(const float* traj, int stride, int T, const float* pose, const float* vel, float* gc, float* w,
                                   float* m, float* v, float* step, float* stats, void* workspace, long long N, float lr, float beta1,
                                   float beta2, float eps, float wd, float r_min, float r_max, int flags, cudaStream_t st) {
    LvTrainParams P;
    P.traj = traj; P.stride = stride; P.pose = pose; P.vel = vel; P.gc = gc; P.w = w; P.m = m; P.v = v; P.step = step; P.stats = stats;
    P.idx = reinterpret_cast<int*>(workspace); P.count = P.idx + N; P.ticket = reinterpret_cast<unsigned int*>(P.idx + N + 1);
    P.partial = reinterpret_cast<float*>(P.idx + N + 4);
    P.lr = lr; P.beta1 = beta1; P.beta2 = beta2; P.eps = eps; P.wd = wd; P.r_min = r_min; P.r_max = r_max; P.flags = flags; P.N = N;
    const bool PO = flags & 1, VE = flags & 2;
#define LV_CASE(TT, PP, VV) if (T == TT && PO == PP && VE == VV) return train_launch<TT, PP, VV>(P, st)
    LV_CASE(13, true, true); LV_CASE(13, true, false); LV_CASE(13, false, true); LV_CASE(13, false, false);
    LV_CASE(5, true, true); LV_CASE(5, true, false); LV_CASE(5, false, true); LV_CASE(5, false, false);
#undef LV_CASE
    return cudaErrorInvalidValue;
}
