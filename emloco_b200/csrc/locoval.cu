// LocoVal: ValuePoseNet forward / input-gradient backward
// (reference pacer/pacer/learning/value_pose_net.py:10-159) as one fused kernel each:
// heading normalisation (:73-103) -> hide toes/spine (:141-144) -> Linear-ReLU-Linear-ReLU-Linear-sigmoid.
// Also the plausibl value MLP (plausibl/test_value_mlp.py:24-113).
//
// One thread per sample; the 6 174 weights live in shared memory, transposed to [in][out] so every
// weight read is a warp-wide broadcast LDS.128 and the hidden activations stay in registers.
// The reference instead builds a [B,2,2] rotation matrix on the CPU and copies it H2D per call (:85-90).
#include "sim.h"

#include "locoval_common.cuh"

template <int T, bool POSE, bool VEL, bool BWD>
__global__ void __launch_bounds__(LV_THREADS) locoval_kernel(const float* __restrict__ traj, int stride, float* pose_rw,
                                                             const float* __restrict__ vel, const float* __restrict__ weights,
                                                             float* __restrict__ value, const float* __restrict__ gvalue,
                                                             float* __restrict__ gtraj, const float* __restrict__ gpose_out,
                                                             float* __restrict__ gpose_in, long long B, int flags) {
    using D = LvDims<T, POSE, VEL>;
    extern __shared__ __align__(16) float smem[];
    float* s_w1t = smem;
    float* s_b1 = s_w1t + D::IN * D::H1P;
    float* s_w2t = s_b1 + D::H1P;
    float* s_b2 = s_w2t + D::H1 * D::H2P;
    float* s_w3 = s_b2 + D::H2P;
    float* s_b3 = s_w3 + D::H2P;
    lv_stage_weights<T, POSE, VEL>(weights, s_w1t, s_b1, s_w2t, s_b2, s_w3, s_b3);
    __syncthreads();
    const bool hide_toe = flags & 4, hide_spine = flags & 8, normalize = flags & 16, writeback = flags & 32;

    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const float* tr = traj + b * T * stride;
        float c, s, xe; bool near0;
        const float x1 = tr[stride], y1 = tr[stride + 1];
        lv_angle(x1, y1, normalize, c, s, xe, near0);

        float h1[D::H1P];
#pragma unroll
        for (int j = 0; j < D::H1P; ++j) h1[j] = s_b1[j];
        auto acc1 = [&](int k, float x) {
            const float4* wr = reinterpret_cast<const float4*>(s_w1t + k * D::H1P);
#pragma unroll
            for (int j4 = 0; j4 < D::H1P / 4; ++j4) {
                float4 w = wr[j4];
                h1[4 * j4] += w.x * x; h1[4 * j4 + 1] += w.y * x; h1[4 * j4 + 2] += w.z * x; h1[4 * j4 + 3] += w.w * x;
            }
        };
        // ---- layer 1, inputs generated on the fly: row-vector times [[c,-s],[s,c]] (:85-100) ----
#pragma unroll 1
        for (int n = 0; n < T; ++n) {
            float x = tr[n * stride], y = tr[n * stride + 1];
            acc1(2 * n, x * c + y * s);
            acc1(2 * n + 1, y * c - x * s);
        }
        if (POSE) {
            float* pp = pose_rw + b * 72;
#pragma unroll 1
            for (int j = 0; j < 24; ++j) {
                float x = pp[3 * j], y = pp[3 * j + 1], z = pp[3 * j + 2];
                float xr = x * c + y * s, yr = y * c - x * s;
                if ((hide_toe && (j == 4 || j == 8)) || (hide_spine && j >= 9 && j <= 11)) { xr = 0.f; yr = 0.f; z = 0.f; }
                if (!BWD && writeback) { pp[3 * j] = xr; pp[3 * j + 1] = yr; pp[3 * j + 2] = z; }
                acc1(2 * T + 3 * j, xr); acc1(2 * T + 3 * j + 1, yr); acc1(2 * T + 3 * j + 2, z);
            }
        }
        if (VEL) {
            float x = vel[b * 2], y = vel[b * 2 + 1];
            acc1(D::IN - 2, x * c + y * s);
            acc1(D::IN - 1, y * c - x * s);
        }
        // ---- layer 2 / 3 ----
        float h2[D::H2P];
#pragma unroll
        for (int o = 0; o < D::H2P; ++o) h2[o] = s_b2[o];
#pragma unroll
        for (int j = 0; j < D::H1; ++j) {
            float a = fmaxf(h1[j], 0.f);
            h1[j] = a;
            const float4* wr = reinterpret_cast<const float4*>(s_w2t + j * D::H2P);
#pragma unroll
            for (int o4 = 0; o4 < D::H2P / 4; ++o4) {
                float4 w = wr[o4];
                h2[4 * o4] += w.x * a; h2[4 * o4 + 1] += w.y * a; h2[4 * o4 + 2] += w.z * a; h2[4 * o4 + 3] += w.w * a;
            }
        }
        float z = s_b3[0];
#pragma unroll
        for (int o = 0; o < D::H2; ++o) { h2[o] = fmaxf(h2[o], 0.f); z += s_w3[o] * h2[o]; }
        const float v = 1.0f / (1.0f + expf(-z));
        if (!BWD) { value[b] = v; continue; }

        // ---- backward: d value / d traj (autograd of calc_embodied_motion_loss, :151-159) ----
        const float dz = gvalue[b] * v * (1.0f - v);
#pragma unroll
        for (int o = 0; o < D::H2P; ++o) h2[o] = (o < D::H2 && h2[o] > 0.f) ? s_w3[o] * dz : 0.f;   // dh2
#pragma unroll
        for (int j = 0; j < D::H1; ++j) {                                                            // dh1 in place
            const float4* wr = reinterpret_cast<const float4*>(s_w2t + j * D::H2P);
            float d = 0.f;
#pragma unroll
            for (int o4 = 0; o4 < D::H2P / 4; ++o4) {
                float4 w = wr[o4];
                d += w.x * h2[4 * o4] + w.y * h2[4 * o4 + 1] + w.z * h2[4 * o4 + 2] + w.w * h2[4 * o4 + 3];
            }
            h1[j] = h1[j] > 0.f ? d : 0.f;
        }
#pragma unroll
        for (int j = D::H1; j < D::H1P; ++j) h1[j] = 0.f;
        auto dx = [&](int k) {
            const float4* wr = reinterpret_cast<const float4*>(s_w1t + k * D::H1P);
            float d = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < D::H1P / 4; ++j4) {
                float4 w = wr[j4];
                d += w.x * h1[4 * j4] + w.y * h1[4 * j4 + 1] + w.z * h1[4 * j4 + 2] + w.w * h1[4 * j4 + 3];
            }
            return d;
        };
        // d/dtheta of (x',y') = (x c + y s, y c - x s) is (y', -x')
        float dtheta = 0.f;
        float* gt = gtraj + b * T * stride;
#pragma unroll 1
        for (int n = 0; n < T; ++n) {
            float x = tr[n * stride], y = tr[n * stride + 1];
            float xr = x * c + y * s, yr = y * c - x * s;
            float gx = dx(2 * n), gy = dx(2 * n + 1);
            dtheta += gx * yr - gy * xr;
            gt[n * stride] = gx * c - gy * s;
            gt[n * stride + 1] = gx * s + gy * c;
            for (int u = 2; u < stride; ++u) gt[n * stride + u] = 0.f;
        }
        if (POSE) {
            // The reference rotates / zeroes init_pose IN PLACE with autograd history (:97,141-144): the tensor the caller
            // holds afterwards depends on theta and on the incoming pose, and a later call that reuses it (the multi-modal
            // loop of social-transmotion/train_jta.py:294-296) sends gradient back through it.  gpose_out is that gradient
            // w.r.t. the rotated / zeroed pose, gpose_in the gradient w.r.t. the pose as it came in.
            const float* pp = pose_rw + b * 72;
#pragma unroll 1
            for (int j = 0; j < 24; ++j) {
                const bool hidden = (hide_toe && (j == 4 || j == 8)) || (hide_spine && j >= 9 && j <= 11);
                float Gx = 0.f, Gy = 0.f, Gz = 0.f;
                if (!hidden) {
                    Gx = dx(2 * T + 3 * j); Gy = dx(2 * T + 3 * j + 1);
                    if (gpose_in) Gz = dx(2 * T + 3 * j + 2);
                    if (gpose_out) { Gx += gpose_out[b * 72 + 3 * j]; Gy += gpose_out[b * 72 + 3 * j + 1]; Gz += gpose_out[b * 72 + 3 * j + 2]; }
                    float x = pp[3 * j], y = pp[3 * j + 1];
                    float xr = x * c + y * s, yr = y * c - x * s;
                    dtheta += Gx * yr - Gy * xr;
                }
                if (gpose_in) {
                    gpose_in[b * 72 + 3 * j] = Gx * c - Gy * s; gpose_in[b * 72 + 3 * j + 1] = Gx * s + Gy * c; gpose_in[b * 72 + 3 * j + 2] = Gz;
                }
            }
        }
        if (VEL) {
            float x = vel[b * 2], y = vel[b * 2 + 1];
            float xr = x * c + y * s, yr = y * c - x * s;
            dtheta += dx(D::IN - 2) * yr - dx(D::IN - 1) * xr;
        }
        if (normalize) {   // theta = atan2(y1, xe); xe = x1 unless |x1| < 1e-10 (:79-84)
            float r2 = xe * xe + y1 * y1;
            gt[stride + 1] += dtheta * xe / r2;
            if (!near0) gt[stride] += -dtheta * y1 / r2;
        }
    }
}

template <int T, bool POSE, bool VEL, bool BWD>
static cudaError_t lv_launch(const float* traj, int stride, float* pose, const float* vel, const float* w, float* value,
                             const float* gv, float* gt, const float* gpo, float* gpi, long long B, int flags, cudaStream_t st) {
    using D = LvDims<T, POSE, VEL>;
    auto k = locoval_kernel<T, POSE, VEL, BWD>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, D::SMEM);
    if (e != cudaSuccess) return e;
    if (B <= 0) return cudaSuccess;
    long long blocks = (B + LV_THREADS - 1) / LV_THREADS;
    // persistent-ish grid: weights are staged once per CTA, so cap at a few CTAs per SM (148 SMs)
    long long cap = 148LL * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    k<<<grid, LV_THREADS, D::SMEM, st>>>(traj, stride, pose, vel, w, value, gv, gt, gpo, gpi, B, flags);
    return cudaGetLastError();
}

template <bool BWD>
static cudaError_t lv_dispatch(const float* traj, int stride, int T, float* pose, const float* vel, const float* w,
                               float* value, const float* gv, float* gt, const float* gpo, float* gpi, long long B, int flags,
                               cudaStream_t st) {
    bool P = flags & 1, V = flags & 2;
#define LV_CASE(TT, PP, VV) if (T == TT && P == PP && V == VV) return lv_launch<TT, PP, VV, BWD>(traj, stride, pose, vel, w, value, gv, gt, gpo, gpi, B, flags, st)
    LV_CASE(13, true, true); LV_CASE(13, true, false); LV_CASE(13, false, true); LV_CASE(13, false, false);
    LV_CASE(5, true, true); LV_CASE(5, true, false); LV_CASE(5, false, true); LV_CASE(5, false, false);
#undef LV_CASE
    return cudaErrorInvalidValue;
}

cudaError_t eml_locoval_forward_tc(const float* traj, int stride, float* pose, const float* vel, const float* w, float* value,
                                   long long B, int flags, cudaStream_t st);

cudaError_t eml_locoval_forward(const float* traj, int stride, int T, float* pose, const float* vel, const float* w,
                                float* value, long long B, int flags, cudaStream_t st) {
    // large batches of the full variant go to the tensor-core kernel (locoval_tc.cu); flag bit 6 forces the CUDA-core kernel
    if (T == 13 && (flags & 3) == 3 && !(flags & 64) && B >= 1024 && stride <= 3 &&
        ((reinterpret_cast<uintptr_t>(pose) | reinterpret_cast<uintptr_t>(traj) | reinterpret_cast<uintptr_t>(vel)) & 15) == 0)
        return eml_locoval_forward_tc(traj, stride, pose, vel, w, value, B, flags, st);
    return lv_dispatch<false>(traj, stride, T, pose, vel, w, value, nullptr, nullptr, nullptr, nullptr, B, flags, st);
}
cudaError_t eml_locoval_backward(const float* traj, int stride, int T, const float* pose, const float* vel, const float* w,
                                 const float* gv, float* gt, const float* gpose_out, float* gpose_in, long long B, int flags,
                                 cudaStream_t st) {
    return lv_dispatch<true>(traj, stride, T, const_cast<float*>(pose), vel, w, nullptr, gv, gt, gpose_out, gpose_in, B, flags, st);
}

// ---- plausibl/test_value_mlp.py:24-113: Linear(24,12)-ReLU-Linear(12,6)-ReLU-Linear(6,1) ----
__global__ void plausibl_mlp_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, long long B) {
    __shared__ float sw[24 * 12 + 12 + 12 * 6 + 6 + 6 + 1];
    for (int i = threadIdx.x; i < 24 * 12 + 12 + 12 * 6 + 6 + 6 + 1; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const float* w1 = sw; const float* b1 = w1 + 288; const float* w2 = b1 + 12; const float* b2 = w2 + 72;
    const float* w3 = b2 + 6; const float* b3 = w3 + 6;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        float in[24];
#pragma unroll
        for (int k = 0; k < 24; ++k) in[k] = x[b * 24 + k];
        float h1[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            float a = b1[j];
#pragma unroll
            for (int k = 0; k < 24; ++k) a += w1[j * 24 + k] * in[k];
            h1[j] = fmaxf(a, 0.f);
        }
        float z = b3[0];
#pragma unroll
        for (int o = 0; o < 6; ++o) {
            float a = b2[o];
#pragma unroll
            for (int j = 0; j < 12; ++j) a += w2[o * 12 + j] * h1[j];
            z += w3[o] * fmaxf(a, 0.f);
        }
        out[b] = z;
    }
}

cudaError_t eml_plausibl_forward(const float* x, const float* w, float* out, long long B, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    long long blocks = (B + 127) / 128;
    int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
    plausibl_mlp_kernel<<<grid, 128, 0, st>>>(x, w, out, B);
    return cudaGetLastError();
}
