// Shared by locoval.cu (forward / input-gradient) and locoval_train.cu (weight-gradient + AdamW): layer sizes of
// ValuePoseNet (reference pacer/pacer/learning/value_pose_net.py:36-53), weight staging, heading angle.
#pragma once
#include "sim.h"

#define LV_THREADS 128

template <int T, bool POSE, bool VEL>
struct LvDims {
    static constexpr int IN = 2 * T + (POSE ? 72 : 0) + (VEL ? 2 : 0);
    static constexpr int H1 = IN / 2 - 1;       // value_pose_net.py:52
    static constexpr int H2 = H1 / 2;           // :53
    static constexpr int H1P = (H1 + 3) & ~3;   // padded to float4
    static constexpr int H2P = (H2 + 3) & ~3;
    static constexpr int NW = IN * H1 + H1 + H1 * H2 + H2 + H2 + 1;
    static constexpr int SMEM = (IN * H1P + H1P + H1 * H2P + H2P + H2P + 4) * 4;
};

template <int T, bool POSE, bool VEL>
__device__ __forceinline__ void lv_stage_weights(const float* __restrict__ w, float* s_w1t, float* s_b1, float* s_w2t,
                                                 float* s_b2, float* s_w3, float* s_b3) {
    using D = LvDims<T, POSE, VEL>;
    const float* w1 = w; const float* b1 = w1 + D::IN * D::H1;
    const float* w2 = b1 + D::H1; const float* b2 = w2 + D::H1 * D::H2;
    const float* w3 = b2 + D::H2; const float* b3 = w3 + D::H2;
    for (int i = threadIdx.x; i < D::IN * D::H1P; i += blockDim.x) {
        int k = i / D::H1P, j = i % D::H1P;
        s_w1t[i] = j < D::H1 ? w1[j * D::IN + k] : 0.f;
    }
    for (int i = threadIdx.x; i < D::H1 * D::H2P; i += blockDim.x) {
        int j = i / D::H2P, o = i % D::H2P;
        s_w2t[i] = o < D::H2 ? w2[o * D::H1 + j] : 0.f;
    }
    for (int i = threadIdx.x; i < D::H1P; i += blockDim.x) s_b1[i] = i < D::H1 ? b1[i] : 0.f;
    for (int i = threadIdx.x; i < D::H2P; i += blockDim.x) { s_b2[i] = i < D::H2 ? b2[i] : 0.f; s_w3[i] = i < D::H2 ? w3[i] : 0.f; }
    if (threadIdx.x == 0) s_b3[0] = b3[0];
}

// rotation angle of _rotate_normalization (:76-84)
__device__ __forceinline__ void lv_angle(float x1, float y1, bool normalize, float& c, float& s, float& xe, bool& near0) {
    near0 = fabsf(x1) < 1e-10f;
    xe = near0 ? 1e-10f : x1;
    if (normalize) { float a = atan2f(y1, xe); sincosf(a, &s, &c); } else { c = 1.f; s = 0.f; }
}

