// Shared device helpers for the emloco_b200 kernels (sm_100a).
// Quaternion convention: xyzw, as in the reference (isaacgym/python/isaacgym/torch_utils.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EML_NB 24          // rigid bodies of the SMPL humanoid
#define EML_NJ 23          // spherical (3-hinge) joints
#define EML_ND 69          // degrees of freedom
#define EML_SELF_OBS 368
#define EML_TASK_OBS 1054
#define EML_OBS 1422
#define EML_AMP_STEP 206
#define EML_AMP_STEPS 15
#define EML_AMP_OBS 3090
#define EML_TRAJ_SAMPLES 15
#define EML_NUM_VERTS 101
#define EML_HEAD 13

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// utils/torch_utils.py:15-25 my_quat_rotate: a + b + c
__device__ __forceinline__ f3 quat_rotate(f4 q, f3 v) {
    float s = 2.0f * q.w * q.w - 1.0f;
    f3 qv = mk3(q.x, q.y, q.z);
    f3 a = v * s;
    f3 b = cross3(qv, v) * q.w * 2.0f;
    f3 c = qv * dot3(qv, v) * 2.0f;
    return a + b + c;
}

// isaacgym torch_utils.py:19-41 quat_mul (same operation order)
__device__ __forceinline__ f4 quat_mul(f4 a, f4 b) {
    float ww = (a.z + a.x) * (b.x + b.y);
    float yy = (a.w - a.y) * (b.w + b.z);
    float zz = (a.w + a.y) * (b.w - b.z);
    float xx = ww + yy + zz;
    float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
    float w = qq - ww + (a.z - a.y) * (b.y - b.z);
    float x = qq - xx + (a.x + a.w) * (b.x + b.w);
    float y = qq - yy + (a.w - a.x) * (b.y + b.z);
    float z = qq - zz + (a.z + a.y) * (b.w - b.x);
    return mk4(x, y, z, w);
}

// plain Hamilton product (used by the physics kernel, where no reference formula exists)
__device__ __forceinline__ f4 qmul(f4 a, f4 b) {
    return mk4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
               a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
               a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
               a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}

// rotate by unit quaternion, standard form v + 2w(q x v) + 2 q x (q x v)
__device__ __forceinline__ f3 qrot(f4 q, f3 v) {
    f3 qv = mk3(q.x, q.y, q.z);
    f3 t = cross3(qv, v) * 2.0f;
    return v + t * q.w + cross3(qv, t);
}
__device__ __forceinline__ f4 qconj(f4 q) { return mk4(-q.x, -q.y, -q.z, q.w); }

// utils/torch_utils.py:137-174: heading angle and the +-heading z-rotation quaternion
__device__ __forceinline__ float calc_heading(f4 q) {
    f3 d = quat_rotate(q, mk3(1.f, 0.f, 0.f));
    return atan2f(d.y, d.x);
}
// quat_from_angle_axis(angle, z) incl. the final quat_unit normalisation (torch_utils.py:98-103)
__device__ __forceinline__ f4 quat_from_angle_z(float angle) {
    float th = angle / 2.0f;
    float s = sinf(th), c = cosf(th);
    float n = fmaxf(sqrtf(s * s + c * c), 1e-9f);
    return mk4(0.f, 0.f, s / n, c / n);
}

// utils/torch_utils.py:66-79 quat_to_tan_norm -> 6 floats
__device__ __forceinline__ void quat_to_tan_norm(f4 q, float* o) {
    f3 t = quat_rotate(q, mk3(1.f, 0.f, 0.f));
    f3 n = quat_rotate(q, mk3(0.f, 0.f, 1.f));
    o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = n.x; o[4] = n.y; o[5] = n.z;
}

// utils/torch_utils.py:89-112 exp_map_to_quat
__device__ __forceinline__ f4 exp_map_to_quat(f3 e) {
    float angle = sqrtf(e.x * e.x + e.y * e.y + e.z * e.z);
    f3 axis = mk3(e.x / angle, e.y / angle, e.z / angle);
    angle = atan2f(sinf(angle), cosf(angle));                 // normalize_angle
    bool ok = fabsf(angle) > 1e-5f;
    if (!ok) { angle = 0.f; axis = mk3(0.f, 0.f, 1.f); }
    float th = angle / 2.0f;
    float an = fmaxf(sqrtf(dot3(axis, axis)), 1e-9f);         // normalize(axis)
    float s = sinf(th), c = cosf(th);
    f4 q = mk4(axis.x / an * s, axis.y / an * s, axis.z / an * s, c);
    float n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-9f);
    return mk4(q.x / n, q.y / n, q.z / n, q.w / n);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
