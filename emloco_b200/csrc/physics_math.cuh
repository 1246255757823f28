// Small fixed-size linear algebra shared by the physics kernels (physics.cu: warp-per-env; physics_soa.cu: lane-per-env).
#pragma once
#include "common.cuh"

#define FULL 0xffffffffu

struct S3 { float xx, xy, xz, yy, yz, zz; };          // symmetric 3x3
struct M3 { float a[9]; };                             // row-major 3x3

__device__ __forceinline__ f3 shfl3(f3 v, int src) {
    return mk3(__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src), __shfl_sync(FULL, v.z, src));
}
__device__ __forceinline__ f4 shfl4(f4 v, int src) {
    return mk4(__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src), __shfl_sync(FULL, v.z, src), __shfl_sync(FULL, v.w, src));
}
__device__ __forceinline__ M3 quat_to_mat(f4 q) {
    M3 R;
    float x = q.x, y = q.y, z = q.z, w = q.w;
    R.a[0] = 1.f - 2.f * (y * y + z * z); R.a[1] = 2.f * (x * y - z * w); R.a[2] = 2.f * (x * z + y * w);
    R.a[3] = 2.f * (x * y + z * w); R.a[4] = 1.f - 2.f * (x * x + z * z); R.a[5] = 2.f * (y * z - x * w);
    R.a[6] = 2.f * (x * z - y * w); R.a[7] = 2.f * (y * z + x * w); R.a[8] = 1.f - 2.f * (x * x + y * y);
    return R;
}
__device__ __forceinline__ f3 mv(const M3& R, f3 v) {
    return mk3(R.a[0] * v.x + R.a[1] * v.y + R.a[2] * v.z, R.a[3] * v.x + R.a[4] * v.y + R.a[5] * v.z,
               R.a[6] * v.x + R.a[7] * v.y + R.a[8] * v.z);
}
__device__ __forceinline__ f3 mtv(const M3& R, f3 v) {
    return mk3(R.a[0] * v.x + R.a[3] * v.y + R.a[6] * v.z, R.a[1] * v.x + R.a[4] * v.y + R.a[7] * v.z,
               R.a[2] * v.x + R.a[5] * v.y + R.a[8] * v.z);
}
__device__ __forceinline__ f3 sv(const S3& s, f3 v) {
    return mk3(s.xx * v.x + s.xy * v.y + s.xz * v.z, s.xy * v.x + s.yy * v.y + s.yz * v.z,
               s.xz * v.x + s.yz * v.y + s.zz * v.z);
}
__device__ __forceinline__ f3 row(const M3& m, int r) { return mk3(m.a[3 * r], m.a[3 * r + 1], m.a[3 * r + 2]); }
__device__ __forceinline__ f3 col(const M3& m, int c) { return mk3(m.a[c], m.a[3 + c], m.a[6 + c]); }
__device__ __forceinline__ void setrow(M3& m, int r, f3 v) { m.a[3 * r] = v.x; m.a[3 * r + 1] = v.y; m.a[3 * r + 2] = v.z; }
__device__ __forceinline__ void setcol(M3& m, int c, f3 v) { m.a[c] = v.x; m.a[3 + c] = v.y; m.a[6 + c] = v.z; }
__device__ __forceinline__ f3 srow(const S3& s, int r) {
    return r == 0 ? mk3(s.xx, s.xy, s.xz) : (r == 1 ? mk3(s.xy, s.yy, s.yz) : mk3(s.xz, s.yz, s.zz));
}
__device__ __forceinline__ S3 inv_s3(const S3& s) {
    float c00 = s.yy * s.zz - s.yz * s.yz, c01 = s.xz * s.yz - s.xy * s.zz, c02 = s.xy * s.yz - s.xz * s.yy;
    float det = s.xx * c00 + s.xy * c01 + s.xz * c02;
    float id = 1.0f / det;
    S3 r;
    r.xx = c00 * id; r.xy = c01 * id; r.xz = c02 * id;
    r.yy = (s.xx * s.zz - s.xz * s.xz) * id; r.yz = (s.xy * s.xz - s.xx * s.yz) * id;
    r.zz = (s.xx * s.yy - s.xy * s.xy) * id;
    return r;
}
__device__ __forceinline__ f4 exp_quat(f3 v) {
    float a2 = dot3(v, v), a = sqrtf(a2);
    float s = a > 1e-4f ? sinf(0.5f * a) / a : 0.5f - a2 / 48.0f;
    return mk4(v.x * s, v.y * s, v.z * s, cosf(0.5f * a));
}
__device__ __forceinline__ f3 log_quat(f4 q) {
    if (q.w < 0.f) q = mk4(-q.x, -q.y, -q.z, -q.w);
    float s = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    float k = s > 1e-6f ? 2.0f * atan2f(s, q.w) / s : 2.0f;
    return mk3(q.x * k, q.y * k, q.z * k);
}
__device__ __forceinline__ f4 qnormalize(f4 q) {
    float n = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return mk4(q.x * n, q.y * n, q.z * n, q.w * n);
}


struct PhysParams {
    const EmlModelDev* model;           // device copy of the model
    const float* env_model;             // [EM_FLOATS][N] per-env body models or NULL
    const float* actions;               // [N,69] or NULL (then pd_target is used as is)
    float* pd_target;                   // [N,69]
    float* actions_copy;                // [N,69] or NULL
    float* root; float* dof; float* jq; float* rb; float* contact; float* dof_force;
    const int16_t* height; int hf_rows, hf_cols; float hf_max;
    const int32_t* env_ids;             // FK-only mode: optional env list
    const int64_t* reset_mask;          // FK-only mode: only envs whose flag is set (device-side reset of done envs)
    const float* init_root;             // FK-only mode: take the state from these buffers instead of root/dof
    const float* init_dof;
    int N; int n_sub; float dt;
    float gz, kn, cn, ct, mu, max_w, max_effort, max_turn;
    int fk_only;
    int epb;                            // lane-per-env kernel: envs per CTA (<= 32)
};

// Body-model access: the shared EmlModelDev, or (ENVM) the env's own arrays.  Compile-time switch: the shared-model kernels are
// unchanged instruction for instruction.
template <bool ENVM> struct ModelView {
    const EmlModelDev& Mo; const float* em; size_t N; size_t env;
    __device__ __forceinline__ float g(int off, int i) const { return __ldg(em + (size_t)(off + i) * N + env); }
    __device__ __forceinline__ f3 offset(int b) const { return ENVM ? mk3(g(EM_OFFSET, 3 * b), g(EM_OFFSET, 3 * b + 1), g(EM_OFFSET, 3 * b + 2)) : mk3(Mo.offset[b][0], Mo.offset[b][1], Mo.offset[b][2]); }
    __device__ __forceinline__ f3 com(int b) const { return ENVM ? mk3(g(EM_COM, 3 * b), g(EM_COM, 3 * b + 1), g(EM_COM, 3 * b + 2)) : mk3(Mo.com[b][0], Mo.com[b][1], Mo.com[b][2]); }
    __device__ __forceinline__ f3 geom_a(int b) const { return ENVM ? mk3(g(EM_GA, 3 * b), g(EM_GA, 3 * b + 1), g(EM_GA, 3 * b + 2)) : mk3(Mo.geom_a[b][0], Mo.geom_a[b][1], Mo.geom_a[b][2]); }
    __device__ __forceinline__ f3 geom_b(int b) const { return ENVM ? mk3(g(EM_GB, 3 * b), g(EM_GB, 3 * b + 1), g(EM_GB, 3 * b + 2)) : mk3(Mo.geom_b[b][0], Mo.geom_b[b][1], Mo.geom_b[b][2]); }
    __device__ __forceinline__ float mass(int b) const { return ENVM ? g(EM_MASS, b) : Mo.mass[b]; }
    __device__ __forceinline__ float inertia(int b, int k) const { return ENVM ? g(EM_INERTIA, 6 * b + k) : Mo.inertia[b][k]; }
    __device__ __forceinline__ float kp(int b) const { return ENVM ? g(EM_KP, b) : Mo.kp[b]; }
    __device__ __forceinline__ float kd(int b) const { return ENVM ? g(EM_KD, b) : Mo.kd[b]; }
    __device__ __forceinline__ float arm(int b) const { return ENVM ? g(EM_ARM, b) : Mo.arm[b]; }
    __device__ __forceinline__ float geom_r(int b) const { return ENVM ? g(EM_GR, b) : Mo.geom_r[b]; }
    __device__ __forceinline__ float geom_bound(int b) const { return ENVM ? g(EM_BOUND, b) : Mo.geom_bound[b]; }
};

// trunc(x / 0.1f) without the IEEE division subroutine: same integer part as the fp32 quotient for every float in [0, 2^24)
// (see div_by_tenth in poststep.cu and scripts/cu/div_by_tenth_check.cu); evaluated for every contact point of every sub-step
__device__ __forceinline__ int cell_index(float x, int n) {
    const float q0 = __fmul_rn(x, 10.0f);
    const float q = __fmaf_rn(__fmaf_rn(-q0, 0.1f, x), 10.0f, q0);
    const int i = __float2int_rz(fminf(fmaxf(q, -1.0f), (float)n));
    return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}
__device__ __forceinline__ float ground_height(const PhysParams& P, float x, float y) {
    if (!P.height) return 0.f;
    const int px = cell_index(x, P.hf_rows), py = cell_index(y, P.hf_cols);
    return (float)__ldg(P.height + px * P.hf_cols + py) * 0.005f;
}

