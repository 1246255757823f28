// Mocap reset state and AMP demo observations on the device (second half).
// Reference: MotionLibSMPL.get_motion_state_smpl (pacer/pacer/utils/motion_lib_smpl.py:485-563: frame pair + blend :596-606,
// lerp of positions / velocities, slerp of rotations utils/torch_utils.py:114-136, local rotations -> exp-map DOFs :611-614 ->
// utils/torch_utils.py:27-65) and HumanoidAMP.build_amp_obs_demo -> build_amp_observations_smpl
// (env/tasks/humanoid_amp.py:186-211,917-971): `steps` observations per sample at times t0 - k dt, newest first.
// The reference runs these as ~60 eager torch launches over gathered [n*15, 24, ...] temporaries per demo batch; here one warp
// handles one (sample[, step]) with lane = body and writes the result rows directly.
#include "sim.h"

struct EmlMotionLibDev {
    const float *gts, *grs, *lrs, *gvs, *gavs, *dvs;      // per frame: [F,24,3] [F,24,4] [F,24,4] [F,24,3] [F,24,3] [F,23,3]
    const float *length, *dt, *bodies;                    // per motion: seconds, frame period, shape parameters [17]
    const int *num_frames, *start;                        // per motion: frame count, index of its first frame
    int num_motions;
};

namespace {

__device__ __forceinline__ f4 slerp_ref(f4 q0, f4 q1, float t) {        // utils/torch_utils.py:114-136
    float c = q0.x * q1.x + q0.y * q1.y + q0.z * q1.z + q0.w * q1.w;
    if (c < 0.f) { q1 = mk4(-q1.x, -q1.y, -q1.z, -q1.w); }
    c = fabsf(c);
    const float half = acosf(fminf(c, 1.0f));
    const float s = sqrtf(fmaxf(1.0f - c * c, 0.f));
    const float ra = sinf((1.0f - t) * half) / s, rb = sinf(t * half) / s;
    f4 q = mk4(ra * q0.x + rb * q1.x, ra * q0.y + rb * q1.y, ra * q0.z + rb * q1.z, ra * q0.w + rb * q1.w);
    if (fabsf(s) < 0.001f) q = mk4(0.5f * q0.x + 0.5f * q1.x, 0.5f * q0.y + 0.5f * q1.y, 0.5f * q0.z + 0.5f * q1.z, 0.5f * q0.w + 0.5f * q1.w);
    if (c >= 1.0f) q = q0;
    return q;
}

__device__ __forceinline__ f3 quat_to_exp_map(f4 q) {                   // utils/torch_utils.py:27-65
    const float sin_t = sqrtf(fmaxf(1.0f - q.w * q.w, 0.f));
    float ang = 2.0f * acosf(fminf(fmaxf(q.w, -1.0f), 1.0f));
    ang = atan2f(sinf(ang), cosf(ang));                                 // normalize_angle
    if (!(fabsf(sin_t) > 1e-5f)) return mk3(0.f, 0.f, 0.f);            // angle 0 (default axis z)
    return mk3(ang * q.x / sin_t, ang * q.y / sin_t, ang * q.z / sin_t);
}

struct Blend { long long f0, f1; float b; };
__device__ __forceinline__ Blend frame_blend(const EmlMotionLibDev& L, int id, float time) {      // :596-606
    const float len = L.length[id], dt = L.dt[id];
    const int nf = L.num_frames[id];
    const float phase = fminf(fmaxf(time / len, 0.f), 1.0f);
    if (time < 0.f) time = 0.f;
    const int i0 = (int)(phase * (float)(nf - 1));
    const int i1 = min(i0 + 1, nf - 1);
    Blend r; r.f0 = (long long)L.start[id] + i0; r.f1 = (long long)L.start[id] + i1; r.b = (time - (float)i0 * dt) / dt;
    return r;
}
__device__ __forceinline__ f3 lerp3(const float* a, long long f0, long long f1, int per, int j, float b) {
    const float* p0 = a + (f0 * per + j) * 3; const float* p1 = a + (f1 * per + j) * 3;
    return mk3((1.0f - b) * p0[0] + b * p1[0], (1.0f - b) * p0[1] + b * p1[1], (1.0f - b) * p0[2] + b * p1[2]);
}
__device__ __forceinline__ f4 ld_q(const float* a, long long f, int j) { const float* p = a + (f * EML_NB + j) * 4; return mk4(p[0], p[1], p[2], p[3]); }

// warp = sample, lane = body
__global__ void __launch_bounds__(128) motion_state_kernel(EmlMotionLibDev L, const int* __restrict__ ids, const float* __restrict__ times,
                                                           long long n, float* __restrict__ root, float* __restrict__ dof,
                                                           float* __restrict__ key_pos, float* __restrict__ rb) {
    const long long s = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= n || lane >= EML_NB) return;
    const int id = ids[s];
    const Blend B = frame_blend(L, id, times[s]);
    const f3 pos = lerp3(L.gts, B.f0, B.f1, EML_NB, lane, B.b);
    const f3 vel = lerp3(L.gvs, B.f0, B.f1, EML_NB, lane, B.b);
    const f3 ang = lerp3(L.gavs, B.f0, B.f1, EML_NB, lane, B.b);
    const f4 rot = slerp_ref(ld_q(L.grs, B.f0, lane), ld_q(L.grs, B.f1, lane), B.b);
    if (rb) {
        float* o = rb + (s * EML_NB + lane) * 13;
        o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = rot.x; o[4] = rot.y; o[5] = rot.z; o[6] = rot.w;
        o[7] = vel.x; o[8] = vel.y; o[9] = vel.z; o[10] = ang.x; o[11] = ang.y; o[12] = ang.z;
    }
    if (lane == 0 && root) {
        float* o = root + s * 13;
        o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = rot.x; o[4] = rot.y; o[5] = rot.z; o[6] = rot.w;
        o[7] = vel.x; o[8] = vel.y; o[9] = vel.z; o[10] = ang.x; o[11] = ang.y; o[12] = ang.z;
    }
    if (lane >= 1 && dof) {
        const int j = lane - 1;
        const f3 e = quat_to_exp_map(slerp_ref(ld_q(L.lrs, B.f0, lane), ld_q(L.lrs, B.f1, lane), B.b));
        const f3 dv = lerp3(L.dvs, B.f0, B.f1, EML_NJ, j, B.b);
        float* o = dof + (s * EML_ND + 3 * j) * 2;
        o[0] = e.x; o[1] = dv.x; o[2] = e.y; o[3] = dv.y; o[4] = e.z; o[5] = dv.z;
    }
    if (key_pos) {
        const int k = lane == 7 ? 0 : lane == 3 ? 1 : lane == 22 ? 2 : lane == 17 ? 3 : -1;        // R_Ankle, L_Ankle, R_Wrist, L_Wrist
        if (k >= 0) { float* o = key_pos + (s * 4 + k) * 3; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; }
    }
}

// warp = (sample, history step), lane = body
__global__ void __launch_bounds__(128) amp_obs_demo_kernel(EmlMotionLibDev L, const int* __restrict__ ids, const float* __restrict__ times0,
                                                           long long n, int steps, float dt, float* __restrict__ out) {
    const long long w = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= n * steps) return;
    const long long s = w / steps;
    const int step = (int)(w - s * steps);
    const int id = ids[s];
    const float time = times0[s] + (-dt * (float)step);
    const Blend B = frame_blend(L, id, time);
    float* o = out + w * EML_AMP_STEP;
    const int b = lane < EML_NB ? lane : 0;
    const f3 pos = lerp3(L.gts, B.f0, B.f1, EML_NB, b, B.b);
    // root quantities, computed by every lane from body 0 (cheaper than shuffling 13 values around)
    const f3 rpos = lerp3(L.gts, B.f0, B.f1, EML_NB, 0, B.b);
    const f4 rrot = slerp_ref(ld_q(L.grs, B.f0, 0), ld_q(L.grs, B.f1, 0), B.b);
    const f4 hinv = quat_from_angle_z(-calc_heading(rrot));
    if (lane == 0) {
        float t6[6];
        quat_to_tan_norm(quat_mul(hinv, rrot), t6);
        const f3 lv = quat_rotate(hinv, lerp3(L.gvs, B.f0, B.f1, EML_NB, 0, B.b));
        const f3 la = quat_rotate(hinv, lerp3(L.gavs, B.f0, B.f1, EML_NB, 0, B.b));
#pragma unroll
        for (int q = 0; q < 6; ++q) o[q] = t6[q];
        o[6] = lv.x; o[7] = lv.y; o[8] = lv.z; o[9] = la.x; o[10] = la.y; o[11] = la.z;
    }
    if (lane >= 1 && lane < EML_NB) {
        const int j = lane - 1;
        if (j != 3 && j != 7 && j != 17 && j != 22) {                    // dof_subset: hands and toes dropped (humanoid.py:290-326)
            const int r = j - (j > 3) - (j > 7) - (j > 17) - (j > 22);
            const f3 e = quat_to_exp_map(slerp_ref(ld_q(L.lrs, B.f0, lane), ld_q(L.lrs, B.f1, lane), B.b));
            float t6[6];
            quat_to_tan_norm(exp_map_to_quat(e), t6);                     // dof_to_obs_smpl (humanoid.py:1327-1338)
#pragma unroll
            for (int q = 0; q < 6; ++q) o[12 + 6 * r + q] = t6[q];
            const f3 dv = lerp3(L.dvs, B.f0, B.f1, EML_NJ, j, B.b);
            o[126 + 3 * r] = dv.x; o[127 + 3 * r] = dv.y; o[128 + 3 * r] = dv.z;
        }
        const int k = lane == 7 ? 0 : lane == 3 ? 1 : lane == 22 ? 2 : lane == 17 ? 3 : -1;
        if (k >= 0) { const f3 l = quat_rotate(hinv, pos - rpos); o[183 + 3 * k] = l.x; o[184 + 3 * k] = l.y; o[185 + 3 * k] = l.z; }
    }
    if (lane < 11) o[195 + lane] = L.bodies[(long long)id * 17 + lane];  // smpl_params[:, :-6]
}

}  // namespace

cudaError_t eml_motion_state(const EmlMotionLibDev& L, const int* ids, const float* times, long long n, float* root, float* dof,
                             float* key_pos, float* rb, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    motion_state_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(L, ids, times, n, root, dof, key_pos, rb);
    return cudaGetLastError();
}
cudaError_t eml_amp_obs_demo(const EmlMotionLibDev& L, const int* ids, const float* times0, long long n, int steps, float dt, float* out,
                             cudaStream_t st) {
    if (n <= 0 || steps <= 0) return cudaSuccess;
    amp_obs_demo_kernel<<<(unsigned)((n * steps + 3) / 4), 128, 0, st>>>(L, ids, times0, n, steps, dt, out);
    return cudaGetLastError();
}
