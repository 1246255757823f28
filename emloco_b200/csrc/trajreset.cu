// Device-side trajectory regeneration for the envs that reset:
//   TrajGenerator.reset            pacer/pacer/env/util/traj_generator.py:60-237
//   HumanoidPedestrianTerrain._reset_task   env/tasks/humanoid_pedestrain_terrain.py:493-516
//   get_waypoint_traj / get_init_pose / get_init_vel   env/tasks/vec_task_wrappers.py:47-63
// The reference does this on the host for `done_indices` (python RNG, python loops over a pickle of real trajectories, a
// device round trip per reset).  Here one warp regenerates one env, driven by the reset flags already on the device; the
// random draws are either an explicit [N, >=405] uniform buffer (parity tests replay the reference's draws) or a
// Philox4x32-10 counter stream keyed by (seed, env, per-env reset count).
#include "sim.h"

namespace {

constexpr int TR_WARPS = 4;
constexpr int NV = EML_NUM_VERTS;      // 101
constexpr int NSEG = NV - 1;           // 100
constexpr int NCHUNK = (NSEG + 31) / 32;

struct TrajParams {
    emloco_traj_cfg c;
    const int64_t* reset; const int64_t* progress;
    const float* root_state; const float* rb_state;
    float* verts; uint32_t* epoch;
    int64_t* reset_w; int64_t* terminate_w;
    int N; int clear; float dt; float traj_dur; float control_dt; float sample_dt;
};

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}

struct Draws {
    const float* row; uint32_t k0, k1, env, epoch;
    __device__ float operator()(int col) const {
        if (row) return row[col];
        uint32_t c[4] = {(uint32_t)(col >> 2), env, epoch, 0x74726a31u};
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) { philox_round(c, a, b); a += 0x9E3779B9u; b += 0xBB67AE85u; }
        return (float)(c[col & 3] >> 8) * (1.0f / 16777216.0f);
    }
};

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}

__device__ __forceinline__ void lerp_pos(const float (*v)[3], float t, float traj_dur, float* out) {
    const float phase = fminf(fmaxf(__fdiv_rn(t, traj_dur), 0.0f), 1.0f);        // calc_pos :278-296
    const float seg = phase * (float)NSEG;
    const int i0 = (int)floorf(seg), i1 = (int)ceilf(seg);
    const float w = seg - (float)i0;
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = (1.0f - w) * v[i0][k] + w * v[i1][k];
}

__global__ void __launch_bounds__(TR_WARPS * 32) traj_reset_kernel(TrajParams P) {
    __shared__ float s_v[TR_WARPS][NV][3];
    __shared__ float s_a[TR_WARPS][NSEG];      // dspeed -> speed -> segment length
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int env = blockIdx.x * TR_WARPS + w;
    if (env >= P.N || P.reset[env] == 0) return;
    const emloco_traj_cfg& c = P.c;
    float (*v)[3] = s_v[w];
    float* a = s_a[w];
    uint32_t ep = 0;
    if (!c.uniform) { ep = P.epoch[env]; __syncwarp(); if (lane == 0) P.epoch[env] = ep + 1; }
    Draws U{c.uniform ? c.uniform + (size_t)env * c.ld_uniform : nullptr, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), (uint32_t)env, ep};
    const float* root = P.root_state + (size_t)env * 13;
    const float px = root[0], py = root[1], vx = root[7], vy = root[8], vz = root[9];
    const float PI = 3.14159265358979323846f;
    float* gv = P.verts + (size_t)env * NV * 3;

    // --- speed random walk (:74-81): a clipped recurrence, walked by one lane
    for (int i = lane; i < NSEG; i += 32)
        a[i] = i == 0 ? (c.speed_max - c.speed_min) * U(401) + c.speed_min : (2.0f * U(301 + i) - 1.0f) * c.accel_max * P.dt;
    __syncwarp();
    if (lane == 0) {
        float s = a[0];
        for (int i = 1; i < NSEG; ++i) { s = fminf(fmaxf(s + a[i], c.speed_min), c.speed_max); a[i] = s; }
    }
    __syncwarp();
    const float root_speed = sqrtf(vx * vx + vy * vy);
    float ratio = 1.0f;
    if (c.flags & EMLOCO_TRAJ_ADJUST_ROOT_VEL) { float s0 = a[0]; if (c.flags & EMLOCO_TRAJ_SLOW) s0 = s0 / 4.0f; ratio = __fdiv_rn(root_speed, s0); }
    // --- headings (:63-71), cumulative (:107), segment vectors and their running sum (:112-118)
    float carry_t = 0.0f, carry_x = 0.0f, carry_y = 0.0f;
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int i = ch * 32 + lane;
        float dth = 0.0f, sp = 0.0f;
        if (i < NSEG) {
            dth = (2.0f * U(i) - 1.0f) * c.dtheta_max * P.dt;
            if (U(200 + i) < c.sharp_turn_prob) dth = PI * (2.0f * U(100 + i) - 1.0f);
            if (i == 0) dth = PI * (2.0f * U(300) - 1.0f);
            sp = a[i];
            if (c.flags & EMLOCO_TRAJ_SLOW) sp = sp / 4.0f;
            if (c.flags & EMLOCO_TRAJ_ADJUST_ROOT_VEL) sp = fminf(fmaxf(ratio * sp, c.speed_min), c.speed_max);
        }
        const float th = warp_incl_scan(dth, lane) + carry_t;
        carry_t = __shfl_sync(0xffffffffu, th, 31);
        float sn, cs; sincosf(th, &sn, &cs);
        const float len = sp * P.dt;
        float dx = cs * len, dy = -sn * len;
        if (i == 0) { dx += px; dy += py; }
        const float cx = warp_incl_scan(dx, lane) + carry_x, cy = warp_incl_scan(dy, lane) + carry_y;
        carry_x = __shfl_sync(0xffffffffu, cx, 31); carry_y = __shfl_sync(0xffffffffu, cy, 31);
        if (i < NSEG) { v[i + 1][0] = cx; v[i + 1][1] = cy; v[i + 1][2] = 0.0f; }
    }
    if (lane == 0) { v[0][0] = px; v[0][1] = py; v[0][2] = gv[2]; }    // vertex 0 keeps its z (:120 writes xy only)
    __syncwarp();
    // --- real-world trajectory from the pool (:116-160)
    if ((c.flags & EMLOCO_TRAJ_REAL_PATH) && c.pool && c.pool_count > 0 && U(402) > c.hybrid_init_prob) {
        // DEVIATION: every env draws its pool index independently (WITH replacement); the reference draws the indices of one
        // reset batch with random.sample (traj_generator.py:122, without replacement inside the batch).  The marginal
        // distribution per env is the same (uniform over the pool); two envs reset in the same step may share a trajectory.
        long long pick = (long long)(U(403) * (float)c.pool_count);
        if (pick > c.pool_count - 1) pick = c.pool_count - 1;
        const float* t = c.pool + (size_t)pick * NV * 3;
        const float ox = t[0], oy = t[1];
        float sc = 1.0f;
        if (c.flags & EMLOCO_TRAJ_ADJUST_ROOT_VEL) {
            const float ax = t[3] - t[0], ay = t[4] - t[1], az = t[5] - t[2];
            const float init_speed = fmaxf(sqrtf(ax * ax + ay * ay + az * az), c.speed_min * P.dt);
            sc = __fdiv_rn(root_speed, init_speed) * P.dt;
        }
        for (int i = lane; i < NV; i += 32) {
            v[i][0] = (t[i * 3 + 0] - ox) * sc + px; v[i][1] = (t[i * 3 + 1] - oy) * sc + py; v[i][2] = t[i * 3 + 2];
        }
        __syncwarp();
    }
    // --- align the first segment with the root velocity (:176-234)
    bool inv = false;
    if (c.flags & EMLOCO_TRAJ_INIT_HEADING) {
        const float ox = v[0][0], oy = v[0][1];
        const float dx = v[1][0] - ox, dy = v[1][1] - oy;
        const float root_rot = sqrtf(vx * vx + vy * vy + vz * vz) > 0.0f ? atan2f(vy, vx) : 0.0f;
        const float init_heading = sqrtf(dx * dx + dy * dy) > 0.0f ? atan2f(dy, dx) : 0.0f;
        float rot = init_heading - root_rot;
        if (c.flags & EMLOCO_TRAJ_HEADING_INVERSION) { inv = U(404) > 0.5f; if (inv) rot = init_heading - root_rot + PI; }
        float sn, cs; sincosf(rot, &sn, &cs);
        __syncwarp();
        for (int i = lane; i < NV; i += 32) {
            const float x = v[i][0] - ox, y = v[i][1] - oy;
            v[i][0] = x * cs + y * sn + ox; v[i][1] = -x * sn + y * cs + oy;
        }
        __syncwarp();
    }
    for (int i = lane; i < NV * 3; i += 32) gv[i] = (&v[0][0])[i];
    if (lane == 0 && c.inverted) c.inverted[env] = inv ? 1 : 0;
    // --- _reset_task outputs (:511-516) as the vec-env getters hand them to LocoVal
    if (c.waypoint_traj) {
        float o[3]; lerp_pos(v, (float)P.progress[env] * P.control_dt, P.traj_dur, o);
        const int nw = c.num_waypoints > 0 ? c.num_waypoints : EML_TRAJ_SAMPLES;
        if (lane < nw) {
            float p[3]; lerp_pos(v, (float)P.progress[env] * P.control_dt + (float)lane * P.sample_dt, P.traj_dur, p);
            float* dst = c.waypoint_traj + ((size_t)env * nw + lane) * 3;
#pragma unroll
            for (int k = 0; k < 3; ++k) dst[k] = c.origin_relative ? p[k] - o[k] : p[k];
        }
    }
    if (c.init_pose) {
        const float* rb = P.rb_state + (size_t)env * EML_NB * 13;
        for (int i = lane; i < EML_NB * 3; i += 32) {
            const int b = i / 3, k = i - 3 * b;
            c.init_pose[(size_t)env * EML_NB * 3 + i] = rb[b * 13 + k] - (c.origin_relative ? rb[k] : 0.0f);
        }
    }
    if (c.init_vel && lane < 2) c.init_vel[(size_t)env * 2 + lane] = lane ? vy : vx;
    if (P.clear && lane == 0) { P.reset_w[env] = 0; P.terminate_w[env] = 0; }
}

}  // namespace

cudaError_t eml_traj_reset(emloco_sim* s, const emloco_traj_cfg& c, int clear_flags, cudaStream_t st) {
    TrajParams P;
    P.c = c; P.reset = s->reset; P.progress = s->progress; P.root_state = s->root_state; P.rb_state = s->rb_state;
    P.verts = s->verts; P.epoch = s->traj_epoch; P.N = s->N; P.reset_w = s->reset; P.terminate_w = s->terminate; P.clear = clear_flags;
    const double dt = (double)s->cfg.sim_dt * s->cfg.control_freq_inv;
    const double tdt = ((double)s->cfg.episode_length * dt) / (NV - 1);            // traj_generator.py:24
    P.dt = (float)tdt; P.traj_dur = (float)(NV * tdt); P.control_dt = (float)dt; P.sample_dt = s->cfg.traj_sample_dt;
    traj_reset_kernel<<<(s->N + TR_WARPS - 1) / TR_WARPS, TR_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
}
