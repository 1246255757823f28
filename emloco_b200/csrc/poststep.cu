// Fused post-physics kernel: everything BaseTask.step does after gym.simulate/fetch_results
// (reference pacer/pacer/env/tasks/base_task.py:258-265 -> humanoid_amp.py:139-157 ->
// humanoid.py:1211-1232), i.e. SURVEY 8a rows a4-a9 in ONE launch:
//   progress += 1                                              humanoid.py:1213
//   self obs   compute_humanoid_observations_smpl_max          humanoid.py:1626-1687
//   task obs   _fetch_traj_samples / calc_pos / location obs   humanoid_traj.py:208-224, traj_generator.py:278-296,
//              head-rooted 32x32 height scan, centre heights    humanoid_pedestrain_terrain.py:394-452,732-815,1212-1288
//   flip obs   mirrored self obs + flipped task obs            humanoid.py:1066-1108, ..terrain.py:455-491
//   reward     exp(-2 d^2) - 0.0005 sum|tau qd|                 ..terrain.py:907-930,1581-1592
//   reset      fallen / too-far / episode end (int64)           ..terrain.py:883-905,1468-1530
//   AMP obs    history shift + new 206-float step               humanoid_amp.py:585-657,917-971
//
// HBM-bound: ~37 KB of traffic per env (27 KB of it the AMP history + obs writes), a few hundred
// flops per output.  One 128-thread CTA per env; inputs are staged in shared memory with vector
// loads, the 1422-float observation row is assembled in shared memory and streamed out (normal and
// mirrored) with coalesced 8-byte stores; the AMP ring is shifted through registers.
#include <cuda_bf16.h>
#include "sim.h"

#define PS_THREADS 128

__constant__ float c_grid32[32];     // np.linspace(-2, 2, 32)   (init_square_height_points, ..terrain.py:650-668)
__constant__ float c_cgx[3];         // np.linspace(-0.1, 0.1, 3) (init_center_height_points, ..terrain.py:631-647)
__constant__ float c_cgy[3];         // np.linspace(-0.2, 0.2, 3)
__constant__ int c_l2r[EML_NB] = {0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 19, 20, 21, 22, 23, 14, 15, 16, 17, 18};
__constant__ int c_amp_joint[19] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 19, 20, 21};
__constant__ int c_key_body[4] = {7, 3, 22, 17};   // R_Ankle, L_Ankle, R_Wrist, L_Wrist (pacer.yaml:50)

struct PostParams {
    const float* rb; const float* dof; const float* contact; const float* dof_force;
    const float* verts; const float* betas; const int16_t* height; int hf_rows, hf_cols;
    int64_t* progress; float* obs; float* flip_obs; float* rew; float* rew_raw;
    int64_t* reset; int64_t* terminate; float* amp;
    const float** ring_ptr; // [N] where each env's AMP ring currently lives: its row of `amp`, or of the experience row that the
                            // last post-step wrote when the rows_only sink mode is on (see emloco_post_sinks)
    int N; int advance; int reset_mode; float dt; float traj_dur; float sample_dt; int max_len;
    float power_coef, loc_coef, fail_dist2;
    emloco_post_sinks k;   // optional extra outputs (experience rows, normalised bf16 hi/lo operands of the nets)
};

// normalise like RunningMeanStd.forward (utils/running_mean_std.py:82-84) and split into bf16 hi + lo (csrc/linear_tc.cu)
__device__ __forceinline__ void norm_split2(float x0, float x1, const float2 m, const float2 is, uint32_t& hi, uint32_t& lo) {
    const float a = fminf(fmaxf((x0 - m.x) * is.x, -5.0f), 5.0f), b = fminf(fmaxf((x1 - m.y) * is.y, -5.0f), 5.0f);
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                 // packed conversions: one F2FP per pair
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// mirrored observation (humanoid.py:1066-1108, ..terrain.py:455-491) as a gather: flip_obs[i] = +-obs[src(i)];
// entry = src | (negate << 15), built once on the host from the left/right body permutation
__device__ uint16_t g_flip_src[EML_OBS];

// Terrain.world_points_to_map + sample (..terrain.py:1212-1218,1282-1288); fp32 division and
// truncation exactly as torch does on the host path.
// trunc(x / 0.1f) as torch computes it (IEEE fp32 division, then .long()) without the division subroutine: one Newton step on the
// product with the rounded reciprocal (1 / 0.1f rounds to 10.0f): q0 = 10 x, q = q0 + 10 (x - q0 * 0.1f), fused multiply-adds.
// q differs from the IEEE quotient in the last bit for 0.3 % of the inputs but its INTEGER PART is identical for every float in
// [0, 2^24) (scripts/cu/div_by_tenth_check.cu, exhaustive on the GPU; negative x clamps to cell 0 either way), and only the
// integer part is used.  The height scan evaluates it 2 x 1033 times per env and step: post-step 75.5 -> 72.5 us.
__device__ __forceinline__ float div_by_tenth(float x) {
    const float q0 = __fmul_rn(x, 10.0f);
    return __fmaf_rn(__fmaf_rn(-q0, 0.1f, x), 10.0f, q0);
}
__device__ __forceinline__ float sample_height(const int16_t* __restrict__ hf, int rows, int cols, float x, float y) {
    // (the quotients are far inside the int32 range: clamp first in float so that the conversion cannot overflow)
    int px = __float2int_rz(fminf(fmaxf(div_by_tenth(x), -1.0f), (float)rows));
    int py = __float2int_rz(fminf(fmaxf(div_by_tenth(y), -1.0f), (float)cols));
    px = px < 0 ? 0 : (px > rows - 2 ? rows - 2 : px);
    py = py < 0 ? 0 : (py > cols - 2 ? cols - 2 : py);
    int h1 = __ldg(hf + px * cols + py);
    int h2 = __ldg(hf + (px + 1) * cols + py + 1);
    return (float)(h1 < h2 ? h1 : h2) * 0.005f;
}

// quat_apply(q=(0,0,qz,qw), (bx,by,0)) + pos, contraction-free so the grid index matches torch's (torch_utils.py:49-56)
__device__ __forceinline__ void yaw_apply(float qz, float qw, float bx, float by, float px, float py, float& ox, float& oy) {
    float tx = __fmul_rn(-__fmul_rn(qz, by), 2.0f);
    float ty = __fmul_rn(__fmul_rn(qz, bx), 2.0f);
    float cx = -__fmul_rn(qz, ty);
    float cy = __fmul_rn(qz, tx);
    float rx = __fadd_rn(__fadd_rn(bx, __fmul_rn(qw, tx)), cx);
    float ry = __fadd_rn(__fadd_rn(by, __fmul_rn(qw, ty)), cy);
    ox = __fadd_rn(rx, px);
    oy = __fadd_rn(ry, py);
}

// TrajGenerator.calc_pos (traj_generator.py:278-296)
__device__ __forceinline__ f3 calc_pos(const float* __restrict__ verts, float t, float traj_dur) {
    float phase = fminf(fmaxf(__fdiv_rn(t, traj_dur), 0.0f), 1.0f);
    float seg = phase * (float)(EML_NUM_VERTS - 1);
    float f0 = floorf(seg), f1 = ceilf(seg);
    int i0 = (int)f0, i1 = (int)f1;
    float l = seg - f0;
    const float* a = verts + i0 * 3;
    const float* b = verts + i1 * 3;
    float w = 1.0f - l;
    return mk3(w * a[0] + l * b[0], w * a[1] + l * b[1], w * a[2] + l * b[2]);
}

__global__ void __launch_bounds__(PS_THREADS) post_step_kernel(PostParams P) {
    __shared__ __align__(16) float s_rb[EML_NB * 13];
    __shared__ __align__(16) float s_dof[EML_ND * 2];
    __shared__ __align__(16) float s_obs[EML_OBS + 2];
    __shared__ __align__(16) float s_amp[EML_AMP_STEP];
    __shared__ float s_misc[16];   // 0-3 hinv, 4-5 head yaw (z,w), 6-7 root yaw (z,w), 8 centre height, 9 time

    const int env = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (env >= P.N) return;
    // reset mode (humanoid.py:455-465 + humanoid_amp.py:284-293,499-502): only envs flagged in reset_buf; progress := 0,
    // observations recomputed, AMP history filled with the current step, reward untouched, flags cleared
    const bool rmode = P.reset_mode != 0;
    if (rmode && P.reset[env] == 0) return;

    // ---- stage state ----
    {
        const float4* g = reinterpret_cast<const float4*>(P.rb + (size_t)env * EML_NB * 13);
        if (tid < 78) reinterpret_cast<float4*>(s_rb)[tid] = __ldg(g + tid);
        const float2* d = reinterpret_cast<const float2*>(P.dof + (size_t)env * EML_ND * 2);
        if (tid < EML_ND) reinterpret_cast<float2*>(s_dof)[tid] = __ldg(d + tid);
    }
    // AMP history: hist[k+1] = old[k] (humanoid_amp.py:585-594); read now, store after the last barrier
    float2 hist[12];
    {
        const float2* a = reinterpret_cast<const float2*>(rmode ? P.amp + (size_t)env * EML_AMP_OBS : P.ring_ptr[env]);
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            int i = tid + PS_THREADS * k;
            if (i < 14 * 103) hist[k] = __ldcs(a + i);      // streamed once: keep it out of the way of the statistics in L1
        }
    }
    long long prog = rmode ? 0 : P.progress[env] + P.advance;
    __syncthreads();

    const f3 root_pos = mk3(s_rb[0], s_rb[1], s_rb[2]);
    if (tid == 0) {
        f4 rr = mk4(s_rb[3], s_rb[4], s_rb[5], s_rb[6]);
        f4 hinv = quat_from_angle_z(-calc_heading(rr));
        s_misc[0] = hinv.x; s_misc[1] = hinv.y; s_misc[2] = hinv.z; s_misc[3] = hinv.w;
        if (P.advance || rmode) P.progress[env] = prog;
        s_misc[9] = (float)prog * P.dt;
    } else if (tid == 32) {
        const float* h = s_rb + EML_HEAD * 13;
        f4 hq = quat_from_angle_z(calc_heading(mk4(h[3], h[4], h[5], h[6])));
        s_misc[4] = hq.z; s_misc[5] = hq.w;
    } else if (tid == 64) {
        // quat_apply_yaw: zero x,y then normalise (..terrain.py:1533-1538)
        float qz = s_rb[5], qw = s_rb[6];
        float n = fmaxf(sqrtf(qz * qz + qw * qw), 1e-9f);
        s_misc[6] = qz / n; s_misc[7] = qw / n;
    }
    __syncthreads();
    const f4 hinv = mk4(s_misc[0], s_misc[1], s_misc[2], s_misc[3]);
    const float time0 = s_misc[9];
    const float* verts = P.verts + (size_t)env * EML_NUM_VERTS * 3;

    if (warp == 0) {
        // ---- self observation, one lane per body (humanoid.py:1626-1687) ----
        if (lane < EML_NB) {
            const float* b = s_rb + lane * 13;
            f3 lp = quat_rotate(hinv, mk3(b[0], b[1], b[2]) - root_pos);
            if (lane > 0) { float* o = s_obs + (lane - 1) * 3; o[0] = lp.x; o[1] = lp.y; o[2] = lp.z; }
            f4 lr = quat_mul(hinv, mk4(b[3], b[4], b[5], b[6]));
            quat_to_tan_norm(lr, s_obs + 69 + lane * 6);
            f3 lv = quat_rotate(hinv, mk3(b[7], b[8], b[9]));
            f3 lw = quat_rotate(hinv, mk3(b[10], b[11], b[12]));
            float* ov = s_obs + 213 + lane * 3; ov[0] = lv.x; ov[1] = lv.y; ov[2] = lv.z;
            float* ow = s_obs + 285 + lane * 3; ow[0] = lw.x; ow[1] = lw.y; ow[2] = lw.z;
            // (the compiler widens the 6-float copy below to aligned 128-bit loads that also touch neighbouring lanes' elements:
            //  order them after those lanes' stores so that racecheck stays clean)
            __syncwarp(0x00ffffffu);
            if (lane == 0) {   // AMP root block reuses the same quantities (humanoid_amp.py:924-938)
                for (int k = 0; k < 6; ++k) s_amp[k] = s_obs[69 + k];
                s_amp[6] = lv.x; s_amp[7] = lv.y; s_amp[8] = lv.z;
                s_amp[9] = lw.x; s_amp[10] = lw.y; s_amp[11] = lw.z;
            }
        }
    } else if (warp == 1) {
        // ---- trajectory samples, target, reward-location, centre height ----
        if (lane < EML_TRAJ_SAMPLES) {
            float t = time0 + (float)lane * P.sample_dt;
            f3 s = calc_pos(verts, t, P.traj_dur);
            f3 l = quat_rotate(hinv, s - root_pos);
            s_obs[EML_SELF_OBS + 2 * lane] = l.x;
            s_obs[EML_SELF_OBS + 2 * lane + 1] = l.y;
        }
        float ch = 0.f;
        if (lane >= 16 && lane < 25) {
            int i = lane - 16;
            float x, y;
            yaw_apply(s_misc[6], s_misc[7], c_cgx[i / 3], c_cgy[i % 3], root_pos.x, root_pos.y, x, y);
            ch = sample_height(P.height, P.hf_rows, P.hf_cols, x, y);
        }
        ch = warp_sum(ch);
        if (lane == 0) s_misc[8] = ch / 9.0f;
    } else if (warp == 2) {
        // ---- AMP step: joints (exp-map -> quat -> tan/norm), dof vel subset, key bodies ----
        if (lane < 19) {
            int j = c_amp_joint[lane];
            f3 e = mk3(s_dof[(3 * j) * 2], s_dof[(3 * j + 1) * 2], s_dof[(3 * j + 2) * 2]);
            quat_to_tan_norm(exp_map_to_quat(e), s_amp + 12 + lane * 6);
            s_amp[126 + lane * 3 + 0] = s_dof[(3 * j) * 2 + 1];
            s_amp[126 + lane * 3 + 1] = s_dof[(3 * j + 1) * 2 + 1];
            s_amp[126 + lane * 3 + 2] = s_dof[(3 * j + 2) * 2 + 1];
        } else if (lane < 23) {
            int k = lane - 19;
            const float* b = s_rb + c_key_body[k] * 13;
            f3 l = quat_rotate(hinv, mk3(b[0], b[1], b[2]) - root_pos);
            s_amp[183 + k * 3] = l.x; s_amp[184 + k * 3] = l.y; s_amp[185 + k * 3] = l.z;
        }
    } else {
        // ---- reward, reset, shape parameters ----
        float pw = 0.f;
        const float* df = P.dof_force + (size_t)env * EML_ND;
        for (int i = lane; i < EML_ND; i += 32) pw += fabsf(__ldg(df + i) * s_dof[2 * i + 1]);
        pw = warp_sum(pw);
        f3 cs = mk3(0.f, 0.f, 0.f);
        if (lane < EML_NB && lane != 3 && lane != 4 && lane != 7 && lane != 8) {   // contactBodies masked (pacer.yaml:51)
            const float* c = P.contact + ((size_t)env * EML_NB + lane) * 3;
            cs = mk3(__ldg(c), __ldg(c + 1), __ldg(c + 2));
        }
        cs.x = warp_sum(cs.x); cs.y = warp_sum(cs.y); cs.z = warp_sum(cs.z);
        if (lane < 11) {
            float b = __ldg(P.betas + (size_t)env * 17 + lane);
            s_obs[357 + lane] = b;
            s_amp[195 + lane] = b;
        }
        if (lane == 31 && !rmode) {
            f3 tar = calc_pos(verts, time0, P.traj_dur);
            float dx = tar.x - root_pos.x, dy = tar.y - root_pos.y;
            float err = dx * dx + dy * dy;
            float loc = P.loc_coef * expf(-2.0f * err);
            float pr = -P.power_coef * pw;
            P.rew[env] = loc + pr;
            P.rew_raw[2 * env] = loc; P.rew_raw[2 * env + 1] = pr;
            bool fallen = sqrtf(cs.x * cs.x + cs.y * cs.y + cs.z * cs.z) > 50.0f && prog > 1;
            bool fail = err > P.fail_dist2;
            long long term = (fallen || fail) ? 1 : 0;
            P.terminate[env] = term;
            P.reset[env] = (prog >= P.max_len - 1) ? 1 : term;
        }
    }
    __syncthreads();

    // ---- head-rooted 32x32 height scan (..terrain.py:761-815), 8 points per thread ----
    {
        const float* h = s_rb + EML_HEAD * 13;
        const float hx = h[0], hy = h[1], qz = s_misc[4], qw = s_misc[5], centre = s_misc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int i = tid + PS_THREADS * k;
            float x, y;
            yaw_apply(qz, qw, c_grid32[i >> 5], c_grid32[i & 31], hx, hy, x, y);
            float m = sample_height(P.height, P.hf_rows, P.hf_cols, x, y);
            s_obs[EML_SELF_OBS + 30 + i] = fminf(fmaxf(centre - m, -3.0f), 3.0f) * 5.0f;
        }
    }
    __syncthreads();

    // ---- stream out: obs, mirrored obs, AMP ring ----
    {
        float2* o = reinterpret_cast<float2*>(P.obs + (size_t)env * EML_OBS);
        float2* fc = (P.k.flip_copy && !rmode) ? reinterpret_cast<float2*>(P.k.flip_copy + (size_t)env * EML_OBS) : nullptr;
        // rows_only: the experience rows are the only destination of the mirrored observation and of the AMP ring
        float2* f = (fc && P.k.rows_only) ? nullptr : reinterpret_cast<float2*>(P.flip_obs + (size_t)env * EML_OBS);
        // The table entries and normalisation statistics are the same for every env: their (L1/L2) loads are issued for all
        // of a thread's elements first, so that the stores below never wait on a load issued one instruction earlier
        constexpr int OB_IT = (EML_OBS / 2 + PS_THREADS - 1) / PS_THREADS;            // 6
        {
            uint32_t e[OB_IT];
#pragma unroll
            for (int k = 0; k < OB_IT; ++k) {
                const int i2 = tid + PS_THREADS * k;
                e[k] = i2 < EML_OBS / 2 ? __ldg(reinterpret_cast<const uint32_t*>(g_flip_src) + i2) : 0u;      // two table entries
            }
            float2* oc = P.k.obs_copy ? reinterpret_cast<float2*>(P.k.obs_copy + (size_t)env * EML_OBS) : nullptr;
#pragma unroll
            for (int k = 0; k < OB_IT; ++k) {
                const int i2 = tid + PS_THREADS * k;
                if (i2 < EML_OBS / 2) {
                    const float2 v = reinterpret_cast<const float2*>(s_obs)[i2];
                    o[i2] = v;
                    if (oc) oc[i2] = v;                                    // experience row of the same observation
                    const float r0 = s_obs[e[k] & 0x7fffu], r1 = s_obs[(e[k] >> 16) & 0x7fffu];
                    const float2 fv = make_float2((e[k] & 0x8000u) ? -r0 : r0, (e[k] & 0x80000000u) ? -r1 : r1);
                    if (f) f[i2] = fv;
                    if (fc) fc[i2] = fv;
                }
            }
        }
        // optional sink: the normalised bf16 hi/lo operands the tensor-core layers read (self-obs part -> actor/critic input,
        // task-obs part -> task MLP input)
        if (P.k.self_hi) {
            float2 m[OB_IT], is[OB_IT];
#pragma unroll
            for (int k = 0; k < OB_IT; ++k) {
                const int i2 = tid + PS_THREADS * k;
                if (i2 < EML_OBS / 2) {
                    m[k] = __ldg(reinterpret_cast<const float2*>(P.k.obs_mean) + i2);
                    is[k] = __ldg(reinterpret_cast<const float2*>(P.k.obs_inv_std) + i2);
                }
            }
#pragma unroll
            for (int k = 0; k < OB_IT; ++k) {
                const int i2 = tid + PS_THREADS * k, i = 2 * i2;
                if (i2 < EML_OBS / 2) {
                    uint32_t hi, lo;
                    norm_split2(s_obs[i], s_obs[i + 1], m[k], is[k], hi, lo);
                    const bool second = P.k.self_hi2 && !rmode;           // the next-observation copy is not touched by resets
                    if (i < EML_SELF_OBS) {
                        *reinterpret_cast<uint32_t*>(P.k.self_hi + (size_t)env * P.k.ld_self + i) = hi;
                        *reinterpret_cast<uint32_t*>(P.k.self_lo + (size_t)env * P.k.ld_self + i) = lo;
                        if (second) {
                            *reinterpret_cast<uint32_t*>(P.k.self_hi2 + (size_t)env * P.k.ld_self + i) = hi;
                            *reinterpret_cast<uint32_t*>(P.k.self_lo2 + (size_t)env * P.k.ld_self + i) = lo;
                        }
                    } else {
                        *reinterpret_cast<uint32_t*>(P.k.task_hi + (size_t)env * P.k.ld_task + (i - EML_SELF_OBS)) = hi;
                        *reinterpret_cast<uint32_t*>(P.k.task_lo + (size_t)env * P.k.ld_task + (i - EML_SELF_OBS)) = lo;
                        if (second) {
                            *reinterpret_cast<uint32_t*>(P.k.task_hi2 + (size_t)env * P.k.ld_task + (i - EML_SELF_OBS)) = hi;
                            *reinterpret_cast<uint32_t*>(P.k.task_lo2 + (size_t)env * P.k.ld_task + (i - EML_SELF_OBS)) = lo;
                        }
                    }
                }
            }
        }
        float2* a = reinterpret_cast<float2*>(P.amp + (size_t)env * EML_AMP_OBS);
        if (rmode) {
            for (int i = tid; i < 15 * 103; i += PS_THREADS) a[i] = reinterpret_cast<const float2*>(s_amp)[i % 103];
            if (tid == 0) {
                P.ring_ptr[env] = P.amp + (size_t)env * EML_AMP_OBS;
                if (P.reset_mode == 1) { P.reset[env] = 0; P.terminate[env] = 0; }
            }
            return;
        }
        float2* ac = P.k.amp_copy ? reinterpret_cast<float2*>(P.k.amp_copy + (size_t)env * EML_AMP_OBS) : nullptr;
        if (ac && P.k.rows_only) {                 // the next step finds the ring in this step's experience row
            a = nullptr;
            if (tid == 0) P.ring_ptr[env] = P.k.amp_copy + (size_t)env * EML_AMP_OBS;
        } else if (tid == 0) {
            P.ring_ptr[env] = P.amp + (size_t)env * EML_AMP_OBS;
        }
        uint32_t* ah = P.k.amp_hi ? reinterpret_cast<uint32_t*>(P.k.amp_hi + (size_t)env * P.k.ld_amp) : nullptr;
        uint32_t* al = P.k.amp_hi ? reinterpret_cast<uint32_t*>(P.k.amp_lo + (size_t)env * P.k.ld_amp) : nullptr;
        // statistics for the thread's 12 ring elements, fetched in groups of 4 ahead of their use
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float2 m[4], is[4];
            if (ah) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = tid + PS_THREADS * (4 * g + j);
                    if (i < 14 * 103) {
                        m[j] = __ldg(reinterpret_cast<const float2*>(P.k.amp_mean) + 103 + i);
                        is[j] = __ldg(reinterpret_cast<const float2*>(P.k.amp_inv_std) + 103 + i);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * g + j, i = tid + PS_THREADS * k;
                if (i < 14 * 103) {
                    if (a) a[103 + i] = hist[k];
                    if (ac) ac[103 + i] = hist[k];
                    if (ah) { uint32_t hi, lo; norm_split2(hist[k].x, hist[k].y, m[j], is[j], hi, lo); ah[103 + i] = hi; al[103 + i] = lo; }
                }
            }
        }
        if (tid < 103) {
            const float2 v = reinterpret_cast<const float2*>(s_amp)[tid];
            if (a) a[tid] = v;
            if (ac) ac[tid] = v;
            if (ah) {
                uint32_t hi, lo;
                norm_split2(v.x, v.y, __ldg(reinterpret_cast<const float2*>(P.k.amp_mean) + tid), __ldg(reinterpret_cast<const float2*>(P.k.amp_inv_std) + tid), hi, lo);
                ah[tid] = hi; al[tid] = lo;
            }
        }
    }
}

static bool g_tables_ready = false;

static cudaError_t launch_post(emloco_sim* s, int advance_progress, int reset_mode, cudaStream_t st);
cudaError_t eml_launch_post_step(emloco_sim* s, int advance_progress, cudaStream_t st) { return launch_post(s, advance_progress, 0, st); }
cudaError_t eml_launch_post_reset(emloco_sim* s, int keep_flags, cudaStream_t st) { return launch_post(s, 0, keep_flags ? 2 : 1, st); }

static cudaError_t launch_post(emloco_sim* s, int advance_progress, int reset_mode, cudaStream_t st) {
    if (!g_tables_ready) {
        float g32[32], cx[3], cy[3];
        for (int i = 0; i < 32; ++i) g32[i] = (float)(-2.0 + (4.0 / 31.0) * i);   // np.linspace step form: start + i*step
        g32[31] = 2.0f;
        double sx = 0.2 / 2.0, sy = 0.4 / 2.0;
        for (int i = 0; i < 3; ++i) { cx[i] = (float)(-0.1 + sx * i); cy[i] = (float)(-0.2 + sy * i); }
        cx[2] = 0.1f; cy[2] = 0.2f;
        cudaError_t e;
        {   // mirrored-observation gather table
            static const int l2r[EML_NB] = {0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 19, 20, 21, 22, 23, 14, 15, 16, 17, 18};
            uint16_t tab[EML_OBS];
            for (int i = 0; i < EML_OBS; ++i) {
                int src = i, neg = 0;
                if (i < 69) { int b = i / 3 + 1, c = i % 3; src = (l2r[b] - 1) * 3 + c; neg = c == 1; }                         // local body pos: y negated
                else if (i < 213) { int j = i - 69, b = j / 6, c = j % 6; src = 69 + l2r[b] * 6 + c; neg = c % 3 == 1; }        // tan/norm: y components
                else if (i < 285) { int j = i - 213, b = j / 3, c = j % 3; src = 213 + l2r[b] * 3 + c; neg = c == 1; }          // linear velocity
                else if (i < 357) { int j = i - 285, b = j / 3, c = j % 3; src = 285 + l2r[b] * 3 + c; neg = c != 1; }          // angular velocity (pseudo-vector)
                else if (i < 368) { src = i; }                                                                                  // shape parameters
                else if (i < 398) { src = i; neg = (i - 368) & 1; }                                                             // trajectory samples: y negated
                else { int j = i - 398; src = 398 + (j & ~31) + (31 - (j & 31)); }                                              // height map flipped along y
                tab[i] = (uint16_t)(src | (neg << 15));
            }
            if ((e = cudaMemcpyToSymbol(g_flip_src, tab, sizeof(tab))) != cudaSuccess) return e;
        }
        if ((e = cudaMemcpyToSymbol(c_grid32, g32, sizeof(g32))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_cgx, cx, sizeof(cx))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_cgy, cy, sizeof(cy))) != cudaSuccess) return e;
        g_tables_ready = true;
    }
    PostParams P;
    P.rb = s->rb_state; P.dof = s->dof_state; P.contact = s->contact; P.dof_force = s->dof_force;
    P.verts = s->verts; P.betas = s->betas; P.height = s->height; P.hf_rows = s->hf_rows; P.hf_cols = s->hf_cols;
    P.progress = s->progress; P.obs = s->obs; P.flip_obs = s->flip_obs; P.rew = s->rew; P.rew_raw = s->rew_raw;
    P.reset = s->reset; P.terminate = s->terminate; P.amp = s->amp_obs; P.ring_ptr = s->ring_ptr;
    P.N = s->N; P.advance = advance_progress; P.reset_mode = reset_mode; P.k = s->sinks;
    double dt = (double)s->cfg.control_freq_inv * (double)s->cfg.sim_dt;          // humanoid.py:89
    P.dt = (float)dt;
    double tdt = ((double)s->cfg.episode_length * dt) / (EML_NUM_VERTS - 1);      // traj_generator.py:24
    P.traj_dur = (float)(EML_NUM_VERTS * tdt);                                    // traj_generator.py:269-272
    P.sample_dt = s->cfg.traj_sample_dt;
    P.max_len = s->cfg.episode_length;
    P.power_coef = s->cfg.power_coefficient; P.loc_coef = s->cfg.location_coefficient;
    P.fail_dist2 = s->cfg.fail_dist * s->cfg.fail_dist;
    post_step_kernel<<<s->N, PS_THREADS, 0, st>>>(P);
    return cudaGetLastError();
}
