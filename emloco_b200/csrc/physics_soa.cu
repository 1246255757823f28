// Physics step, lane-per-env formulation (the default stepping kernel; physics.cu keeps the warp-per-env kernel for the
// forward-kinematics / reset paths and as a cross-check).  Same algorithm and arithmetic as physics.cu / oracle/physics_oracle.c
// (reference seam: gym.simulate x controlFrequencyInv, pacer/pacer/env/tasks/base_task.py:792-797, and pre_physics_step,
// humanoid.py:1184-1209) - only the mapping onto the machine differs:
//
//   CTA = 32 envs x 12 warps.  LANE = ENV: every per-body quantity is a struct-of-arrays [field][32 envs] in shared memory
//   (bank-conflict free, no shuffles), so all 32 lanes of a warp do the same body of 32 different envs - the tree recursions
//   run at full lane occupancy instead of the 2-5 active lanes per level of the warp-per-env mapping (7x fewer issue slots).
//   WARPS = the 5 kinematic chains (legs, spine+head, arms) for the tree passes, all 12 warps x 2 bodies for the per-body
//   work (inertia, contacts, drive).  ~211 KB of shared memory per CTA: one CTA per SM, 128 CTAs for 4096 envs.
//
// Per integration part (a sub-step, or a piece of one when the adaptive refinement triggers):
//   A0  chains   kinematics root->leaf in the frame translating with the pelvis (arms recompute torso..chest on the way)
//   A1  bodies   spatial inertia about the pelvis, bias force, gravity, ground contact folded in, implicit PD drive
//   B1  chains   articulated inertias leaf->root inside each chain (spine chain: head, neck), chain totals to smem
//   B2  1 warp   chest, spine, torso (+ arm totals), pelvis (+ leg totals): 6x6 root solve; accelerations and integration
//                of torso..chest and of the root
//   C   chains   accelerations root->leaf, contact force at the end-of-part velocity, drive torque, joint integration
#include "sim.h"
#include "physics_math.cuh"

#define SOA_WARPS 12        // 12 x 166 registers x 32 lanes just fit the register file; the per-body phases run 2 bodies per warp
#define SOA_BPW (EML_NB / SOA_WARPS)
#define SOA_THREADS (SOA_WARPS * 32)
// tree-pass loops (B1, B2, C): NOT unrolled - unrolling by two (168 registers) made the kernel 15 % slower (116 vs 100 us): the
// compiler does not interleave consecutive bodies and the loop body outgrows the instruction cache
#ifndef SOA_TREE_UNROLL_N
#define SOA_TREE_UNROLL_N 1
#endif
#define SOA_PRAGMA_(x) _Pragma(#x)
#define SOA_PRAGMA(x) SOA_PRAGMA_(x)
#define SOA_TREE_UNROLL SOA_PRAGMA(unroll SOA_TREE_UNROLL_N)

// floats per body in shared memory
enum {
    F_JQ = 0,    // 4  joint rotation (state)
    F_JW = 4,    // 3  joint rate, child frame (state)
    F_X = 7,     // 3  anchor position relative to the pelvis, world axes
    F_QW = 10,   // 4  world orientation
    F_VW = 14,   // 3  angular velocity
    F_VL = 17,   // 3  linear velocity at the pelvis point, in the pelvis-translating frame
    F_C = 20,    // 6  velocity-product acceleration (cw, cl)
    F_DYN = 26,  // 27 A(6) B(9) M(6) pn(3) pf(3)   ->   after pass B: Ut(9) Ub(9) Di(6) u(3)
    F_TAU = 53,  // 3  drive torque at the start-of-part state, world axes
    F_DD = 56,   // 1  implicit drive inertia (saturation applied)
    F_CS = 57,   // 8  contact sums F0z Sbt Sbn Stz Sty Stx Sny Snx
    F_PER_BODY = 65
};
#define SOA_BODY_FLOATS (F_PER_BODY * EML_NB * 32)
#define SOA_ROOT (SOA_BODY_FLOATS)                 // 2 x 13 x 32: root state, double-buffered by part parity
#define SOA_XCHG (SOA_ROOT + 2 * 13 * 32)          // 4 chains x 27 x 32: chain totals (legs -> pelvis, arms -> chest)
#define SOA_ACC (SOA_XCHG + 4 * 27 * 32)           // 2 x 6 x 32: pelvis and chest accelerations
#define SOA_WMAX (SOA_ACC + 2 * 6 * 32)            // 5 x 32: per-chain max body angular speed
#define SOA_KMAX (SOA_WMAX + 5 * 32)               // SOA_WARPS ints: per-warp largest piece count
#define SOA_FSUM (SOA_KMAX + SOA_WARPS)                    // 24 x 3 x 32: contact force accumulated over the parts
#define SOA_KIN (SOA_FSUM + EML_NB * 3 * 32)         // 13 x 32: the chest's kinematic state of the coming part, for the arm chains
#define SOA_FLOATS (SOA_KIN + 13 * 32)
#define SOA_SMEM_BYTES (SOA_FLOATS * 4)

__constant__ int c_chain_len[5] = {4, 4, 5, 5, 5};
__constant__ int c_chain_body[5][5] = {{1, 2, 3, 4, 0}, {5, 6, 7, 8, 0}, {9, 10, 11, 12, 13}, {14, 15, 16, 17, 18}, {19, 20, 21, 22, 23}};

struct Sp { S3 A; M3 B; S3 M; f3 pn, pf; };          // a spatial inertia (about the pelvis, world axes) and a bias force

#define SM(b, f) smem[((b) * F_PER_BODY + (f)) * 32 + lane]

__device__ __forceinline__ f3 ld3(const float* smem, int lane, int b, int f) { return mk3(SM(b, f), SM(b, f + 1), SM(b, f + 2)); }
__device__ __forceinline__ f4 ld4(const float* smem, int lane, int b, int f) { return mk4(SM(b, f), SM(b, f + 1), SM(b, f + 2), SM(b, f + 3)); }
__device__ __forceinline__ void st3(float* smem, int lane, int b, int f, f3 v) { SM(b, f) = v.x; SM(b, f + 1) = v.y; SM(b, f + 2) = v.z; }
__device__ __forceinline__ void st4(float* smem, int lane, int b, int f, f4 v) { SM(b, f) = v.x; SM(b, f + 1) = v.y; SM(b, f + 2) = v.z; SM(b, f + 3) = v.w; }

__device__ __forceinline__ void ld_sp(const float* smem, int lane, int b, Sp& s) {
    const int f = F_DYN;
    s.A.xx = SM(b, f); s.A.xy = SM(b, f + 1); s.A.xz = SM(b, f + 2); s.A.yy = SM(b, f + 3); s.A.yz = SM(b, f + 4); s.A.zz = SM(b, f + 5);
#pragma unroll
    for (int k = 0; k < 9; ++k) s.B.a[k] = SM(b, f + 6 + k);
    s.M.xx = SM(b, f + 15); s.M.xy = SM(b, f + 16); s.M.xz = SM(b, f + 17); s.M.yy = SM(b, f + 18); s.M.yz = SM(b, f + 19); s.M.zz = SM(b, f + 20);
    s.pn = mk3(SM(b, f + 21), SM(b, f + 22), SM(b, f + 23)); s.pf = mk3(SM(b, f + 24), SM(b, f + 25), SM(b, f + 26));
}
__device__ __forceinline__ void st_sp(float* smem, int lane, int b, const Sp& s) {
    const int f = F_DYN;
    SM(b, f) = s.A.xx; SM(b, f + 1) = s.A.xy; SM(b, f + 2) = s.A.xz; SM(b, f + 3) = s.A.yy; SM(b, f + 4) = s.A.yz; SM(b, f + 5) = s.A.zz;
#pragma unroll
    for (int k = 0; k < 9; ++k) SM(b, f + 6 + k) = s.B.a[k];
    SM(b, f + 15) = s.M.xx; SM(b, f + 16) = s.M.xy; SM(b, f + 17) = s.M.xz; SM(b, f + 18) = s.M.yy; SM(b, f + 19) = s.M.yz; SM(b, f + 20) = s.M.zz;
    SM(b, f + 21) = s.pn.x; SM(b, f + 22) = s.pn.y; SM(b, f + 23) = s.pn.z; SM(b, f + 24) = s.pf.x; SM(b, f + 25) = s.pf.y; SM(b, f + 26) = s.pf.z;
}
__device__ __forceinline__ void add_sp(Sp& a, const Sp& b) {
    a.A.xx += b.A.xx; a.A.xy += b.A.xy; a.A.xz += b.A.xz; a.A.yy += b.A.yy; a.A.yz += b.A.yz; a.A.zz += b.A.zz;
#pragma unroll
    for (int k = 0; k < 9; ++k) a.B.a[k] += b.B.a[k];
    a.M.xx += b.M.xx; a.M.xy += b.M.xy; a.M.xz += b.M.xz; a.M.yy += b.M.yy; a.M.yz += b.M.yz; a.M.zz += b.M.zz;
    a.pn = a.pn + b.pn; a.pf = a.pf + b.pf;
}
// chain totals through the exchange area: slot in 0..3, 27 floats
__device__ __forceinline__ void st_xchg(float* smem, int lane, int slot, const Sp& s) {
    float* x = smem + SOA_XCHG + slot * 27 * 32 + lane;
    x[0] = s.A.xx; x[32] = s.A.xy; x[64] = s.A.xz; x[96] = s.A.yy; x[128] = s.A.yz; x[160] = s.A.zz;
#pragma unroll
    for (int k = 0; k < 9; ++k) x[(6 + k) * 32] = s.B.a[k];
    x[15 * 32] = s.M.xx; x[16 * 32] = s.M.xy; x[17 * 32] = s.M.xz; x[18 * 32] = s.M.yy; x[19 * 32] = s.M.yz; x[20 * 32] = s.M.zz;
    x[21 * 32] = s.pn.x; x[22 * 32] = s.pn.y; x[23 * 32] = s.pn.z; x[24 * 32] = s.pf.x; x[25 * 32] = s.pf.y; x[26 * 32] = s.pf.z;
}
__device__ __forceinline__ void add_xchg(const float* smem, int lane, int slot, Sp& s) {
    const float* x = smem + SOA_XCHG + slot * 27 * 32 + lane;
    s.A.xx += x[0]; s.A.xy += x[32]; s.A.xz += x[64]; s.A.yy += x[96]; s.A.yz += x[128]; s.A.zz += x[160];
#pragma unroll
    for (int k = 0; k < 9; ++k) s.B.a[k] += x[(6 + k) * 32];
    s.M.xx += x[15 * 32]; s.M.xy += x[16 * 32]; s.M.xz += x[17 * 32]; s.M.yy += x[18 * 32]; s.M.yz += x[19 * 32]; s.M.zz += x[20 * 32];
    s.pn.x += x[21 * 32]; s.pn.y += x[22 * 32]; s.pn.z += x[23 * 32]; s.pf.x += x[24 * 32]; s.pf.y += x[25 * 32]; s.pf.z += x[26 * 32];
}

struct Kin { f4 q; f3 x, w, l; };                    // world orientation, anchor, angular / linear velocity

// one kinematic step parent -> child (physics.cu pass 1)
__device__ __forceinline__ Kin kin_step(const Kin& p, f3 offset, f4 jq, f3 jw, f3& cw, f3& cl) {
    Kin k;
    k.x = p.x + qrot(p.q, offset);
    k.q = qmul(p.q, jq);
    f3 ww = qrot(k.q, jw);                               // joint rate in world axes
    f3 jl = cross3(k.x, ww);
    k.w = p.w + ww; k.l = p.l + jl;
    cw = cross3(p.w, ww);                                // c = v_parent x vJ
    cl = cross3(p.w, jl) + cross3(p.l, ww);
    return k;
}

__device__ __forceinline__ void store_kin(float* smem, int lane, int b, const Kin& k, f3 cw, f3 cl) {
    st3(smem, lane, b, F_X, k.x); st4(smem, lane, b, F_QW, k.q); st3(smem, lane, b, F_VW, k.w); st3(smem, lane, b, F_VL, k.l);
    st3(smem, lane, b, F_C, cw); st3(smem, lane, b, F_C + 3, cl);
}
__device__ __forceinline__ void put_kin(float* smem, int lane, const Kin& k) {
    float* p = smem + SOA_KIN + lane;
    p[0] = k.q.x; p[32] = k.q.y; p[64] = k.q.z; p[96] = k.q.w; p[128] = k.x.x; p[160] = k.x.y; p[192] = k.x.z;
    p[224] = k.w.x; p[256] = k.w.y; p[288] = k.w.z; p[320] = k.l.x; p[352] = k.l.y; p[384] = k.l.z;
}
__device__ __forceinline__ Kin get_kin(const float* smem, int lane) {
    const float* p = smem + SOA_KIN + lane;
    Kin k; k.q = mk4(p[0], p[32], p[64], p[96]); k.x = mk3(p[128], p[160], p[192]);
    k.w = mk3(p[224], p[256], p[288]); k.l = mk3(p[320], p[352], p[384]);
    return k;
}

// articulated-body step of one joint (physics.cu pass 2 body): consumes the body's accumulated (IA, pA), leaves
// U, Dinv, u in the body's F_DYN slot for pass C and returns what the parent has to add
__device__ __forceinline__ void aba_body(float* smem, int lane, int b, Sp& s, float arm) {
    const f3 x = ld3(smem, lane, b, F_X);
    const f3 cw = ld3(smem, lane, b, F_C), cl = ld3(smem, lane, b, F_C + 3);
    const f3 tau0 = ld3(smem, lane, b, F_TAU);
    const float dd = arm + SM(b, F_DD);
    M3 Ut, Ub;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        setrow(Ut, r, srow(s.A, r) + cross3(row(s.B, r), x));
        setrow(Ub, r, col(s.B, r) + cross3(srow(s.M, r), x));
    }
    S3 D;
    f3 d0 = col(Ut, 0) - cross3(x, col(Ub, 0));
    f3 d1 = col(Ut, 1) - cross3(x, col(Ub, 1));
    f3 d2 = col(Ut, 2) - cross3(x, col(Ub, 2));
    D.xx = d0.x + dd; D.xy = 0.5f * (d0.y + d1.x); D.xz = 0.5f * (d0.z + d2.x);
    D.yy = d1.y + dd; D.yz = 0.5f * (d1.z + d2.y); D.zz = d2.z + dd;
    const S3 Di = inv_s3(D);
    const f3 u = tau0 - s.pn + cross3(x, s.pf);
    M3 Wt, Wb;
#pragma unroll
    for (int r = 0; r < 3; ++r) { setrow(Wt, r, sv(Di, row(Ut, r))); setrow(Wb, r, sv(Di, row(Ub, r))); }
    s.A.xx -= dot3(row(Wt, 0), row(Ut, 0)); s.A.xy -= dot3(row(Wt, 0), row(Ut, 1)); s.A.xz -= dot3(row(Wt, 0), row(Ut, 2));
    s.A.yy -= dot3(row(Wt, 1), row(Ut, 1)); s.A.yz -= dot3(row(Wt, 1), row(Ut, 2)); s.A.zz -= dot3(row(Wt, 2), row(Ut, 2));
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c2 = 0; c2 < 3; ++c2) s.B.a[3 * r + c2] -= dot3(row(Wt, r), row(Ub, c2));
    s.M.xx -= dot3(row(Wb, 0), row(Ub, 0)); s.M.xy -= dot3(row(Wb, 0), row(Ub, 1)); s.M.xz -= dot3(row(Wb, 0), row(Ub, 2));
    s.M.yy -= dot3(row(Wb, 1), row(Ub, 1)); s.M.yz -= dot3(row(Wb, 1), row(Ub, 2)); s.M.zz -= dot3(row(Wb, 2), row(Ub, 2));
    s.pn = s.pn + sv(s.A, cw) + mv(s.B, cl) + mv(Wt, u);
    s.pf = s.pf + mtv(s.B, cw) + sv(s.M, cl) + mv(Wb, u);
    // keep U, Dinv, u for pass C (overwrites the body's own IA / pA, which are consumed)
    const int f = F_DYN;
#pragma unroll
    for (int k = 0; k < 9; ++k) { SM(b, f + k) = Ut.a[k]; SM(b, f + 9 + k) = Ub.a[k]; }
    SM(b, f + 18) = Di.xx; SM(b, f + 19) = Di.xy; SM(b, f + 20) = Di.xz; SM(b, f + 21) = Di.yy; SM(b, f + 22) = Di.yz; SM(b, f + 23) = Di.zz;
    SM(b, f + 24) = u.x; SM(b, f + 25) = u.y; SM(b, f + 26) = u.z;
}

// acceleration of one joint and its body from the parent's spatial acceleration (physics.cu pass 3 body)
__device__ __forceinline__ f3 acc_body(const float* smem, int lane, int b, f3& aw, f3& al) {
    const int f = F_DYN;
    M3 Ut, Ub; S3 Di;
#pragma unroll
    for (int k = 0; k < 9; ++k) { Ut.a[k] = SM(b, f + k); Ub.a[k] = SM(b, f + 9 + k); }
    Di.xx = SM(b, f + 18); Di.xy = SM(b, f + 19); Di.xz = SM(b, f + 20); Di.yy = SM(b, f + 21); Di.yz = SM(b, f + 22); Di.zz = SM(b, f + 23);
    const f3 u = mk3(SM(b, f + 24), SM(b, f + 25), SM(b, f + 26));
    const f3 x = ld3(smem, lane, b, F_X);
    f3 apw = aw + ld3(smem, lane, b, F_C), apl = al + ld3(smem, lane, b, F_C + 3);
    f3 rhs = u - mtv(Ut, apw) - mtv(Ub, apl);
    f3 wdot = sv(Di, rhs);
    aw = apw + wdot; al = apl + cross3(x, wdot);
    return wdot;
}

struct SoaStep { float dt; float wgt; float max_w; bool live; bool last; };

// contact force of the part at the end-of-part velocity (physics.cu "contact force at the end-of-step velocity"),
// accumulated into the per-body sums in shared memory
__device__ __forceinline__ void contact_force(float* smem, int lane, int b, f3 aw, f3 al, f3 v0, const SoaStep& st) {
    if (!st.live) return;
    const f3 vw = ld3(smem, lane, b, F_VW), vl = ld3(smem, lane, b, F_VL);
    const float F0z = SM(b, F_CS), Sbt = SM(b, F_CS + 1), Sbn = SM(b, F_CS + 2), Stz = SM(b, F_CS + 3), Sty = SM(b, F_CS + 4),
                Stx = SM(b, F_CS + 5), Sny = SM(b, F_CS + 6), Snx = SM(b, F_CS + 7);
    f3 wn = vw + aw * st.dt, ln = v0 + vl + al * st.dt;
    float* fs = smem + SOA_FSUM + b * 3 * 32 + lane;
    fs[0] += st.wgt * (-Sbt * ln.x + (-Stz * wn.y + Sty * wn.z));
    fs[32] += st.wgt * (-Sbt * ln.y + (Stz * wn.x - Stx * wn.z));
    fs[64] += st.wgt * (F0z - Sbn * ln.z + (-Sny * wn.x + Snx * wn.y));
}

// joint acceleration from the parent's, contact force, drive torque, joint integration (physics.cu pass 3 + integrate).
// aw/al: in = parent's spatial acceleration, out = this body's.  The drive torque of the env's last live part goes to dof_force.
__device__ __forceinline__ void finish_body(float* smem, int lane, int b, f3& aw, f3& al, f3 v0, const SoaStep& st, float* dof_force_row) {
    const f3 wdot = acc_body(smem, lane, b, aw, al);
    contact_force(smem, lane, b, aw, al, v0, st);
    const M3 R = quat_to_mat(ld4(smem, lane, b, F_QW));
    if (st.live) {
        if (st.last && dof_force_row) {
            const f3 tau0 = ld3(smem, lane, b, F_TAU);
            const f3 drive = mtv(R, tau0 - wdot * SM(b, F_DD));
            float* df = dof_force_row + 3 * (b - 1);
            df[0] = drive.x; df[1] = drive.y; df[2] = drive.z;
        }
        f3 jw = ld3(smem, lane, b, F_JW) + mtv(R, wdot) * st.dt;
        float n2 = dot3(jw, jw);
        if (n2 > st.max_w * st.max_w) jw = jw * (st.max_w * rsqrtf(n2));
        f4 jq = qnormalize(qmul(ld4(smem, lane, b, F_JQ), exp_quat(jw * st.dt)));
        st3(smem, lane, b, F_JW, jw);
        st4(smem, lane, b, F_JQ, jq);
    }
}

template <bool ENVM>
__global__ void __launch_bounds__(SOA_THREADS, 1) physics_soa_kernel(PhysParams P) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // envs per CTA: 32 lanes, of which P.epb are used - chosen by the launcher so that the CTAs cover all SMs in one wave
    // (4096 envs: 147 CTAs x 28 envs instead of 128 x 32)
    const int env_raw = blockIdx.x * P.epb + lane;
    const bool env_ok = env_raw < P.N && lane < P.epb;
    const int env = env_ok ? env_raw : P.N - 1;                    // clamped: tail lanes recompute the last env, stores masked
    const EmlModelDev& Mo = *P.model;
    const ModelView<ENVM> MV{Mo, P.env_model, (size_t)P.N, (size_t)env};

    // ---------------- load: root -> smem (warp 0), joints (bodies warp, warp+8, warp+16) -> smem, drive targets -> registers
    if (warp == 0) {
        const float* r = P.root + (size_t)env * 13;
#pragma unroll
        for (int k = 0; k < 13; ++k) smem[SOA_ROOT + k * 32 + lane] = r[k];
    }
    f4 qtgt[SOA_BPW];
#pragma unroll
    for (int s = 0; s < SOA_BPW; ++s) {
        const int b = warp + SOA_WARPS * s;
        qtgt[s] = mk4(0, 0, 0, 1);
        if (b == 0) continue;
        const int d = 3 * (b - 1);
        const float4 t4 = *reinterpret_cast<const float4*>(P.jq + ((size_t)env * EML_NJ + (b - 1)) * 4);
        st4(smem, lane, b, F_JQ, mk4(t4.x, t4.y, t4.z, t4.w));
        const float2* ds = reinterpret_cast<const float2*>(P.dof + ((size_t)env * EML_ND + d) * 2);
        st3(smem, lane, b, F_JW, mk3(ds[0].y, ds[1].y, ds[2].y));
        float* tp = P.pd_target + (size_t)env * EML_ND + d;
        f3 target;
        if (P.actions) {   // pre_physics_step: pd_tar = offset + scale*action; hands and toes forced to 0 (humanoid.py:1184-1199)
            const float* ap = P.actions + (size_t)env * EML_ND + d;
            const float a0 = ap[0], a1 = ap[1], a2 = ap[2];
            const bool frozen = (b == 4 || b == 8 || b == 18 || b == 23);
            target = frozen ? mk3(0, 0, 0)
                            : mk3(Mo.pd_offset[d] + Mo.pd_scale[d] * a0, Mo.pd_offset[d + 1] + Mo.pd_scale[d + 1] * a1,
                                  Mo.pd_offset[d + 2] + Mo.pd_scale[d + 2] * a2);
            if (env_ok) {
                tp[0] = target.x; tp[1] = target.y; tp[2] = target.z;
                if (P.actions_copy) { float* ac = P.actions_copy + (size_t)env * EML_ND + d; ac[0] = a0; ac[1] = a1; ac[2] = a2; }
            }
        } else {
            target = mk3(tp[0], tp[1], tp[2]);
        }
        qtgt[s] = exp_quat(target);
    }
    // contact-force sums start at zero (24 x 3 x 32 floats: 9 per thread)
    for (int i = threadIdx.x; i < EML_NB * 3 * 32; i += SOA_THREADS) smem[SOA_FSUM + i] = 0.f;
    __syncthreads();

    const int chain = warp;                                        // warps 0..4 walk chains
    const int clen = chain < 5 ? c_chain_len[chain] : 0;
    float* const df_row = env_ok ? P.dof_force + (size_t)env * EML_ND : nullptr;
    int sub = 0, part = 0, parts = 1, kmax = 1, rb = 0;            // rb: root buffer holding the current root state
    float dt = P.dt;
    // ---- kinematics of the initial state (afterwards pass C computes the next part's kinematics right after integrating) ----
    {
        const float* root = smem + SOA_ROOT;
        const f4 q0 = mk4(root[3 * 32 + lane], root[4 * 32 + lane], root[5 * 32 + lane], root[6 * 32 + lane]);
        const f3 w0 = mk3(root[10 * 32 + lane], root[11 * 32 + lane], root[12 * 32 + lane]);
        if (chain < 5) {
            Kin k; k.q = q0; k.x = mk3(0, 0, 0); k.w = w0; k.l = mk3(0, 0, 0);
            float wm = 0.f;
            if (chain == 2) {                                      // the spine warp also owns the pelvis
                st3(smem, lane, 0, F_X, k.x); st4(smem, lane, 0, F_QW, k.q); st3(smem, lane, 0, F_VW, k.w); st3(smem, lane, 0, F_VL, k.l);
                wm = dot3(w0, w0);
            }
            const int pre = chain >= 3 ? 3 : 0;                    // arms hang off the chest: walk torso, spine, chest first
#pragma unroll 1
            for (int i = -pre; i < clen; ++i) {
                const int b = i < 0 ? 12 + i : c_chain_body[chain][i];
                f3 cw, cl;
                k = kin_step(k, MV.offset(b), ld4(smem, lane, b, F_JQ), ld3(smem, lane, b, F_JW), cw, cl);
                if (i >= 0) {
                    st3(smem, lane, b, F_X, k.x); st4(smem, lane, b, F_QW, k.q); st3(smem, lane, b, F_VW, k.w); st3(smem, lane, b, F_VL, k.l);
                    st3(smem, lane, b, F_C, cw); st3(smem, lane, b, F_C + 3, cl);
                    wm = fmaxf(wm, dot3(k.w, k.w));
                }
            }
            smem[SOA_WMAX + chain * 32 + lane] = wm;
        }
    }
    __syncthreads();
#pragma unroll 1
    while (sub < P.n_sub) {
        const float* root = smem + SOA_ROOT + rb * 13 * 32;
        const f3 p0 = mk3(root[lane], root[32 + lane], root[64 + lane]);
        const f4 q0 = mk4(root[3 * 32 + lane], root[4 * 32 + lane], root[5 * 32 + lane], root[6 * 32 + lane]);
        const f3 v0 = mk3(root[7 * 32 + lane], root[8 * 32 + lane], root[9 * 32 + lane]);
        const f3 w0 = mk3(root[10 * 32 + lane], root[11 * 32 + lane], root[12 * 32 + lane]);

        if (part == 0) {
            // adaptive refinement: split the sub-step until no body turns more than max_turn per piece (<= 8 pieces); the
            // piece count is per env (lane), the CTA loops to the largest and finished lanes stop updating their state
            parts = 1;
            if (P.max_turn > 0.f) {
                float wm = 0.f;
#pragma unroll
                for (int c = 0; c < 5; ++c) wm = fmaxf(wm, smem[SOA_WMAX + c * 32 + lane]);
                int k = (int)ceilf(sqrtf(wm) * P.dt / P.max_turn);
                parts = k < 1 ? 1 : (k > 8 ? 8 : k);
            }
            dt = P.dt / (float)parts;
            // CTA-wide maximum (note: __syncthreads_or is a LOGICAL or, it cannot carry a bit mask)
            const int wk = __reduce_max_sync(FULL, parts);
            int* s_k = reinterpret_cast<int*>(smem + SOA_KMAX);
            if (lane == 0) s_k[warp] = wk;
            __syncthreads();
            kmax = 1;
#pragma unroll
            for (int w = 0; w < SOA_WARPS; ++w) kmax = max(kmax, s_k[w]);
        }
        SoaStep st; st.dt = dt; st.wgt = 1.0f / (float)parts; st.max_w = P.max_w; st.live = part < parts;
        st.last = (sub == P.n_sub - 1) && (part == parts - 1);

        // ================= A1: per-body inertia, bias force, contacts, drive (bodies warp, warp + SOA_WARPS, ...) =================
#pragma unroll
        for (int s = 0; s < SOA_BPW; ++s) {
            const int b = warp + SOA_WARPS * s;
            const f3 x = ld3(smem, lane, b, F_X), vw = ld3(smem, lane, b, F_VW), vl = ld3(smem, lane, b, F_VL);
            const M3 R = quat_to_mat(ld4(smem, lane, b, F_QW));
            const float mass = MV.mass(b);
            Sp sp;
            {
                S3 Ib; Ib.xx = MV.inertia(b, 0); Ib.xy = MV.inertia(b, 1); Ib.xz = MV.inertia(b, 2); Ib.yy = MV.inertia(b, 3); Ib.yz = MV.inertia(b, 4); Ib.zz = MV.inertia(b, 5);
                M3 T;                                                    // T = R * Ib
#pragma unroll
                for (int r = 0; r < 3; ++r) setrow(T, r, sv(Ib, row(R, r)));
                f3 c = x + mv(R, MV.com(b));
                float c2 = dot3(c, c);
                sp.A.xx = dot3(row(T, 0), row(R, 0)) + mass * (c2 - c.x * c.x);
                sp.A.xy = dot3(row(T, 0), row(R, 1)) - mass * c.x * c.y;
                sp.A.xz = dot3(row(T, 0), row(R, 2)) - mass * c.x * c.z;
                sp.A.yy = dot3(row(T, 1), row(R, 1)) + mass * (c2 - c.y * c.y);
                sp.A.yz = dot3(row(T, 1), row(R, 2)) - mass * c.y * c.z;
                sp.A.zz = dot3(row(T, 2), row(R, 2)) + mass * (c2 - c.z * c.z);
                f3 mc = c * mass;
                sp.B.a[0] = 0; sp.B.a[1] = -mc.z; sp.B.a[2] = mc.y; sp.B.a[3] = mc.z; sp.B.a[4] = 0; sp.B.a[5] = -mc.x;
                sp.B.a[6] = -mc.y; sp.B.a[7] = mc.x; sp.B.a[8] = 0;
                sp.M.xx = sp.M.yy = sp.M.zz = mass; sp.M.xy = sp.M.xz = sp.M.yz = 0;
                f3 hn = sv(sp.A, vw) + mv(sp.B, vl);                     // I v
                f3 hf = mtv(sp.B, vw) + vl * mass;
                sp.pn = cross3(vw, hn) + cross3(vl, hf);                 // v x* (I v)
                sp.pf = cross3(vw, hf);
                f3 g = mk3(0, 0, mass * P.gz);
                sp.pn = sp.pn - cross3(c, g); sp.pf = sp.pf - g;
            }
            // ---- ground contact: implicit spring-damper folded into (A,B,M), p ----
            float F0z = 0, Sbt = 0, Sbn = 0, Stz = 0, Sty = 0, Stx = 0, Sny = 0, Snx = 0;
            {
                const int gt = Mo.geom_type[b];
                const f3 ga = MV.geom_a(b);
                const f3 gb = MV.geom_b(b);
                const float drop = gt == 2 ? 0.f : MV.geom_r(b);
                const int np = gt == 0 ? 1 : (gt == 1 ? 2 : 8);
                const float bn = P.kn * dt + P.cn;
                // no contact point of this body can be below the highest terrain sample: skip the loop when that holds for
                // every env of the warp (same result: each point would find gap >= 0)
                const bool reach = p0.z + x.z - MV.geom_bound(b) < P.hf_max;
                const int npw = __any_sync(FULL, reach) ? np : 0;          // evaluated once, while the warp is converged
#pragma unroll 1
                for (int k = 0; k < npw; ++k) {
                    f3 pb;
                    if (gt == 2) pb = mk3(ga.x + ((k & 1) ? gb.x : -gb.x), ga.y + ((k & 2) ? gb.y : -gb.y), ga.z + ((k & 4) ? gb.z : -gb.z));
                    else pb = (k == 0) ? ga : gb;
                    f3 r = x + mv(R, pb);
                    r.z -= drop;
                    float gap = p0.z + r.z - ground_height(P, p0.x + r.x, p0.y + r.y);
                    if (gap >= 0.f) continue;
                    f3 vp = v0 + vl + cross3(vw, r);                    // absolute velocity of the contact point
                    float fn = -P.kn * gap - bn * vp.z;
                    if (fn <= 0.f) continue;                             // separating: no adhesion
                    float vt = sqrtf(vp.x * vp.x + vp.y * vp.y);
                    float bt = P.ct;
                    if (bt * vt > P.mu * fn) bt = P.mu * fn / vt;        // regularised Coulomb cone
                    float f0z = -P.kn * gap;
                    float dbt = dt * bt, dbn = dt * bn;
                    sp.A.xx += dbt * r.z * r.z + dbn * r.y * r.y; sp.A.yy += dbt * r.z * r.z + dbn * r.x * r.x;
                    sp.A.zz += dbt * (r.x * r.x + r.y * r.y);
                    sp.A.xy -= dbn * r.x * r.y; sp.A.xz -= dbt * r.x * r.z; sp.A.yz -= dbt * r.y * r.z;
                    sp.B.a[1] -= r.z * dbt; sp.B.a[2] += r.y * dbn; sp.B.a[3] += r.z * dbt; sp.B.a[5] -= r.x * dbn;
                    sp.B.a[6] -= r.y * dbt; sp.B.a[7] += r.x * dbt;
                    sp.M.xx += dbt; sp.M.yy += dbt; sp.M.zz += dbn;
                    f3 w = mk3(bt * vp.x, bt * vp.y, bn * vp.z - f0z);
                    sp.pn = sp.pn + cross3(r, w); sp.pf = sp.pf + w;
                    F0z += f0z; Sbt += bt; Sbn += bn; Stz += bt * r.z; Sty += bt * r.y; Stx += bt * r.x;
                    Sny += bn * r.y; Snx += bn * r.x;
                }
            }
            SM(b, F_CS) = F0z; SM(b, F_CS + 1) = Sbt; SM(b, F_CS + 2) = Sbn; SM(b, F_CS + 3) = Stz; SM(b, F_CS + 4) = Sty;
            SM(b, F_CS + 5) = Stx; SM(b, F_CS + 6) = Sny; SM(b, F_CS + 7) = Snx;
            // ---- implicit PD drive ----
            if (b > 0) {
                const float kp = MV.kp(b), kd = MV.kd(b);
                const f4 jq = ld4(smem, lane, b, F_JQ); const f3 jw = ld3(smem, lane, b, F_JW);
                f3 e = log_quat(qmul(qconj(jq), qtgt[s]));              // position error on SO(3), child frame
                float tm = fmaxf(fmaxf(fabsf(kp * e.x - kd * jw.x), fabsf(kp * e.y - kd * jw.y)), fabsf(kp * e.z - kd * jw.z));
                float sat = (P.max_effort > 0.f && tm > P.max_effort) ? P.max_effort / tm : 1.0f;   // effort limit
                float kk = kd + kp * dt;
                f3 t0 = mk3(sat * (kp * e.x - kk * jw.x), sat * (kp * e.y - kk * jw.y), sat * (kp * e.z - kk * jw.z));
                st3(smem, lane, b, F_TAU, mv(R, t0));
                SM(b, F_DD) = sat * dt * (kd + kp * dt);
            }
            st_sp(smem, lane, b, sp);
        }
        __syncthreads();

        // ================= B1: articulated inertias inside each chain, leaf -> chain root =================
        Sp carry;
        if (chain < 5) {
            const int stop = chain == 2 ? 3 : 0;                       // spine warp: head, neck now; chest.. after the arms
SOA_TREE_UNROLL
            for (int i = clen - 1; i >= stop; --i) {
                const int b = c_chain_body[chain][i];
                Sp sp; ld_sp(smem, lane, b, sp);
                if (i < clen - 1) add_sp(sp, carry);
                aba_body(smem, lane, b, sp, MV.arm(b));
                carry = sp;
            }
            if (chain != 2) st_xchg(smem, lane, chain < 2 ? chain : chain - 1, carry);   // slots: 0,1 legs; 2,3 arms
        }
        __syncthreads();

        // ================= B2 (spine warp): chest, spine, torso; pelvis and the 6x6 root solve; accelerations of torso..chest ==========
        if (chain == 2) {
SOA_TREE_UNROLL
            for (int b = 11; b >= 9; --b) {
                Sp sp; ld_sp(smem, lane, b, sp);
                add_sp(sp, carry);
                if (b == 11) { add_xchg(smem, lane, 2, sp); add_xchg(smem, lane, 3, sp); }
                aba_body(smem, lane, b, sp, MV.arm(b));
                carry = sp;
            }
            Sp sp; ld_sp(smem, lane, 0, sp);
            add_sp(sp, carry); add_xchg(smem, lane, 0, sp); add_xchg(smem, lane, 1, sp);
            // floating base: [[A,B],[B^T,M]] [alpha; l] = -[pn; pf]
            S3 Mi = inv_s3(sp.M);
            M3 T;                                                        // T = B Minv
#pragma unroll
            for (int r = 0; r < 3; ++r) setrow(T, r, sv(Mi, row(sp.B, r)));
            S3 Sc;                                                       // A - T B^T
            Sc.xx = sp.A.xx - dot3(row(T, 0), row(sp.B, 0)); Sc.xy = sp.A.xy - dot3(row(T, 0), row(sp.B, 1)); Sc.xz = sp.A.xz - dot3(row(T, 0), row(sp.B, 2));
            Sc.yy = sp.A.yy - dot3(row(T, 1), row(sp.B, 1)); Sc.yz = sp.A.yz - dot3(row(T, 1), row(sp.B, 2)); Sc.zz = sp.A.zz - dot3(row(T, 2), row(sp.B, 2));
            f3 rhs = mv(T, sp.pf) - sp.pn;
            f3 aw = sv(inv_s3(Sc), rhs);
            f3 al = sv(Mi, mk3(0, 0, 0) - sp.pf - mtv(sp.B, aw));
            float* acc = smem + SOA_ACC + lane;
            acc[0] = aw.x; acc[32] = aw.y; acc[64] = aw.z; acc[96] = al.x; acc[128] = al.y; acc[160] = al.z;
            contact_force(smem, lane, 0, aw, al, v0, st);
            // root integration into the other root buffer (the current one is still read by the other warps in pass C)
            {
                float* nr = smem + SOA_ROOT + (rb ^ 1) * 13 * 32 + lane;
                f3 np0 = p0, nv0 = v0, nw0 = w0; f4 nq0 = q0;
                if (st.live) {
                    f3 wn = w0 + aw * dt, vO = al * dt;                  // vO: in-frame velocity gained by the pelvis point
                    float n2 = dot3(wn, wn);
                    if (n2 > P.max_w * P.max_w) wn = wn * (P.max_w * rsqrtf(n2));    // maxAngularVelocity (humanoid.py:685-688)
                    nq0 = qnormalize(qmul(exp_quat(wn * dt), q0));
                    f3 vn = v0 + vO;
                    np0 = p0 + vn * dt;
                    nv0 = vn + cross3(wn, vO * dt);                      // re-reference to the moved origin (second order)
                    nw0 = wn;
                }
                nr[0] = np0.x; nr[32] = np0.y; nr[64] = np0.z; nr[96] = nq0.x; nr[128] = nq0.y; nr[160] = nq0.z; nr[192] = nq0.w;
                nr[224] = nv0.x; nr[256] = nv0.y; nr[288] = nv0.z; nr[320] = nw0.x; nr[352] = nw0.y; nr[384] = nw0.z;
            }
            // torso, spine, chest: accelerations, contact force, integration (the arms and the neck start from the chest's)
            Kin k;                                                         // kinematics of the COMING part, from the new state
            {
                const float* nr = smem + SOA_ROOT + (rb ^ 1) * 13 * 32 + lane;
                k.q = mk4(nr[96], nr[128], nr[160], nr[192]); k.x = mk3(0, 0, 0);
                k.w = mk3(nr[320], nr[352], nr[384]); k.l = mk3(0, 0, 0);
            }
            float wm = dot3(k.w, k.w);
            st3(smem, lane, 0, F_X, k.x); st4(smem, lane, 0, F_QW, k.q); st3(smem, lane, 0, F_VW, k.w); st3(smem, lane, 0, F_VL, k.l);
SOA_TREE_UNROLL
            for (int b = 9; b <= 11; ++b) {
                finish_body(smem, lane, b, aw, al, v0, st, df_row);
                f3 cw, cl;
                k = kin_step(k, MV.offset(b), ld4(smem, lane, b, F_JQ), ld3(smem, lane, b, F_JW), cw, cl);
                store_kin(smem, lane, b, k, cw, cl);
                wm = fmaxf(wm, dot3(k.w, k.w));
            }
            put_kin(smem, lane, k);                                        // the arms and the neck continue from the chest
            smem[SOA_WMAX + 2 * 32 + lane] = wm;                           // (the spine warp adds neck and head in pass C)
            acc[6 * 32] = aw.x; acc[7 * 32] = aw.y; acc[8 * 32] = aw.z; acc[9 * 32] = al.x; acc[10 * 32] = al.y; acc[11 * 32] = al.z;
        }
        __syncthreads();

        // ================= C: accelerations root -> leaf inside each chain, contact force, drive torque, integration =================
        if (chain < 5) {
            const float* acc = smem + SOA_ACC + (chain >= 2 ? 6 * 32 : 0) + lane;   // legs start at the pelvis, the rest at the chest
            f3 aw = mk3(acc[0], acc[32], acc[64]), al = mk3(acc[96], acc[128], acc[160]);
            Kin k;
            float wm = 0.f;
            if (chain >= 2) {
                k = get_kin(smem, lane);
                if (chain == 2) wm = smem[SOA_WMAX + 2 * 32 + lane];
            } else {
                const float* nr = smem + SOA_ROOT + (rb ^ 1) * 13 * 32 + lane;
                k.q = mk4(nr[96], nr[128], nr[160], nr[192]); k.x = mk3(0, 0, 0);
                k.w = mk3(nr[320], nr[352], nr[384]); k.l = mk3(0, 0, 0);
            }
SOA_TREE_UNROLL
            for (int i = chain == 2 ? 3 : 0; i < clen; ++i) {
                const int b = c_chain_body[chain][i];
                finish_body(smem, lane, b, aw, al, v0, st, df_row);
                f3 cw, cl;                                                 // kinematics of the coming part
                k = kin_step(k, MV.offset(b), ld4(smem, lane, b, F_JQ), ld3(smem, lane, b, F_JW), cw, cl);
                store_kin(smem, lane, b, k, cw, cl);
                wm = fmaxf(wm, dot3(k.w, k.w));
            }
            smem[SOA_WMAX + chain * 32 + lane] = wm;
        }
        rb ^= 1;
        if (++part == kmax) { part = 0; ++sub; }
        __syncthreads();
    }

    // ================= refresh: forward kinematics -> rigid-body state, DOF state, contact forces =================
    // Lane = env means that a direct store of a body's 13 state floats is 32 scattered 4-byte writes per instruction (env rows are
    // 1248 B apart): the refresh took ~20 % of the kernel.  The walk therefore leaves its results in shared memory - body b's 26
    // output floats in the dead fields of b's own block, pitch 33 so that the transposed reads below are conflict free - and all
    // 12 warps then copy each env's contiguous global rows (rb 312, contact 72, dof 138, joint quaternions 92, root 13 floats)
    // with coalesced stores.  Same values, same arithmetic.
    constexpr int ST_OFF = 7 * 32, ST_PITCH = 33;         // after F_JQ / F_JW (still read by the walk); k-th staged float of a body
    constexpr int BLK = F_PER_BODY * 32;
    static_assert(ST_OFF + 29 * ST_PITCH + 32 <= BLK, "staging fits the body block");
#define STG(b, k) smem[(b) * BLK + ST_OFF + (k) * ST_PITCH + lane]
    if (chain < 5) {
        const float* root = smem + SOA_ROOT + rb * 13 * 32;
        const f3 p0 = mk3(root[lane], root[32 + lane], root[64 + lane]);
        const f4 q0 = mk4(root[3 * 32 + lane], root[4 * 32 + lane], root[5 * 32 + lane], root[6 * 32 + lane]);
        const f3 v0 = mk3(root[7 * 32 + lane], root[8 * 32 + lane], root[9 * 32 + lane]);
        const f3 w0 = mk3(root[10 * 32 + lane], root[11 * 32 + lane], root[12 * 32 + lane]);
        const float inv = P.n_sub > 0 ? 1.0f / (float)P.n_sub : 0.f;   // mean force over the sub-steps
        f4 qw = q0; f3 x = p0, wv = w0, lv = v0;
        auto put_body = [&](int b) {                                   // k 0..12 rigid-body state, 13..15 contact force
            STG(b, 0) = x.x; STG(b, 1) = x.y; STG(b, 2) = x.z; STG(b, 3) = qw.x; STG(b, 4) = qw.y; STG(b, 5) = qw.z; STG(b, 6) = qw.w;
            STG(b, 7) = lv.x; STG(b, 8) = lv.y; STG(b, 9) = lv.z; STG(b, 10) = wv.x; STG(b, 11) = wv.y; STG(b, 12) = wv.z;
            const float* fs = smem + SOA_FSUM + b * 3 * 32 + lane;
            STG(b, 13) = fs[0] * inv; STG(b, 14) = fs[32] * inv; STG(b, 15) = fs[64] * inv;
        };
        if (chain == 2) {
            put_body(0);                                               // k 16..28 of body 0: the root state row
            STG(0, 16) = p0.x; STG(0, 17) = p0.y; STG(0, 18) = p0.z; STG(0, 19) = q0.x; STG(0, 20) = q0.y; STG(0, 21) = q0.z; STG(0, 22) = q0.w;
            STG(0, 23) = v0.x; STG(0, 24) = v0.y; STG(0, 25) = v0.z; STG(0, 26) = w0.x; STG(0, 27) = w0.y; STG(0, 28) = w0.z;
        }
        const int pre = chain >= 3 ? 3 : 0;
#pragma unroll 1
        for (int i = -pre; i < clen; ++i) {
            const int b = i < 0 ? 12 + i : c_chain_body[chain][i];
            const f4 jq = ld4(smem, lane, b, F_JQ); const f3 jw = ld3(smem, lane, b, F_JW);
            const f3 t = qrot(qw, MV.offset(b));
            lv = lv + cross3(wv, t);
            x = x + t;
            qw = qmul(qw, jq);
            wv = wv + qrot(qw, jw);
            if (i < 0) continue;
            put_body(b);
            const f3 e = log_quat(jq);                                 // k 16..18 exp-map position, 19..21 rate, 22..25 joint quaternion
            STG(b, 16) = e.x; STG(b, 17) = e.y; STG(b, 18) = e.z; STG(b, 19) = jw.x; STG(b, 20) = jw.y; STG(b, 21) = jw.z;
            STG(b, 22) = jq.x; STG(b, 23) = jq.y; STG(b, 24) = jq.z; STG(b, 25) = jq.w;
        }
    }
    __syncthreads();
    {
        const int env0 = blockIdx.x * P.epb;
        for (int e = warp; e < P.epb && env0 + e < P.N; e += SOA_WARPS) {
            const size_t g = (size_t)(env0 + e);
            const float* st = smem + ST_OFF + e;                       // + body * BLK + k * ST_PITCH
            float* o = P.rb + g * (EML_NB * 13);
            for (int r = lane; r < EML_NB * 13; r += 32) { const int b = r / 13, k = r - b * 13; o[r] = st[b * BLK + k * ST_PITCH]; }
            o = P.contact + g * (EML_NB * 3);
            for (int r = lane; r < EML_NB * 3; r += 32) { const int b = r / 3, k = r - b * 3; o[r] = st[b * BLK + (13 + k) * ST_PITCH]; }
            o = P.dof + g * (EML_ND * 2);
            for (int r = lane; r < EML_ND * 2; r += 32) {              // (position, velocity) pairs of the 69 DOFs
                const int d = r >> 1, b = d / 3 + 1, k = 16 + 3 * (r & 1) + (d - (b - 1) * 3);
                o[r] = st[b * BLK + k * ST_PITCH];
            }
            o = P.jq + g * (EML_NJ * 4);
            for (int r = lane; r < EML_NJ * 4; r += 32) { const int b = (r >> 2) + 1; o[r] = st[b * BLK + (22 + (r & 3)) * ST_PITCH]; }
            if (lane < 13) P.root[g * 13 + lane] = st[(16 + lane) * ST_PITCH];
        }
    }
#undef STG
}

const EmlModelDev* eml_model_dev();
void eml_fill_phys_params(emloco_sim* s, PhysParams& P);

cudaError_t eml_launch_physics_soa(emloco_sim* s, const float* d_actions, int n_substeps, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(physics_soa_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOA_SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(physics_soa_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOA_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    PhysParams P; eml_fill_phys_params(s, P);
    P.actions = d_actions; P.actions_copy = d_actions ? s->actions : nullptr; P.n_sub = n_substeps;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    // one CTA per SM is resident (214 KB of shared memory): spread the envs over as many SMs as possible, <= 32 per CTA
    int epb = (s->N + sms - 1) / sms;
    epb = epb < 1 ? 1 : (epb > 32 ? 32 : epb);
    if ((s->N + epb - 1) / epb > sms && epb < 32) epb = 32;      // more than one wave anyway: use full warps
    P.epb = epb;
    const int blocks = (s->N + epb - 1) / epb;
    if (P.env_model) physics_soa_kernel<true><<<blocks, SOA_THREADS, SOA_SMEM_BYTES, st>>>(P);        // per-env body models (row f3)
    else physics_soa_kernel<false><<<blocks, SOA_THREADS, SOA_SMEM_BYTES, st>>>(P);
    return cudaGetLastError();
}
