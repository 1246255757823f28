// Physics step kernel: replaces gym.simulate x controlFrequencyInv (reference
// pacer/pacer/env/tasks/base_task.py:792-797) and the PD-target set-up of pre_physics_step
// (humanoid.py:1184-1209, _action_to_pd_targets :1281-1283) - SURVEY 8a rows a1+a2.
//
// ONE WARP PER ENV, ONE LANE PER RIGID BODY (24 of 32 lanes).  The whole env step (4 sub-steps of
// 1/120 s) runs inside one launch with the reduced state in registers; HBM is touched once on the
// way in (actions, root state, joint rotations/rates: ~1.2 KB) and once on the way out (reduced state,
// rigid-body state, contact forces, DOF forces: ~2.8 KB).
//
// Algorithm per sub-step (DESIGN.md "Physics"; fp64 restatement in oracle/physics_oracle.c):
//   pass 1  root->leaves   kinematics in world-aligned axes about the pelvis (warp shuffles from the parent lane)
//   local   rigid-body spatial inertia, bias force, gravity; ground contact as an implicit spring-damper
//           folded into the body's spatial inertia (regularised Coulomb friction); implicit PD drive
//   pass 2  leaves->root   articulated-body inertias; children are summed into the parent through shuffles
//   root    6x6 floating-base solve (Schur complement) on lane 0
//   pass 3  root->leaves   joint and body accelerations
//   integrate (semi-implicit Euler), joints as unit quaternions, DOF position = exponential map
// The 3-hinge MJCF joints are spherical joints with exp-map coordinates, as the reference treats them
// (humanoid.py:1359-1360, utils/motion_lib_smpl.py:611-614).
#include "sim.h"
#include "physics_math.cuh"

#define PH_WARPS 4


template <bool ENVM>
__global__ void __launch_bounds__(PH_WARPS * 32) physics_kernel(PhysParams P) {
    __shared__ __align__(16) float s_rb[PH_WARPS][EML_NB * 13];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * PH_WARPS + warp;
    if (slot >= P.N) return;                       // warp-uniform
    const int env = P.env_ids ? P.env_ids[slot] : slot;
    if (P.reset_mask && P.reset_mask[env] == 0) return;   // warp-uniform
    const bool body = lane < EML_NB;
    const int b = body ? lane : 0;
    const EmlModelDev& Mo = *P.model;
    const ModelView<ENVM> MV{Mo, P.env_model, (size_t)P.N, (size_t)env};

    // ---- per-lane model constants ----
    const int parent = body ? Mo.parent[b] : 0;
    const int level = body ? Mo.level[b] : 99;
    const int par = parent < 0 ? 0 : parent;
    int ch0 = body ? Mo.child[b][0] : -1, ch1 = body ? Mo.child[b][1] : -1, ch2 = body ? Mo.child[b][2] : -1;
    const f3 offset = MV.offset(b);
    const float mass = MV.mass(b);
    const f3 com = MV.com(b);
    const float kp = MV.kp(b), kd = MV.kd(b), arm = MV.arm(b);
    const int max_level = Mo.max_level;
    const bool joint = body && b > 0;

    // ---- load reduced state ----
    f3 p0 = mk3(0, 0, 0), v0 = p0, w0 = p0; f4 q0 = mk4(0, 0, 0, 1);
    {
        const float* r = (P.init_root ? P.init_root : P.root) + (size_t)env * 13;   // every lane reads the root (broadcast load)
        p0 = mk3(r[0], r[1], r[2]); q0 = mk4(r[3], r[4], r[5], r[6]);
        v0 = mk3(r[7], r[8], r[9]); w0 = mk3(r[10], r[11], r[12]);
    }
    f4 jq = mk4(0, 0, 0, 1); f3 jw = mk3(0, 0, 0), target = mk3(0, 0, 0);
    if (joint) {
        const int d = 3 * (b - 1);
        const float2* ds = reinterpret_cast<const float2*>((P.init_dof ? P.init_dof : P.dof) + ((size_t)env * EML_ND + d) * 2);
        float2 d0 = ds[0], d1 = ds[1], d2 = ds[2];
        jw = mk3(d0.y, d1.y, d2.y);
        if (P.fk_only) {
            jq = exp_quat(mk3(d0.x, d1.x, d2.x));           // caller wrote exp-map DOF positions
        } else {
            float4 t4 = *reinterpret_cast<const float4*>(P.jq + ((size_t)env * EML_NJ + (b - 1)) * 4);
            jq = mk4(t4.x, t4.y, t4.z, t4.w);
            float* tp = P.pd_target + (size_t)env * EML_ND + d;
            if (P.actions) {
                // pre_physics_step: pd_tar = offset + scale*action; hands and toes forced to 0 (humanoid.py:1184-1199)
                const float* ap = P.actions + (size_t)env * EML_ND + d;
                float a0 = ap[0], a1 = ap[1], a2 = ap[2];
                bool frozen = (b == 4 || b == 8 || b == 18 || b == 23);
                target = frozen ? mk3(0, 0, 0)
                                : mk3(Mo.pd_offset[d] + Mo.pd_scale[d] * a0, Mo.pd_offset[d + 1] + Mo.pd_scale[d + 1] * a1,
                                      Mo.pd_offset[d + 2] + Mo.pd_scale[d + 2] * a2);
                tp[0] = target.x; tp[1] = target.y; tp[2] = target.z;
                if (P.actions_copy) { float* ac = P.actions_copy + (size_t)env * EML_ND + d; ac[0] = a0; ac[1] = a1; ac[2] = a2; }
            } else {
                target = mk3(tp[0], tp[1], tp[2]);
            }
        }
    }
    f3 fsum = mk3(0, 0, 0);                               // contact force accumulated over the sub-steps
    f3 drive = mk3(0, 0, 0);                              // last drive torque, child frame
    f4 qtgt = mk4(0, 0, 0, 1);
    if (joint && !P.fk_only) qtgt = exp_quat(target);     // drive target as a rotation

    const int n_sub = P.fk_only ? 0 : P.n_sub;
    // Adaptive refinement: the explicit velocity-product terms gain energy like (|w| dt)^2, so a nominal sub-step is split
    // into `parts` equal pieces until no body turns more than max_turn radians per piece (parts <= 8).  One warp = one env,
    // so the trip count is warp-uniform.
    int sub = 0, part = 0, parts = 1;
    float dt = P.dt;
#pragma unroll 1
    while (sub < n_sub) {
        // ================= pass 1: kinematics (in the inertial frame translating with the pelvis velocity v0) =================
        f4 qw = q0; f3 x = mk3(0, 0, 0); f3 vw = w0, vl = mk3(0, 0, 0);   // lane 0 (and template for the others)
        f3 cw_ = mk3(0, 0, 0), cl_ = mk3(0, 0, 0);                  // velocity-product acceleration c_i
        f3 ww = mk3(0, 0, 0);
#pragma unroll 1
        for (int L = 1; L <= max_level; ++L) {
            f4 qp = shfl4(qw, par); f3 xp = shfl3(x, par); f3 wp = shfl3(vw, par); f3 lp = shfl3(vl, par);
            if (level == L) {
                x = xp + qrot(qp, offset);
                qw = qmul(qp, jq);
                ww = qrot(qw, jw);                                   // joint rate in world axes
                f3 jl = cross3(x, ww);
                vw = wp + ww; vl = lp + jl;
                cw_ = cross3(wp, ww);                                // c = v_parent x vJ
                cl_ = cross3(wp, jl) + cross3(lp, ww);
            }
        }
        if (part == 0) {
            parts = 1; dt = P.dt;
            if (P.max_turn > 0.f) {
                float wm = body ? sqrtf(dot3(vw, vw)) : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(FULL, wm, o));
                int k = (int)ceilf(wm * P.dt / P.max_turn);
                parts = k < 1 ? 1 : (k > 8 ? 8 : k);
                dt = P.dt / (float)parts;
            }
        }
        const float dd_pd_full = dt * (kd + kp * dt);         // implicit PD: extra joint-space inertia
        const M3 R = quat_to_mat(qw);
        // ================= rigid-body inertia about O, bias force, gravity =================
        S3 A, Mm; M3 Bm; f3 pn, pf;
        {
            S3 Ib; Ib.xx = MV.inertia(b, 0); Ib.xy = MV.inertia(b, 1); Ib.xz = MV.inertia(b, 2); Ib.yy = MV.inertia(b, 3); Ib.yz = MV.inertia(b, 4); Ib.zz = MV.inertia(b, 5);
            M3 T;                                                    // T = R * Ib
#pragma unroll
            for (int r = 0; r < 3; ++r) setrow(T, r, sv(Ib, row(R, r)));
            f3 c = x + mv(R, com);
            float c2 = dot3(c, c);
            A.xx = dot3(row(T, 0), row(R, 0)) + mass * (c2 - c.x * c.x);
            A.xy = dot3(row(T, 0), row(R, 1)) - mass * c.x * c.y;
            A.xz = dot3(row(T, 0), row(R, 2)) - mass * c.x * c.z;
            A.yy = dot3(row(T, 1), row(R, 1)) + mass * (c2 - c.y * c.y);
            A.yz = dot3(row(T, 1), row(R, 2)) - mass * c.y * c.z;
            A.zz = dot3(row(T, 2), row(R, 2)) + mass * (c2 - c.z * c.z);
            f3 mc = c * mass;
            Bm.a[0] = 0; Bm.a[1] = -mc.z; Bm.a[2] = mc.y; Bm.a[3] = mc.z; Bm.a[4] = 0; Bm.a[5] = -mc.x;
            Bm.a[6] = -mc.y; Bm.a[7] = mc.x; Bm.a[8] = 0;
            Mm.xx = Mm.yy = Mm.zz = mass; Mm.xy = Mm.xz = Mm.yz = 0;
            f3 hn = sv(A, vw) + mv(Bm, vl);                          // I v
            f3 hf = mtv(Bm, vw) + vl * mass;
            pn = cross3(vw, hn) + cross3(vl, hf);                    // v x* (I v)
            pf = cross3(vw, hf);
            f3 g = mk3(0, 0, mass * P.gz);
            pn = pn - cross3(c, g); pf = pf - g;
            if (!body) { A.xx = A.yy = A.zz = 1.f; Mm.xx = Mm.yy = Mm.zz = 1.f; }   // keep idle lanes finite
        }
        // ================= ground contact: implicit spring-damper folded into (A,B,M), p =================
        float F0z = 0, Sbt = 0, Sbn = 0, Stz = 0, Sty = 0, Stx = 0, Sny = 0, Snx = 0;
        if (body) {
            const int gt = Mo.geom_type[b];
            const f3 ga = MV.geom_a(b);
            const f3 gb = MV.geom_b(b);
            const float drop = gt == 2 ? 0.f : MV.geom_r(b);
            const int np = gt == 0 ? 1 : (gt == 1 ? 2 : 8);
            const float bn = P.kn * dt + P.cn;
#pragma unroll 1
            for (int k = 0; k < np; ++k) {
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
                A.xx += dbt * r.z * r.z + dbn * r.y * r.y; A.yy += dbt * r.z * r.z + dbn * r.x * r.x;
                A.zz += dbt * (r.x * r.x + r.y * r.y);
                A.xy -= dbn * r.x * r.y; A.xz -= dbt * r.x * r.z; A.yz -= dbt * r.y * r.z;
                Bm.a[1] -= r.z * dbt; Bm.a[2] += r.y * dbn; Bm.a[3] += r.z * dbt; Bm.a[5] -= r.x * dbn;
                Bm.a[6] -= r.y * dbt; Bm.a[7] += r.x * dbt;
                Mm.xx += dbt; Mm.yy += dbt; Mm.zz += dbn;
                f3 w = mk3(bt * vp.x, bt * vp.y, bn * vp.z - f0z);
                pn = pn + cross3(r, w); pf = pf + w;
                F0z += f0z; Sbt += bt; Sbn += bn; Stz += bt * r.z; Sty += bt * r.y; Stx += bt * r.x;
                Sny += bn * r.y; Snx += bn * r.x;
            }
        }
        // ================= implicit PD drive =================
        f3 tau0 = mk3(0, 0, 0);
        float dd_pd = dd_pd_full;
        if (joint) {
            // position error on SO(3): e = log(q^-1 * exp(target)), child frame; = target - log q to first order, defined
            // for every target (component-wise targets of norm > pi are legal actions), zero exactly at the target
            f3 e = log_quat(qmul(qconj(jq), qtgt));
            // effort limit (MJCF motor gear -> DOF effort): scale the whole drive, implicit part included
            float tm = fmaxf(fmaxf(fabsf(kp * e.x - kd * jw.x), fabsf(kp * e.y - kd * jw.y)), fabsf(kp * e.z - kd * jw.z));
            float sat = (P.max_effort > 0.f && tm > P.max_effort) ? P.max_effort / tm : 1.0f;
            float kk = kd + kp * dt;
            f3 t0 = mk3(sat * (kp * e.x - kk * jw.x), sat * (kp * e.y - kk * jw.y), sat * (kp * e.z - kk * jw.z));
            tau0 = mv(R, t0);
            dd_pd = sat * dd_pd_full;
        }
        // ================= pass 2: articulated inertias, leaves -> root =================
        M3 Ut, Ub; S3 Di; f3 u = mk3(0, 0, 0);
#pragma unroll
        for (int k = 0; k < 9; ++k) { Ut.a[k] = 0; Ub.a[k] = 0; }
        Di.xx = Di.yy = Di.zz = 1.f; Di.xy = Di.xz = Di.yz = 0;
#pragma unroll 1
        for (int L = max_level; L >= 1; --L) {
            if (level == L) {
                // U = IA S with S = [1; [x]x]:  U_top = A + rows(B) x x,  U_bot = B^T + rows(M) x x
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    setrow(Ut, r, srow(A, r) + cross3(row(Bm, r), x));
                    setrow(Ub, r, col(Bm, r) + cross3(srow(Mm, r), x));
                }
                // D = S^T U = U_top - x x cols(U_bot), + (armature + implicit PD) on the diagonal
                S3 D;
                f3 d0 = col(Ut, 0) - cross3(x, col(Ub, 0));
                f3 d1 = col(Ut, 1) - cross3(x, col(Ub, 1));
                f3 d2 = col(Ut, 2) - cross3(x, col(Ub, 2));
                float dd = arm + dd_pd;
                D.xx = d0.x + dd; D.xy = 0.5f * (d0.y + d1.x); D.xz = 0.5f * (d0.z + d2.x);
                D.yy = d1.y + dd; D.yz = 0.5f * (d1.z + d2.y); D.zz = d2.z + dd;
                Di = inv_s3(D);
                u = tau0 - pn + cross3(x, pf);
                // W = U Dinv
                M3 Wt, Wb;
#pragma unroll
                for (int r = 0; r < 3; ++r) { setrow(Wt, r, sv(Di, row(Ut, r))); setrow(Wb, r, sv(Di, row(Ub, r))); }
                // Ia = IA - W U^T
                A.xx -= dot3(row(Wt, 0), row(Ut, 0)); A.xy -= dot3(row(Wt, 0), row(Ut, 1)); A.xz -= dot3(row(Wt, 0), row(Ut, 2));
                A.yy -= dot3(row(Wt, 1), row(Ut, 1)); A.yz -= dot3(row(Wt, 1), row(Ut, 2)); A.zz -= dot3(row(Wt, 2), row(Ut, 2));
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c2 = 0; c2 < 3; ++c2) Bm.a[3 * r + c2] -= dot3(row(Wt, r), row(Ub, c2));
                Mm.xx -= dot3(row(Wb, 0), row(Ub, 0)); Mm.xy -= dot3(row(Wb, 0), row(Ub, 1)); Mm.xz -= dot3(row(Wb, 0), row(Ub, 2));
                Mm.yy -= dot3(row(Wb, 1), row(Ub, 1)); Mm.yz -= dot3(row(Wb, 1), row(Ub, 2)); Mm.zz -= dot3(row(Wb, 2), row(Ub, 2));
                // pa = pA + Ia c + W u
                pn = pn + sv(A, cw_) + mv(Bm, cl_) + mv(Wt, u);
                pf = pf + mtv(Bm, cw_) + sv(Mm, cl_) + mv(Wb, u);
            }
            // children (level L) -> parents (level L-1); only bodies 0 and 11 have more than one child
            const int nslots = (L == 1 || L == 4) ? 3 : 1;
            for (int s = 0; s < nslots; ++s) {
                int ch = s == 0 ? ch0 : (s == 1 ? ch1 : ch2);
                bool take = (level == L - 1) && ch >= 0;
                int src = ch >= 0 ? ch : lane;
                float r;
#define ACC(field) r = __shfl_sync(FULL, field, src); if (take) field += r;
                ACC(A.xx) ACC(A.xy) ACC(A.xz) ACC(A.yy) ACC(A.yz) ACC(A.zz)
                ACC(Bm.a[0]) ACC(Bm.a[1]) ACC(Bm.a[2]) ACC(Bm.a[3]) ACC(Bm.a[4]) ACC(Bm.a[5]) ACC(Bm.a[6]) ACC(Bm.a[7]) ACC(Bm.a[8])
                ACC(Mm.xx) ACC(Mm.xy) ACC(Mm.xz) ACC(Mm.yy) ACC(Mm.yz) ACC(Mm.zz)
                ACC(pn.x) ACC(pn.y) ACC(pn.z) ACC(pf.x) ACC(pf.y) ACC(pf.z)
#undef ACC
            }
        }
        // ================= floating base: [[A,B],[B^T,M]] [alpha; l] = -[pn; pf] (lane 0) =================
        f3 aw = mk3(0, 0, 0), al = mk3(0, 0, 0);
        if (lane == 0) {
            S3 Mi = inv_s3(Mm);
            M3 T;                                                    // T = B Minv
#pragma unroll
            for (int r = 0; r < 3; ++r) setrow(T, r, sv(Mi, row(Bm, r)));
            S3 Sc;                                                   // A - T B^T
            Sc.xx = A.xx - dot3(row(T, 0), row(Bm, 0)); Sc.xy = A.xy - dot3(row(T, 0), row(Bm, 1)); Sc.xz = A.xz - dot3(row(T, 0), row(Bm, 2));
            Sc.yy = A.yy - dot3(row(T, 1), row(Bm, 1)); Sc.yz = A.yz - dot3(row(T, 1), row(Bm, 2)); Sc.zz = A.zz - dot3(row(T, 2), row(Bm, 2));
            f3 rhs = mv(T, pf) - pn;
            aw = sv(inv_s3(Sc), rhs);
            al = sv(Mi, mk3(0, 0, 0) - pf - mtv(Bm, aw));
        }
        // ================= pass 3: accelerations, root -> leaves =================
        f3 wdot = mk3(0, 0, 0);
#pragma unroll 1
        for (int L = 1; L <= max_level; ++L) {
            f3 apw = shfl3(aw, par), apl = shfl3(al, par);
            if (level == L) {
                apw = apw + cw_; apl = apl + cl_;
                f3 rhs = u - mtv(Ut, apw) - mtv(Ub, apl);
                wdot = sv(Di, rhs);
                aw = apw + wdot; al = apl + cross3(x, wdot);
            }
        }
        // ================= contact force at the end-of-step velocity, drive torque =================
        {
            f3 wn = vw + aw * dt, ln = v0 + vl + al * dt;
            const float wgt = 1.0f / (float)parts;
            fsum.x += wgt * (-Sbt * ln.x + (-Stz * wn.y + Sty * wn.z));
            fsum.y += wgt * (-Sbt * ln.y + (Stz * wn.x - Stx * wn.z));
            fsum.z += wgt * (F0z - Sbn * ln.z + (-Sny * wn.x + Snx * wn.y));
        }
        // ================= integrate =================
        if (joint) {
            drive = mtv(R, tau0 - wdot * dd_pd);
            jw = jw + mtv(R, wdot) * dt;
            float n2 = dot3(jw, jw);
            if (n2 > P.max_w * P.max_w) jw = jw * (P.max_w * rsqrtf(n2));
            jq = qnormalize(qmul(jq, exp_quat(jw * dt)));
        }
        {   // root, computed redundantly by every lane from lane 0's acceleration
            f3 a0w = shfl3(aw, 0), a0l = shfl3(al, 0);
            f3 wn = w0 + a0w * dt, vO = a0l * dt;                    // vO: in-frame velocity gained by the pelvis point
            float n2 = dot3(wn, wn);
            if (n2 > P.max_w * P.max_w) wn = wn * (P.max_w * rsqrtf(n2));    // maxAngularVelocity (humanoid.py:685-688)
            q0 = qnormalize(qmul(exp_quat(wn * dt), q0));
            f3 vn = v0 + vO;
            p0 = p0 + vn * dt;
            v0 = vn + cross3(wn, vO * dt);                           // re-reference to the moved origin (second order)
            w0 = wn;
        }
        if (++part == parts) { part = 0; ++sub; }
    }

    // ================= refresh: forward kinematics -> rigid-body state, DOF state =================
    {
        f4 qw = q0; f3 x = p0, wv = w0, lv = v0;
#pragma unroll 1
        for (int L = 1; L <= max_level; ++L) {
            f4 qp = shfl4(qw, par); f3 xp = shfl3(x, par); f3 wp = shfl3(wv, par); f3 lp = shfl3(lv, par);
            if (level == L) {
                f3 t = qrot(qp, offset);
                x = xp + t;
                qw = qmul(qp, jq);
                lv = lp + cross3(wp, t);
                wv = wp + qrot(qw, jw);
            }
        }
        if (body) {
            float* o = s_rb[warp] + b * 13;
            o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = qw.x; o[4] = qw.y; o[5] = qw.z; o[6] = qw.w;
            o[7] = lv.x; o[8] = lv.y; o[9] = lv.z; o[10] = wv.x; o[11] = wv.y; o[12] = wv.z;
        }
        __syncwarp();
        float4* g = reinterpret_cast<float4*>(P.rb + (size_t)env * EML_NB * 13);
        const float4* s4 = reinterpret_cast<const float4*>(s_rb[warp]);
        for (int i = lane; i < 78; i += 32) g[i] = s4[i];
        if (lane == 0) {
            float* r = P.root + (size_t)env * 13;
            r[0] = p0.x; r[1] = p0.y; r[2] = p0.z; r[3] = q0.x; r[4] = q0.y; r[5] = q0.z; r[6] = q0.w;
            r[7] = v0.x; r[8] = v0.y; r[9] = v0.z; r[10] = w0.x; r[11] = w0.y; r[12] = w0.z;
        }
        if (body) {
            float* c = P.contact + ((size_t)env * EML_NB + b) * 3;
            float inv = n_sub > 0 ? 1.0f / (float)n_sub : 0.f;      // mean force over the sub-steps
            c[0] = fsum.x * inv; c[1] = fsum.y * inv; c[2] = fsum.z * inv;
        }
        if (joint) {
            const int d = 3 * (b - 1);
            f3 e = log_quat(jq);
            float2* ds = reinterpret_cast<float2*>(P.dof + ((size_t)env * EML_ND + d) * 2);
            ds[0] = make_float2(e.x, jw.x); ds[1] = make_float2(e.y, jw.y); ds[2] = make_float2(e.z, jw.z);
            *reinterpret_cast<float4*>(P.jq + ((size_t)env * EML_NJ + (b - 1)) * 4) = make_float4(jq.x, jq.y, jq.z, jq.w);
            float* df = P.dof_force + (size_t)env * EML_ND + d;
            df[0] = drive.x; df[1] = drive.y; df[2] = drive.z;
        }
    }
}

static EmlModelDev* g_model_dev = nullptr;
const EmlModelDev* eml_model_dev() { return g_model_dev; }

cudaError_t eml_upload_model(const EmlModelDev* m) {
    cudaError_t e;
    if (!g_model_dev && (e = cudaMalloc(&g_model_dev, sizeof(EmlModelDev))) != cudaSuccess) return e;
    return cudaMemcpy(g_model_dev, m, sizeof(EmlModelDev), cudaMemcpyHostToDevice);
}

void eml_fill_phys_params(emloco_sim* s, PhysParams& P);
static void fill_params(emloco_sim* s, PhysParams& P) { eml_fill_phys_params(s, P); }
void eml_fill_phys_params(emloco_sim* s, PhysParams& P) {
    P.model = g_model_dev; P.env_model = s->env_model;
    P.actions = nullptr; P.pd_target = s->pd_target; P.actions_copy = nullptr;
    P.root = s->root_state; P.dof = s->dof_state; P.jq = s->joint_quat; P.rb = s->rb_state;
    P.contact = s->contact; P.dof_force = s->dof_force;
    P.height = s->height; P.hf_rows = s->hf_rows; P.hf_cols = s->hf_cols; P.hf_max = s->hf_max;
    P.env_ids = nullptr; P.reset_mask = nullptr; P.init_root = nullptr; P.init_dof = nullptr; P.N = s->N; P.n_sub = 0; P.dt = s->cfg.sim_dt / (float)s->cfg.substeps;
    P.gz = s->cfg.gravity_z; P.kn = s->cfg.contact_stiffness; P.cn = s->cfg.contact_damping;
    P.ct = s->cfg.friction_damping; P.mu = s->cfg.friction_mu; P.max_w = s->cfg.max_ang_vel;
    P.max_effort = s->cfg.max_effort; P.max_turn = s->cfg.max_turn;
    P.fk_only = 0;
}

cudaError_t eml_launch_physics_soa(emloco_sim* s, const float* d_actions, int n_substeps, cudaStream_t st);

cudaError_t eml_launch_physics(emloco_sim* s, const float* d_actions, int n_substeps, int, cudaStream_t st) {
    if (s->physics_impl == 0) return eml_launch_physics_soa(s, d_actions, n_substeps, st);   // lane-per-env kernel (physics_soa.cu)
    PhysParams P; fill_params(s, P);
    P.actions = d_actions; P.actions_copy = d_actions ? s->actions : nullptr; P.n_sub = n_substeps;
    int blocks = (s->N + PH_WARPS - 1) / PH_WARPS;
    if (P.env_model) physics_kernel<true><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    else physics_kernel<false><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t eml_launch_fk(emloco_sim* s, const int32_t* d_env_ids, int n, cudaStream_t st) {
    PhysParams P; fill_params(s, P);
    P.fk_only = 1; P.env_ids = d_env_ids; P.N = d_env_ids ? n : s->N;
    if (P.N <= 0) return cudaSuccess;
    int blocks = (P.N + PH_WARPS - 1) / PH_WARPS;
    if (P.env_model) physics_kernel<true><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    else physics_kernel<false><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
}

// Device-side reset of the envs whose reset flag is set (env_reset(done_indices) of play_steps,
// amp_continuous_value.py:45 -> humanoid.py:455-481 with a fixed synthetic initial state): no host round trip.
cudaError_t eml_reset_done(emloco_sim* s, const float* d_init_root, const float* d_init_dof, cudaStream_t st) {
    PhysParams P; fill_params(s, P);
    P.fk_only = 1; P.reset_mask = s->reset; P.init_root = d_init_root; P.init_dof = d_init_dof;
    int blocks = (P.N + PH_WARPS - 1) / PH_WARPS;
    if (P.env_model) physics_kernel<true><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    else physics_kernel<false><<<blocks, PH_WARPS * 32, 0, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = eml_launch_post_reset(s, s->traj_on, st);
    if (e != cudaSuccess || s->traj_on != 1) return e;   // 2 = deferred: the caller runs the stage (emloco_traj_reset(sim, NULL))
    return eml_traj_reset(s, s->traj, 1, st);       // _reset_task runs after the observations (humanoid_amp_task.py:54-57)
}
