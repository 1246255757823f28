/* TEST INFRASTRUCTURE ONLY - fp64 CPU restatement of the articulated-body sub-step that the CUDA
 * kernel emloco_b200/csrc/physics.cu implements.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this.
 *
 * PARITY UNPINNED: the reference's physics is NVIDIA Isaac Gym 1.0.preview4 / PhysX 5 (TGS), whose
 * binaries are git-ignored and absent from the reference tree (isaacgym/python/isaacgym/_bindings
 * missing; call sites pacer/pacer/env/tasks/base_task.py:792-797, humanoid.py:137-155,1202).  The
 * reference holds no test, golden vector or fixture for the physics step.  What is restated here is
 * the *contract* the reference code relies on, from its own files:
 *   - model: pacer/pacer/data/assets/mjcf/smpl_humanoid.xml (24 bodies, 23 x 3 hinges, densities)
 *   - drive: DOF_MODE_POS, torque = kp*(target-q) - kd*qd  (isaacgym/docs programming/physics.rst;
 *            humanoid.py:905-910), armature 0.02, 3-hinge joints treated as spherical joints whose
 *            position is an exponential map (humanoid.py:1359-1360, utils/motion_lib_smpl.py:611-614)
 *   - stepping: dt 1/60, 2 substeps per simulate, 2 simulate per env step (pacer.yaml:42,94)
 *   - gravity -9.81 z, ground friction 1.0, restitution 0 (base_task.py:225-231, pacer.yaml:71-73)
 *   - tensors: root [13], rigid body [24,13], dof (pos,vel), net contact force [24,3], dof force
 *            (isaacgym/docs programming/tensors.rst)
 * and the algorithm is the builder's own (DESIGN.md "Physics"): Featherstone ABA in world-aligned
 * coordinates about the pelvis, implicit PD (stiffness and damping folded into the joint-space
 * inertia), ground contact as an implicit spring-damper folded into the contacting body's spatial
 * inertia, regularised Coulomb friction.  This file uses dense 6x6 spatial algebra on purpose: it is
 * structurally independent of the CUDA kernel's block-wise formulation.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC physics_oracle.c -o _build/libphysics_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define NB 24
#define ND 69

typedef struct {
    int    parent[NB];
    double offset[NB][3];
    double mass[NB];
    double com[NB][3];
    double inertia[NB][6];     /* xx xy xz yy yz zz */
    double kp[NB], kd[NB], arm[NB];   /* per joint (index = body) */
    int    geom_type[NB];
    double geom_a[NB][3], geom_b[NB][3], geom_r[NB];
} OModel;

typedef struct {
    double dt;                 /* sub-step */
    double gravity_z;
    double kn, cn, ct, mu;
    double max_ang_vel;
    double max_turn;           /* adaptive refinement threshold: largest body rotation per integration part [rad]; <= 0: off */
    double max_effort;         /* drive torque limit per DOF (MJCF motor gear 500 -> Isaac Gym DOF `effort`); <= 0: unlimited */
    int    hf_rows, hf_cols;
} OCfg;

/* ---------- small linear algebra ---------- */
static void cross(const double* a, const double* b, double* o) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static void quat_to_mat(const double* q, double R[3][3]) { /* xyzw */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0][0] = 1 - 2 * (y * y + z * z); R[0][1] = 2 * (x * y - z * w); R[0][2] = 2 * (x * z + y * w);
    R[1][0] = 2 * (x * y + z * w); R[1][1] = 1 - 2 * (x * x + z * z); R[1][2] = 2 * (y * z - x * w);
    R[2][0] = 2 * (x * z - y * w); R[2][1] = 2 * (y * z + x * w); R[2][2] = 1 - 2 * (x * x + y * y);
}
static void qmul(const double* a, const double* b, double* o) {
    double r[4];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    r[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    memcpy(o, r, sizeof r);
}
static void qnormalize(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
}
static void exp_quat(const double* v, double* q) { /* rotation vector -> quaternion */
    double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double s = a > 1e-8 ? sin(0.5 * a) / a : 0.5 - a * a / 48.0;
    q[0] = v[0] * s; q[1] = v[1] * s; q[2] = v[2] * s; q[3] = cos(0.5 * a);
}
static void log_quat(const double* qin, double* v) { /* quaternion -> rotation vector in (-pi, pi] */
    double q[4] = {qin[0], qin[1], qin[2], qin[3]};
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    double s = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    double k = s > 1e-8 ? 2.0 * atan2(s, q[3]) / s : 2.0;
    v[0] = q[0] * k; v[1] = q[1] * k; v[2] = q[2] * k;
}
static void mat3_vec(double R[3][3], const double* v, double* o) {
    double r[3];
    for (int i = 0; i < 3; ++i) r[i] = R[i][0] * v[0] + R[i][1] * v[1] + R[i][2] * v[2];
    memcpy(o, r, sizeof r);
}
static void mat3T_vec(double R[3][3], const double* v, double* o) {
    double r[3];
    for (int i = 0; i < 3; ++i) r[i] = R[0][i] * v[0] + R[1][i] * v[1] + R[2][i] * v[2];
    memcpy(o, r, sizeof r);
}
static void skew(const double* r, double X[3][3]) {
    X[0][0] = 0; X[0][1] = -r[2]; X[0][2] = r[1];
    X[1][0] = r[2]; X[1][1] = 0; X[1][2] = -r[0];
    X[2][0] = -r[1]; X[2][1] = r[0]; X[2][2] = 0;
}
/* spatial motion cross v x m and force cross v x* f; vectors are (angular, linear) */
static void crm(const double* v, const double* m, double* o) {
    double a[3], b[3], c[3];
    cross(v, m, a); cross(v, m + 3, b); cross(v + 3, m, c);
    o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
    o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
static void crf(const double* v, const double* f, double* o) {
    double a[3], b[3], c[3];
    cross(v, f, a); cross(v + 3, f + 3, b); cross(v, f + 3, c);
    o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
    o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
/* solve A x = b, A n x n SPD-ish, Gaussian elimination with partial pivoting (n <= 6) */
static void solve(int n, double A[6][6], double* b) {
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (p != c) { for (int k = 0; k < n; ++k) { double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; } double t = b[c]; b[c] = b[p]; b[p] = t; }
        for (int r = c + 1; r < n; ++r) {
            double f = A[r][c] / A[c][c];
            for (int k = c; k < n; ++k) A[r][k] -= f * A[c][k];
            b[r] -= f * b[c];
        }
    }
    for (int c = n - 1; c >= 0; --c) {
        for (int k = c + 1; k < n; ++k) b[c] -= A[c][k] * b[k];
        b[c] /= A[c][c];
    }
}
static void inv3(double D[3][3], double Di[3][3]) {
    double A[6][6]; double e[3];
    for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = D[i][j];
        e[0] = e[1] = e[2] = 0; e[c] = 1;
        solve(3, A, e);
        for (int i = 0; i < 3; ++i) Di[i][c] = e[i];
    }
}

static double ground_height(const OCfg* c, const int16_t* hf, double x, double y) {
    if (!hf) return 0.0;
    long px = (long)(x / 0.1), py = (long)(y / 0.1);
    if (px < 0) px = 0; if (px > c->hf_rows - 1) px = c->hf_rows - 1;
    if (py < 0) py = 0; if (py > c->hf_cols - 1) py = c->hf_cols - 1;
    return hf[px * c->hf_cols + py] * 0.005;
}

/* candidate contact points of body i in its own frame; returns count */
static int contact_points(const OModel* m, int i, double pts[8][3], double* drop) {
    *drop = 0.0;
    if (m->geom_type[i] == 0) { memcpy(pts[0], m->geom_a[i], 24); *drop = m->geom_r[i]; return 1; }
    if (m->geom_type[i] == 1) { memcpy(pts[0], m->geom_a[i], 24); memcpy(pts[1], m->geom_b[i], 24); *drop = m->geom_r[i]; return 2; }
    for (int k = 0; k < 8; ++k)
        for (int a = 0; a < 3; ++a) pts[k][a] = m->geom_a[i][a] + ((k >> a) & 1 ? 1.0 : -1.0) * m->geom_b[i][a];
    return 8;
}

/* One articulated sub-step for one env.
 * root[13] = pos3 quat4 lin3 ang3 (world; lin = velocity of the pelvis origin)
 * jq[23][4] joint rotations parent->child; jw[69] joint rates in the child frame; target[69] exp-map targets
 * outputs (may be NULL): contact[24][3] += force of this sub-step, dof_force[69] = drive torque (child frame) */
static void substep(const OModel* m, const OCfg* c, double* root, double* jq, double* jw, const double* target,
                    const int16_t* hf, double* contact, double* dof_force) {
    const double dt = c->dt;
    double R[NB][3][3], x[NB][3], v[NB][6], cj[NB][6], ww[NB][3];
    double IA[NB][6][6], pA[NB][6], U[NB][6][3], Dinv[NB][3][3], u[NB][3], tau0[NB][3], sat[NB];
    double fc0[NB][8][3], Bc[NB][8][3], rc[NB][8][3]; int nact[NB];
    const double* O = root;   /* reference point: pelvis position, fixed during the sub-step */

    /* ---- pass 1: kinematics, root -> leaves ---- */
    quat_to_mat(root + 3, R[0]);
    x[0][0] = x[0][1] = x[0][2] = 0;
    /* Dynamics are evaluated in the inertial frame that translates with the pelvis' velocity V at the start of the sub-step
     * (Galilean invariance): spatial velocities stay O(relative motion), so the velocity-product terms never contain the
     * large, mutually cancelling  w x V  pieces.  V re-enters only where absolute velocities matter (contacts). */
    const double V[3] = {root[7], root[8], root[9]};
    for (int k = 0; k < 3; ++k) { v[0][k] = root[10 + k]; v[0][3 + k] = 0.0; }
    memset(cj[0], 0, sizeof cj[0]);
    for (int i = 1; i < NB; ++i) {
        int p = m->parent[i];
        double t[3], Rj[3][3];
        mat3_vec(R[p], m->offset[i], t);
        for (int k = 0; k < 3; ++k) x[i][k] = x[p][k] + t[k];
        quat_to_mat(jq + 4 * (i - 1), Rj);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
            R[i][a][b] = R[p][a][0] * Rj[0][b] + R[p][a][1] * Rj[1][b] + R[p][a][2] * Rj[2][b];
        mat3_vec(R[i], jw + 3 * (i - 1), ww[i]);          /* joint rate in world axes */
        double vJ[6];
        vJ[0] = ww[i][0]; vJ[1] = ww[i][1]; vJ[2] = ww[i][2];
        cross(x[i], ww[i], vJ + 3);                        /* velocity at O of a rotation about the anchor */
        for (int k = 0; k < 6; ++k) v[i][k] = v[p][k] + vJ[k];
        crm(v[p], vJ, cj[i]);
    }
    /* ---- rigid-body inertias about O in world axes, bias forces, contacts, drive ---- */
    for (int i = 0; i < NB; ++i) {
        double cw[3], Ib[3][3], Iw[3][3], T[3][3], C[3][3];
        mat3_vec(R[i], m->com[i], cw);
        for (int k = 0; k < 3; ++k) cw[k] += x[i][k];
        const double* I6 = m->inertia[i];
        Ib[0][0] = I6[0]; Ib[0][1] = Ib[1][0] = I6[1]; Ib[0][2] = Ib[2][0] = I6[2];
        Ib[1][1] = I6[3]; Ib[1][2] = Ib[2][1] = I6[4]; Ib[2][2] = I6[5];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
            T[a][b] = R[i][a][0] * Ib[0][b] + R[i][a][1] * Ib[1][b] + R[i][a][2] * Ib[2][b];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
            Iw[a][b] = T[a][0] * R[i][b][0] + T[a][1] * R[i][b][1] + T[a][2] * R[i][b][2];
        skew(cw, C);
        double ms = m->mass[i];
        memset(IA[i], 0, sizeof IA[i]);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
            double cc = 0; for (int k = 0; k < 3; ++k) cc += C[a][k] * C[b][k];      /* C C^T */
            IA[i][a][b] = Iw[a][b] + ms * cc;
            IA[i][a][3 + b] = ms * C[a][b];
            IA[i][3 + a][b] = ms * C[b][a];
        }
        for (int a = 0; a < 3; ++a) IA[i][3 + a][3 + a] = ms;
        double Iv[6], gf[6];
        for (int a = 0; a < 6; ++a) { Iv[a] = 0; for (int b = 0; b < 6; ++b) Iv[a] += IA[i][a][b] * v[i][b]; }
        crf(v[i], Iv, pA[i]);
        double g[3] = {0, 0, ms * c->gravity_z};
        cross(cw, g, gf); gf[3] = g[0]; gf[4] = g[1]; gf[5] = g[2];
        for (int a = 0; a < 6; ++a) pA[i][a] -= gf[a];

        /* ground contact: implicit spring-damper folded into IA / pA */
        double pts[8][3], drop; int np = contact_points(m, i, pts, &drop);
        nact[i] = 0;
        for (int k = 0; k < np; ++k) {
            double r[3];
            mat3_vec(R[i], pts[k], r);
            for (int a = 0; a < 3; ++a) r[a] += x[i][a];
            r[2] -= drop;
            double h = ground_height(c, hf, O[0] + r[0], O[1] + r[1]);
            double gap = O[2] + r[2] - h;
            if (gap >= 0) continue;
            double vp[3], wr[3];
            cross(v[i], r, wr);
            for (int a = 0; a < 3; ++a) vp[a] = V[a] + v[i][3 + a] + wr[a];
            double bn = c->kn * dt + c->cn;
            double fn_est = -c->kn * gap - bn * vp[2];
            if (fn_est <= 0) continue;                      /* separating: no adhesion */
            double vt = sqrt(vp[0] * vp[0] + vp[1] * vp[1]);
            double bt = c->ct;
            if (bt * vt > c->mu * fn_est) bt = c->mu * fn_est / vt;   /* regularised Coulomb cone */
            int q = nact[i]++;
            rc[i][q][0] = r[0]; rc[i][q][1] = r[1]; rc[i][q][2] = r[2];
            Bc[i][q][0] = bt; Bc[i][q][1] = bt; Bc[i][q][2] = bn;
            fc0[i][q][0] = 0; fc0[i][q][1] = 0; fc0[i][q][2] = -c->kn * gap;
            double J[3][6], X[3][3];
            skew(r, X);
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { J[a][b] = -X[a][b]; J[a][3 + b] = (a == b); }
            double w[3];   /* B J v - f0 */
            for (int a = 0; a < 3; ++a) w[a] = Bc[i][q][a] * vp[a] - fc0[i][q][a];
            for (int a = 0; a < 6; ++a) {
                for (int b = 0; b < 6; ++b) {
                    double sacc = 0; for (int k2 = 0; k2 < 3; ++k2) sacc += J[k2][a] * Bc[i][q][k2] * J[k2][b];
                    IA[i][a][b] += dt * sacc;
                }
                double sacc = 0; for (int k2 = 0; k2 < 3; ++k2) sacc += J[k2][a] * w[k2];
                pA[i][a] += sacc;
            }
        }
        if (i > 0) {   /* implicit PD: tau = kp (tgt - q - dt w') - kd w',  w' = w + dt wdot */
            /* position error on SO(3): the rotation that takes the joint from q to exp(target), as a rotation vector in
             * the child frame, e = log(q^-1 * exp(target)).  Equal to (target - log q) to first order, but defined for
             * every target (component-wise targets of norm > pi are legal actions) and zero exactly at the target. */
            double qt[4], qi[4], qd[4], e[3], t0[3];
            exp_quat(target + 3 * (i - 1), qt);
            qi[0] = -jq[4 * (i - 1)]; qi[1] = -jq[4 * (i - 1) + 1]; qi[2] = -jq[4 * (i - 1) + 2]; qi[3] = jq[4 * (i - 1) + 3];
            qmul(qi, qt, qd);
            log_quat(qd, e);
            /* effort limit: when the PD torque at the current state exceeds max_effort on any axis, the whole drive
             * (spring, damper and their implicit part) is scaled back so that the largest component equals the limit */
            double tmax = 0;
            for (int k = 0; k < 3; ++k) {
                double te = m->kp[i] * e[k] - m->kd[i] * jw[3 * (i - 1) + k];
                if (fabs(te) > tmax) tmax = fabs(te);
            }
            sat[i] = (c->max_effort > 0 && tmax > c->max_effort) ? c->max_effort / tmax : 1.0;
            for (int k = 0; k < 3; ++k)
                t0[k] = sat[i] * (m->kp[i] * e[k] - (m->kd[i] + m->kp[i] * dt) * jw[3 * (i - 1) + k]);
            mat3_vec(R[i], t0, tau0[i]);
        }
    }
    /* ---- pass 2: articulated inertias, leaves -> root (bodies are in DFS order: child index > parent) ---- */
    for (int i = NB - 1; i >= 1; --i) {
        int p = m->parent[i];
        double S[6][3], X[3][3];
        skew(x[i], X);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { S[a][b] = (a == b); S[3 + a][b] = X[a][b]; }
        for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) { U[i][a][b] = 0; for (int k = 0; k < 6; ++k) U[i][a][b] += IA[i][a][k] * S[k][b]; }
        double D[3][3];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { D[a][b] = 0; for (int k = 0; k < 6; ++k) D[a][b] += S[k][a] * U[i][k][b]; }
        double dd = m->arm[i] + sat[i] * dt * (m->kd[i] + m->kp[i] * dt);
        for (int a = 0; a < 3; ++a) D[a][a] += dd;
        inv3(D, Dinv[i]);
        for (int a = 0; a < 3; ++a) { u[i][a] = tau0[i][a]; for (int k = 0; k < 6; ++k) u[i][a] -= S[k][a] * pA[i][k]; }
        double W[6][3];
        for (int a = 0; a < 6; ++a) for (int b = 0; b < 3; ++b) { W[a][b] = 0; for (int k = 0; k < 3; ++k) W[a][b] += U[i][a][k] * Dinv[i][k][b]; }
        double Ia[6][6], pa[6];
        for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) { Ia[a][b] = IA[i][a][b]; for (int k = 0; k < 3; ++k) Ia[a][b] -= W[a][k] * U[i][b][k]; }
        for (int a = 0; a < 6; ++a) {
            pa[a] = pA[i][a];
            for (int b = 0; b < 6; ++b) pa[a] += Ia[a][b] * cj[i][b];
            for (int k = 0; k < 3; ++k) pa[a] += W[a][k] * u[i][k];
        }
        for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) IA[p][a][b] += Ia[a][b]; pA[p][a] += pa[a]; }
    }
    /* ---- pass 3: accelerations, root -> leaves ---- */
    double acc[NB][6], wdot[NB][3];
    {
        double A[6][6], b[6];
        memcpy(A, IA[0], sizeof A);
        for (int a = 0; a < 6; ++a) b[a] = -pA[0][a];
        solve(6, A, b);
        memcpy(acc[0], b, sizeof b);
    }
    for (int i = 1; i < NB; ++i) {
        int p = m->parent[i];
        double ap[6], rhs[3];
        for (int a = 0; a < 6; ++a) ap[a] = acc[p][a] + cj[i][a];
        for (int a = 0; a < 3; ++a) { rhs[a] = u[i][a]; for (int k = 0; k < 6; ++k) rhs[a] -= U[i][k][a] * ap[k]; }
        for (int a = 0; a < 3; ++a) wdot[i][a] = Dinv[i][a][0] * rhs[0] + Dinv[i][a][1] * rhs[1] + Dinv[i][a][2] * rhs[2];
        double sx[3];
        cross(x[i], wdot[i], sx);
        for (int a = 0; a < 3; ++a) { acc[i][a] = ap[a] + wdot[i][a]; acc[i][3 + a] = ap[3 + a] + sx[a]; }
    }
    /* ---- outputs that use the end-of-step velocities ---- */
    for (int i = 0; i < NB; ++i) {
        double vn[6];
        for (int a = 0; a < 6; ++a) vn[a] = v[i][a] + dt * acc[i][a];
        if (contact) for (int q = 0; q < nact[i]; ++q) {
            double wr[3]; cross(vn, rc[i][q], wr);
            for (int a = 0; a < 3; ++a) contact[3 * i + a] += fc0[i][q][a] - Bc[i][q][a] * (V[a] + vn[3 + a] + wr[a]);
        }
        if (i > 0) {
            double tw[3], tb[3];
            double dd = sat[i] * dt * (m->kd[i] + m->kp[i] * dt);
            for (int a = 0; a < 3; ++a) tw[a] = tau0[i][a] - dd * wdot[i][a];
            mat3T_vec(R[i], tw, tb);
            if (dof_force) for (int a = 0; a < 3; ++a) dof_force[3 * (i - 1) + a] = tb[a];
        }
    }
    /* ---- integrate (semi-implicit Euler) ---- */
    for (int i = 1; i < NB; ++i) {
        double wb[3], dq[4], h[3];
        mat3T_vec(R[i], wdot[i], wb);
        double* w = jw + 3 * (i - 1);
        for (int a = 0; a < 3; ++a) w[a] += dt * wb[a];
        double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (n > c->max_ang_vel) for (int a = 0; a < 3; ++a) w[a] *= c->max_ang_vel / n;
        for (int a = 0; a < 3; ++a) h[a] = dt * w[a];
        exp_quat(h, dq);
        qmul(jq + 4 * (i - 1), dq, jq + 4 * (i - 1));
        qnormalize(jq + 4 * (i - 1));
    }
    {
        double wn[3], vO[3], h[3], dq[4], t[3];
        for (int a = 0; a < 3; ++a) { wn[a] = v[0][a] + dt * acc[0][a]; vO[a] = v[0][3 + a] + dt * acc[0][3 + a]; }
        double wlen = sqrt(wn[0] * wn[0] + wn[1] * wn[1] + wn[2] * wn[2]);
        if (wlen > c->max_ang_vel) for (int a = 0; a < 3; ++a) wn[a] *= c->max_ang_vel / wlen;   /* maxAngularVelocity, humanoid.py:685-688 */
        for (int a = 0; a < 3; ++a) h[a] = dt * wn[a];
        exp_quat(h, dq);
        qmul(dq, root + 3, root + 3);
        qnormalize(root + 3);
        double dp[3] = {dt * vO[0], dt * vO[1], dt * vO[2]};   /* in-frame displacement of the pelvis (second order) */
        cross(wn, dp, t);                                   /* re-reference the spatial velocity to the new origin */
        for (int a = 0; a < 3; ++a) {
            root[7 + a] = V[a] + vO[a] + t[a];
            root[a] += dt * (V[a] + vO[a]);
            root[10 + a] = wn[a];
        }
    }
}

/* forward kinematics -> rigid-body state [24][13] and dof positions (exp map) */
static void refresh(const OModel* m, const double* root, const double* jq, const double* jw, double* rb, double* dof_pos) {
    double R[NB][3][3], x[NB][3], w[NB][3], vl[NB][3], q[NB][4];
    quat_to_mat(root + 3, R[0]);
    for (int k = 0; k < 3; ++k) { x[0][k] = root[k]; vl[0][k] = root[7 + k]; w[0][k] = root[10 + k]; }
    memcpy(q[0], root + 3, 32);
    for (int i = 1; i < NB; ++i) {
        int p = m->parent[i];
        double t[3], wt[3], ww[3];
        mat3_vec(R[p], m->offset[i], t);
        for (int k = 0; k < 3; ++k) x[i][k] = x[p][k] + t[k];
        qmul(q[p], jq + 4 * (i - 1), q[i]);
        quat_to_mat(q[i], R[i]);
        cross(w[p], t, wt);
        mat3_vec(R[i], jw + 3 * (i - 1), ww);
        for (int k = 0; k < 3; ++k) { vl[i][k] = vl[p][k] + wt[k]; w[i][k] = w[p][k] + ww[k]; }
        if (dof_pos) log_quat(jq + 4 * (i - 1), dof_pos + 3 * (i - 1));
    }
    for (int i = 0; i < NB; ++i) {
        double* o = rb + 13 * i;
        for (int k = 0; k < 3; ++k) { o[k] = x[i][k]; o[7 + k] = vl[i][k]; o[10 + k] = w[i][k]; }
        for (int k = 0; k < 4; ++k) o[3 + k] = q[i][k];
    }
}

/* largest rigid-body angular speed of the current state */
static double max_body_ang_vel(const OModel* m, const double* root, const double* jq, const double* jw) {
    double q[NB][4], w[NB][3], best = 0;
    memcpy(q[0], root + 3, 32);
    for (int k = 0; k < 3; ++k) w[0][k] = root[10 + k];
    for (int i = 0; i < NB; ++i) {
        if (i > 0) {
            int p = m->parent[i];
            double R[3][3], ww[3];
            qmul(q[p], jq + 4 * (i - 1), q[i]);
            quat_to_mat(q[i], R);
            mat3_vec(R, jw + 3 * (i - 1), ww);
            for (int k = 0; k < 3; ++k) w[i][k] = w[p][k] + ww[k];
        }
        double n = sqrt(w[i][0] * w[i][0] + w[i][1] * w[i][1] + w[i][2] * w[i][2]);
        if (n > best) best = n;
    }
    return best;
}

/* Public: n_sub sub-steps for N envs (OpenMP over envs).  contact = mean force over the sub-steps. */
void emloco_oracle_step(const OModel* m, const OCfg* c, int N, int n_sub, double* root, double* jq, double* jw,
                        const double* target, const int16_t* hf, double* rb, double* dof_pos, double* contact,
                        double* dof_force) {
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; ++e) {
        double* ct = contact + (size_t)e * NB * 3;
        memset(ct, 0, sizeof(double) * NB * 3);
        for (int s = 0; s < n_sub; ++s) {
            /* adaptive refinement: the explicit velocity-product terms gain energy like (|w| dt)^2, so a sub-step is
             * split into k equal parts until no body turns more than max_turn radians in one part (k <= 8) */
            double* rt = root + (size_t)e * 13; double* q = jq + (size_t)e * 92; double* w = jw + (size_t)e * ND;
            int k = 1;
            if (c->max_turn > 0) {
                double wmax = max_body_ang_vel(m, rt, q, w);
                k = (int)ceil(wmax * c->dt / c->max_turn);
                k = k < 1 ? 1 : (k > 8 ? 8 : k);
            }
            OCfg cc = *c; cc.dt = c->dt / k;
            double part[NB * 3];
            memset(part, 0, sizeof part);
            for (int j = 0; j < k; ++j) substep(m, &cc, rt, q, w, target + (size_t)e * ND, hf, part, dof_force + (size_t)e * ND);
            for (int j = 0; j < NB * 3; ++j) ct[j] += part[j] / k;
        }
        for (int k = 0; k < NB * 3; ++k) ct[k] /= (double)n_sub;
        refresh(m, root + (size_t)e * 13, jq + (size_t)e * 92, jw + (size_t)e * ND, rb + (size_t)e * NB * 13,
                dof_pos + (size_t)e * ND);
    }
}

void emloco_oracle_refresh(const OModel* m, int N, const double* root, const double* jq, const double* jw, double* rb,
                           double* dof_pos) {
    for (int e = 0; e < N; ++e)
        refresh(m, root + (size_t)e * 13, jq + (size_t)e * 92, jw + (size_t)e * ND, rb + (size_t)e * NB * 13,
                dof_pos + (size_t)e * ND);
}

/* exp-map -> joint quaternion (used when the caller writes dof positions, set_dof_state_tensor_indexed) */
void emloco_oracle_expmap_to_quat(int n, const double* e, double* q) {
    for (int i = 0; i < n; ++i) exp_quat(e + 3 * i, q + 4 * i);
}

int emloco_oracle_sizeof_model(void) { return (int)sizeof(OModel); }
int emloco_oracle_sizeof_cfg(void) { return (int)sizeof(OCfg); }
