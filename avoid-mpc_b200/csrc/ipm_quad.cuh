// Batched interior-point solve of the quadrotor collision-avoidance NLP (sm_100a, FP64):
// FOUR LANES PER MPC INSTANCE, up to eight instances per warp, persistent warps that refill
// from a queue.  Same algorithm as the warp-per-instance kernel (ipm_solve.cuh); DESIGN.md section 2.
//
// Replaces ObstacleAvoidanceMPC::Solve -> casadi::nlpsol("ipopt") and the CasADi-generated
// nlp_f / nlp_grad_f / nlp_hess_l of tools/mpc_obstacle_casadi.py
// (src/HighLvlMpc.cpp:93-137; tools/mpc_obstacle_casadi.py:51-242,338-357).
//
// NLP (tools/mpc_obstacle_casadi.py:156-220): w = [X_0,U_0,...,U_{N-1},X_N],
//   min  sum_k (U_k-u_ref)'Q_u(U_k-u_ref) + l_k(X_{k+1})
//   s.t. X_0 = x0,  X_{k+1} = Phi X_k + Gam U_k + gam  (RK4x4 of an affine ODE),
//        lb <= U_k <= ub                                 (src/HighLvlMpc.cpp:70-92)
//   l_k = path quadratic in the yaw-rotated error + sum_j lambda*softplus(-32(r-R))*|v.n|
//   (k < N-1),  terminal quadratic (k = N-1).
//
// Method (DESIGN.md "Solver"): primal-dual log barrier on the control box (IPOPT's monotone mu
// rule, fraction-to-the-boundary rule, inertia-correction ladder, multiplier safeguard);
// iterates stay on the dynamics manifold; the Newton system of the multiple-shooting NLP is
// solved exactly by a stage-wise Riccati sweep; projected Armijo line search on the barrier
// objective; |v.n| smoothed inside the solver as sqrt(s^2+eps^2)-eps, eps = max(eps_min, mu).
//
// Mapping.  The work of one instance is ~850 FMA per Riccati stage and 16 collision terms per
// cost stage: far too little for 32 lanes (the warp-per-instance kernel of round 1 spent 37 k
// warp-instructions per iteration, 15x the arithmetic).  Here
//   * a QUAD (4 adjacent lanes) owns one instance: lane a < 3 = axis chain a (p_a, v_a, acc_a;
//     control a), lane 3 = the yaw chain (control 3).  Every vector of the iterate is walked by
//     "its" lane, the Riccati recursion keeps block row a of P (three 3x3 blocks) in the
//     registers of lane a and exchanges only S (3x3), b and Y = S^-1 Bm' per stage;
//   * the cost / gradient / Hessian EVALUATION is pooled over the warp: every (instance, stage)
//     pair that some quad asked for in this pass is an item, items are dealt to all 32 lanes
//     stage-major, so the cost of a pass follows the number of requests, not the slowest quad,
//     and a lone straggler has its 20 stages evaluated by 20 lanes at once;
//   * the warp runs PASSES: [quad phase: Armijo test of the trial evaluated last pass; if
//     accepted, multiplier update + convergence test + barrier update in ONE loop over the
//     stages (adjoint recursion), then ONE Riccati sweep, then forward sweep + multiplier step
//     length + next trial point in ONE loop] -> [pooled evaluation of the requested points].
//     A failed inertia test or a rejected trial costs that quad one more pass while the others
//     carry on; a quad whose instance has finished pulls the next one from an atomic counter;
//   * every evaluation is a full one (value, gradient, Hessian) AT THE TRIAL POINT: an accepted
//     trial (the common case) has the next iteration's derivatives ready, so an iteration is
//     one evaluation + three stage loops.
// Control flow of the quad phase is warp-uniform (a block runs if any quad needs it, effects are
// predicated), so that every shuffle uses the full mask.
// Memory.  Per instance, in SHARED memory [element][quad]: controls, multipliers, their slack
// reciprocals and the step (3.8 KB at N = 20: six warps = 48 instances per SM); in a GLOBAL
// workspace [element][quad] per resident warp, sized to stay in L2: the two state trajectories
// (iterate / trial), trial controls, cost gradients, reduced gradient, stage Hessians, feedback
// gains -- everything a stage loop reads exactly once, fetched one stage ahead.  Results do not
// depend on which instances share a warp.
#pragma once
#include "common.cuh"

namespace ampc {

#ifndef AMPC_GZ
#define AMPC_GZ 9.81 // tools/mpc_obstacle_casadi.py:39
#endif

struct SolveOut {
    double cost, kkt_dual, kkt_compl, mu;
    int32_t iters, status, n_reg, n_backtrack;
};

#define AMPC_QUADS 8 // quads per warp (at most)

// ---- shared-memory state of one quad, in rows (one row = one double per quad of the warp).
// Per stage k a block of 32 rows: eight per-control fields x 4 controls, so that one pointer
// bumped per stage reaches all of them with compile-time offsets; then the cost gradients q.
enum { QF_U = 0, QF_ZL = 4, QF_ZU = 8, QF_DU = 12, QF_ISL = 16, QF_ISU = 20, QF_STAGE = 24 };
//   u, zl, zu          controls and bound multipliers
//   du                 feed-forward kf during the backward sweep, then the Newton step
//   isl, isu           1 / (u - lb), 1 / (ub - u)
__host__ __device__ inline int quad_smem_rows(int N) { return QF_STAGE * N; }
// per-stage block of the global workspace read once per loop, one stage ahead:
//   r   grad_u of the objective          gr  r_k + Gam' lam_{k+1}: reduced gradient
enum { GF_R = 0, GF_GR = 4, GF_STAGE = 8 };

// rows of the global workspace
struct QuadGmem {
    int x0, x1, ut, q, rg, Hc, Kg, cs, fp, fd, total;
    __host__ __device__ explicit QuadGmem(int N) {
        int o = 0;
        x0 = o, o += 10 * (N + 1);
        x1 = o, o += 10 * (N + 1);
        ut = o, o += 4 * N;
        q = o, o += 10 * (N + 1); // grad_x of the objective, q_k[i] at 10 k + i
        rg = o, o += GF_STAGE * N;
        Hc = o, o += 27 * N; // cost stage kc at 27 kc: pp[a][b] at 3a+b, pv[a][b] at 9+3a+b, vv[a][b] at 18+3a+b
        Kg = o, o += 28 * N; // stage k: axis a at 9a (control l, component c at 3l+c), yaw at 27
        cs = o, o += 2 * N;
        fp = o, o += N; // objective of cost stage kc (smoothed)
        fd = o, o += N; // un-smoothed minus smoothed collision cost of stage kc
        total = o;
    }
    __host__ __device__ int x(int b) const { return b ? x1 : x0; }
};

// per-quad mailbox the pooled evaluation reads
struct QuadBox {
    double eps;     // smoothing of the requested evaluation
    int32_t inst;   // instance id
    int16_t buf;    // which x buffer to read
    int16_t utrial; // controls: 1 = trial controls (global ut), 0 = iterate (shared u)
};

struct QuadTables {
    Chain ch[4];
    double lb[4], ub[4], qu[4];
    double qg[4][3], qp[4][3], gam[4][3]; // goal / path weights and gam by chain component
    QuadBox box[AMPC_QUADS];
    double Y[AMPC_QUADS][3][9]; // per-stage exchange of Y = S^-1 Bm' between the axis lanes
    // scheduling hints of the pooled evaluation (never results): bit kc of heavy[q] = cost stage kc
    // of quad q had collision terms within range at its last evaluation; order[] = this pass's
    // items, the heavy ones first, so that a round of 32 lanes is all-heavy or all-light
    unsigned long long heavy[AMPC_QUADS];
    unsigned short order[AMPC_QUADS * 64];
};
__host__ __device__ inline size_t quad_smem_bytes(int N, int Q) {
    return ((sizeof(QuadTables) + 15) & ~(size_t)15) + (size_t)quad_smem_rows(N) * Q * 8;
}
__host__ __device__ inline size_t quad_ws_bytes_per_warp(int N, int Q) { return (size_t)QuadGmem(N).total * Q * 8; }

// state index of component c of chain i
__device__ __forceinline__ int chain_state(int i, int c) { return i < 3 ? (c == 0 ? i : (c == 1 ? 4 + i : 7 + i)) : 3; }
__device__ __forceinline__ int sym3(int a, int b) { // a <= b in 0..2 -> 0..5
    return a * 3 - (a * (a - 1)) / 2 + (b - a);
}
// FP64 min / max as one compare + select (fmin / fmax are emulated with NaN handling, ~10
// instructions each; here a NaN operand simply loses the comparison, NaNs are tracked apart)
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }
// Quad collectives: ALWAYS executed by the whole warp with the full mask (see header).
__device__ __forceinline__ double qshfl(double v, int src) { return __shfl_sync(AMPC_FULL_MASK, v, src, 4); }
__device__ __forceinline__ double qsum(double v) {
    v += __shfl_xor_sync(AMPC_FULL_MASK, v, 1);
    v += __shfl_xor_sync(AMPC_FULL_MASK, v, 2);
    return v;
}
__device__ __forceinline__ double qmax(double v) {
    v = dmax(v, __shfl_xor_sync(AMPC_FULL_MASK, v, 1));
    return dmax(v, __shfl_xor_sync(AMPC_FULL_MASK, v, 2));
}
__device__ __forceinline__ double qmin(double v) {
    v = dmin(v, __shfl_xor_sync(AMPC_FULL_MASK, v, 1));
    return dmin(v, __shfl_xor_sync(AMPC_FULL_MASK, v, 2));
}
#define QANY(p) __any_sync(AMPC_FULL_MASK, (p))

// element `row` relative to a pointer into a quad's column: rows are (1 << QS) doubles apart
#define RW(ptr, row) (ptr)[(row) << QS]

// accumulators of one cost stage
struct StageAcc {
    double acc, accd;   // objective; un-smoothed minus smoothed collision cost
    double g[7];        // gradient wrt p (0..2) and v (4..6)
    double hpp[6], hpv[9], hvv[6];
};

// one collision term lambda * softplus(-32 (r - R)) * psi(v.n)  (mpc_obstacle_casadi.py:186-204;
// closed forms: SURVEY.md 8a, generalised to the smoothed |s|: psi = sqrt(s^2+eps^2) - eps).
// d = o - p, n = d/r, s = v.n, w = (v - s n)/r, Pi = I - nn'.  With e = exp(-32 (r - R)),
// sp = log(1+e), sig = e/(1+e):
//   grad_p = a_n n - a_w w,  grad_v = a_w n                 a_n = 32 lam sig psi, a_w = lam sp psi'
//   H_pp = (c_nn + c_d) nn' - c_1 (nw' + wn') + c_ww ww' - c_d I
//   H_pv = (c_nw + c_sp) nn' - c_ww wn' - c_sp I,   H_vv = c_ww nn'
// (c_1 = c_nw + c_sp, c_d = c_pi + c_sp s / r): every block is a sum of two outer products and a
// multiple of I, accumulated as such.  Norms come from rsqrt (1 ulp) instead of sqrt + divide;
// for e < 2^-7 (clearance > 0.15 m: almost every term) log(1+e) and 1/(1+e) are their series to
// degree 8 (relative error < 2e-18; the reference's un-stabilised log(1+exp(x)) is itself only
// good to 1.1e-16 / e there), beyond that the library log and a division.
__device__ __forceinline__ void collision_term(StageAcc &A, const double *x, double d0, double d1, double d2, double r2,
                                               double radius, double lam, double eps, double eps2) {
    const double ir = rsqrt(r2);
    const double rr = r2 * ir;
    const double n0 = d0 * ir, n1 = d1 * ir, n2 = d2 * ir;
    const double sv = x[4] * n0 + x[5] * n1 + x[6] * n2;
    const double e = exp((rr - radius) * -32.0);
    double sp, sig;
    if (e < 0.0078125) {
        const double l = 1.0 + e * (-1.0 / 2 + e * (1.0 / 3 + e * (-1.0 / 4 + e * (1.0 / 5 + e * (-1.0 / 6 + e * (1.0 / 7 + e * (-1.0 / 8)))))));
        const double v = 1.0 + e * (-1.0 + e * (1.0 + e * (-1.0 + e * (1.0 + e * (-1.0 + e * (1.0 + e * (-1.0 + e)))))));
        sp = e * l;
        sig = e * v;
    } else {
        sp = log(1.0 + e);
        sig = e / (1.0 + e);
    }
    const double h2 = sv * sv + eps2;
    const double ih = rsqrt(h2);
    const double hyp = h2 * ih;
    const double psi = hyp - eps;
    const double lsp = lam * sp;
    A.acc += lsp * psi;
    A.accd += lsp * (fabs(sv) - psi);
    const double dpsi = sv * ih;
    const double ddpsi = eps2 * ih * ih * ih;
    const double w0 = (x[4] - sv * n0) * ir, w1 = (x[5] - sv * n1) * ir, w2 = (x[6] - sv * n2) * ir;
    const double nn[3] = {n0, n1, n2}, ww[3] = {w0, w1, w2};
    const double ls32 = lam * 32.0 * sig;
    const double a_n = ls32 * psi; // grad p along n
    const double a_w = lsp * dpsi; // grad p along -w, grad v along n
    const double c_nn = 32.0 * ls32 * (1.0 - sig) * psi;
    const double c_nw = ls32 * dpsi;
    const double c_ww = lsp * ddpsi;
    const double c_sp = a_w * ir;
    const double c_1 = c_nw + c_sp;
    const double c_d = a_n * ir + c_sp * sv * ir;
    const double c_a = c_nn + c_d;
    double pa[3], qa[3], ta[3], ua[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        A.g[a] += a_n * nn[a] - a_w * ww[a];
        A.g[4 + a] += a_w * nn[a];
        pa[a] = c_a * nn[a] - c_1 * ww[a];  // H_pp = p n' + q w' - c_d I
        qa[a] = c_ww * ww[a] - c_1 * nn[a];
        ta[a] = c_1 * nn[a] - c_ww * ww[a]; // H_pv = t n' - c_sp I
        ua[a] = c_ww * nn[a];               // H_vv = u n'
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (b >= a) {
                A.hpp[sym3(a, b)] += pa[a] * nn[b] + qa[a] * ww[b];
                A.hvv[sym3(a, b)] += ua[a] * nn[b];
            }
            A.hpv[a * 3 + b] += ta[a] * nn[b];
        }
    A.hpp[0] -= c_d, A.hpp[3] -= c_d, A.hpp[5] -= c_d;
    A.hpv[0] -= c_sp, A.hpv[4] -= c_sp, A.hpv[8] -= c_sp;
}

// ---- pooled evaluation: objective, gradient and Hessian of cost stage kc of one instance
// (mpc_obstacle_casadi.py:162-214).  One lane per item.  Reads the state from x buffer `buf`, the
// controls from the trial controls or the iterate; writes q, r (shared) and the stage Hessian,
// row by row for the three axis lanes of the sweep (global).
template <int QS>
__device__ __forceinline__ void eval_item(const SolveConsts &c, const QuadGmem &LG, double *__restrict__ S,
                                          double *__restrict__ G, int kc, const double *__restrict__ prefix, int buf,
                                          int utrial, double eps, unsigned long long *heavy) {
    const int N = c.N, K = c.K, k = kc + 1;
    const double *qg = c.wgt, *qp = c.wgt + 10, *qu = c.wgt + 20;
    const double lam = c.wgt[24];
    StageAcc A;
    A.acc = 0.0, A.accd = 0.0;
    double *pc = S + ((QF_STAGE * kc) << QS);
    // control term (mpc_obstacle_casadi.py:209-210)
    {
        const double *up = utrial ? G + ((LG.ut + 4 * kc) << QS) : pc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double d = RW(up, i) - (i == 2 ? AMPC_GZ : 0.0);
            A.acc += qu[i] * d * d;
            RW(G, LG.rg + GF_STAGE * kc + GF_R + i) = 2.0 * qu[i] * d;
        }
    }
    double x[10];
    {
        const double *xp = G + ((LG.x(buf) + 10 * k) << QS);
#pragma unroll
        for (int i = 0; i < 10; ++i)
            x[i] = RW(xp, i);
    }
    double *qo = G + ((LG.q + 10 * k) << QS);
    if (kc == N - 1) { // terminal (mpc_obstacle_casadi.py:168-170)
        const double *tg = prefix + 10 + 10 * N + 3 * K * N;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const double d = x[i] - tg[i];
            A.acc += qg[i] * d * d;
            RW(qo, i) = 2.0 * qg[i] * d;
        }
        RW(G, LG.fp + kc) = A.acc;
        RW(G, LG.fd + kc) = 0.0;
        return;
    }
    bool any_near = false;
    // path term in the yaw-rotated frame (mpc_obstacle_casadi.py:172-185,206-208)
    const double *ref = prefix + 10 + 10 * kc;
    const double cy = RW(G, LG.cs + 2 * kc), sy = RW(G, LG.cs + 2 * kc + 1); // cos(yaw), sin(-yaw)
    double dl[10];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        dl[i] = x[i] - ref[i];
    const double rd0 = cy * dl[0] - sy * dl[1], rd1 = sy * dl[0] + cy * dl[1];
    const double rd4 = cy * dl[4] - sy * dl[5], rd5 = sy * dl[4] + cy * dl[5];
    A.acc += qp[0] * rd0 * rd0 + qp[1] * rd1 * rd1 + qp[4] * rd4 * rd4 + qp[5] * rd5 * rd5;
    A.acc += qp[2] * dl[2] * dl[2] + qp[3] * dl[3] * dl[3] + qp[6] * dl[6] * dl[6] + qp[7] * dl[7] * dl[7] +
             qp[8] * dl[8] * dl[8] + qp[9] * dl[9] * dl[9];
    A.g[0] = 2.0 * (cy * qp[0] * rd0 + sy * qp[1] * rd1);
    A.g[1] = 2.0 * (-sy * qp[0] * rd0 + cy * qp[1] * rd1);
    A.g[4] = 2.0 * (cy * qp[4] * rd4 + sy * qp[5] * rd5);
    A.g[5] = 2.0 * (-sy * qp[4] * rd4 + cy * qp[5] * rd5);
    A.g[2] = 2.0 * qp[2] * dl[2];
    A.g[6] = 2.0 * qp[6] * dl[6];
    A.g[3] = 0.0;
    RW(qo, 3) = 2.0 * qp[3] * dl[3];
    RW(qo, 7) = 2.0 * qp[7] * dl[7];
    RW(qo, 8) = 2.0 * qp[8] * dl[8];
    RW(qo, 9) = 2.0 * qp[9] * dl[9];
    A.hpp[0] = 2.0 * (cy * cy * qp[0] + sy * sy * qp[1]);
    A.hpp[1] = 2.0 * (cy * sy * (qp[1] - qp[0]));
    A.hpp[2] = 0.0;
    A.hpp[3] = 2.0 * (sy * sy * qp[0] + cy * cy * qp[1]);
    A.hpp[4] = 0.0;
    A.hpp[5] = 2.0 * qp[2];
    A.hvv[0] = 2.0 * (cy * cy * qp[4] + sy * sy * qp[5]);
    A.hvv[1] = 2.0 * (cy * sy * (qp[5] - qp[4]));
    A.hvv[2] = 0.0;
    A.hvv[3] = 2.0 * (sy * sy * qp[4] + cy * cy * qp[5]);
    A.hvv[4] = 0.0;
    A.hvv[5] = 2.0 * qp[6];
#pragma unroll
    for (int i = 0; i < 9; ++i)
        A.hpv[i] = 0.0;
    // collision terms, two per trip so that their dependent chains (sqrt, 1/r, exp, log, ...)
    // interleave.  A term whose softplus argument is below -40 (clearance > 1.25 m, which also
    // covers the (1e4,1e4,1e4) padding points) is < 1e-16*|v.n| in value and < 3e-15 in any
    // derivative -- below the rounding of the sums it is added to: a pair is skipped when both
    // of its terms are that far (stages are almost always all-near or all-far: 71 % / 23 % of
    // the benchmark's stages).
    const double *ob = prefix + 10 + 10 * N + 3 * K * kc;
    const double far2 = (c.radius + 1.25) * (c.radius + 1.25);
    const double eps2 = eps * eps;
    const double radius = c.radius;
    // the points of the next pair are fetched while this pair computes; an odd K reads one point
    // past its stage (the next stage's first neighbour, or the target for the last one: valid
    // memory) and replaces it by the padding point
    double o[6];
#pragma unroll
    for (int i = 0; i < 6; ++i)
        o[i] = ob[i];
    for (int j = 0; j < K; j += 2) {
        const bool two = j + 1 < K;
        const double a0 = o[0] - x[0], a1 = o[1] - x[1], a2 = o[2] - x[2];
        const double b0 = (two ? o[3] : 1e4) - x[0], b1 = (two ? o[4] : 1e4) - x[1], b2 = (two ? o[5] : 1e4) - x[2];
        if (j + 2 < K) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
                o[i] = ob[3 * (j + 2) + i];
        }
        const double ra = a0 * a0 + a1 * a1 + a2 * a2, rb = b0 * b0 + b1 * b1 + b2 * b2;
        if (ra > far2 && rb > far2)
            continue;
        any_near = true;
        // both terms of the pair are evaluated (one straight-line block): the far one of a mixed
        // pair adds its true, negligible value instead of exactly nothing
        collision_term(A, x, a0, a1, a2, ra, radius, lam, eps, eps2);
        collision_term(A, x, b0, b1, b2, rb, radius, lam, eps, eps2);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        RW(qo, i) = A.g[i];
        RW(qo, 4 + i) = A.g[4 + i];
    }
    double *ho = G + ((LG.Hc + 27 * kc) << QS);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const int lo = a < b ? a : b, hi = a < b ? b : a;
            RW(ho, 3 * a + b) = A.hpp[sym3(lo, hi)];
            RW(ho, 9 + 3 * a + b) = A.hpv[3 * a + b]; // d2 / dp_a dv_b
            RW(ho, 18 + 3 * a + b) = A.hvv[sym3(lo, hi)];
        }
    RW(G, LG.fp + kc) = A.acc;
    RW(G, LG.fd + kc) = A.accd;
    if (kc < 64) {
        if (any_near)
            atomicOr(heavy, 1ull << kc);
        else
            atomicAnd(heavy, ~(1ull << kc));
    }
}

// y = F' x
__device__ __forceinline__ void chain_FT(const Chain &f, const double x[3], double y[3]) {
    y[0] = f.d1 * x[0];
    y[1] = f.c1 * x[0] + f.d2 * x[1];
    y[2] = f.c2 * x[0] + f.c3 * x[1] + f.c4 * x[2];
}

// ---- loop 1 of an iteration, stages N-1 .. 0, lane a = control a / chain a.  For quads with
// `acc`: the multiplier step of the accepted trial (IPOPT eq. (16) safeguard), u := u_t, new
// slack reciprocals.  For all: complementarity measures, and the adjoint recursion
//   lam_k = q_k + Phi'lam_{k+1},  gr_k = r_k + Gam'lam_{k+1}  (reduced gradient, stored),
//   |grad_{U_k} L|_inf = |gr_k - zl_k + zu_k|_inf.
// Effects are stored for quads with `on`.  Outputs are per lane; the caller reduces.
template <int QS>
__device__ __forceinline__ void quad_adjoint(const SolveConsts &c, const QuadGmem &LG, double *__restrict__ S,
                                             const double *__restrict__ G, const QuadTables *tb, int a, bool on,
                                             bool acc, double mu, double a_du, double *ed_out, double *ec_out,
                                             double *cm_out, bool *nan_out) {
    const int N = c.N;
    const Chain fa = tb->ch[a];
    const double lo = tb->lb[a], hi = tb->ub[a];
    const double kappa_sigma = 1e10, inv_kappa = 1e-10;
    const bool axis = a < 3;
    double *pc = S + ((QF_STAGE * (N - 1) + a) << QS);               // control block of stage k
    const double *pq = G + ((LG.q + 10 * N + (axis ? a : 3)) << QS); // q_{k+1}, this chain
    const double *pu = G + ((LG.ut + 4 * (N - 1) + a) << QS);
    double *pg = const_cast<double *>(G) + ((LG.rg + GF_STAGE * (N - 1) + a) << QS);
    double lv[3];
    lv[0] = RW(pq, 0);
    lv[1] = axis ? RW(pq, 4) : 0.0;
    lv[2] = axis ? RW(pq, 7) : 0.0;
    double ed = 0.0, ec = 0.0, cm = 0.0;
    bool nan_seen = false;
    // operands from the global workspace are fetched one stage ahead
    double un_next = *pu, r_next = RW(pg, GF_R), q_next[3];
    pq -= 10 << QS;
    q_next[0] = RW(pq, 0), q_next[1] = axis ? RW(pq, 4) : 0.0, q_next[2] = axis ? RW(pq, 7) : 0.0;
#pragma unroll 2
    for (int k = N - 1; k >= 0; --k) {
        const double un = un_next, rk = r_next;
        const double qk0 = q_next[0], qk1 = q_next[1], qk2 = q_next[2];
        if (k > 0) {
            pu -= 4 << QS;
            un_next = *pu;
            r_next = RW(pg, GF_R - GF_STAGE);
            pq -= 10 << QS;
            q_next[0] = RW(pq, 0);
            if (axis)
                q_next[1] = RW(pq, 4), q_next[2] = RW(pq, 7);
        }
        double uu = RW(pc, QF_U), zl = RW(pc, QF_ZL), zu = RW(pc, QF_ZU);
        if (acc) {
            const double isl0 = RW(pc, QF_ISL), isu0 = RW(pc, QF_ISU), du = RW(pc, QF_DU);
            const double dzl = mu * isl0 - zl - zl * isl0 * du;
            const double dzu = mu * isu0 - zu + zu * isu0 * du;
            uu = un;
            const double isl = 1.0 / (uu - lo), isu = 1.0 / (hi - uu);
            const double mil = mu * isl, miu = mu * isu;
            zl += a_du * dzl;
            zu += a_du * dzu;
            zl = dmax(dmin(zl, kappa_sigma * mil), mil * inv_kappa);
            zu = dmax(dmin(zu, kappa_sigma * miu), miu * inv_kappa);
            RW(pc, QF_U) = uu;
            RW(pc, QF_ZL) = zl;
            RW(pc, QF_ZU) = zu;
            RW(pc, QF_ISL) = isl;
            RW(pc, QF_ISU) = isu;
        }
        const double sl = uu - lo, su = hi - uu;
        const double cl = sl * zl, cu = su * zu;
        ec = dmax(ec, dmax(cl, cu));
        cm = dmax(cm, dmax(fabs(cl - mu), fabs(cu - mu)));
        const double gr = rk + (fa.g1 * lv[0] + fa.g2 * lv[1] + fa.g3 * lv[2]);
        const double gu = gr - zl + zu;
        if (on)
            RW(pg, GF_GR) = gr;
        ed = dmax(ed, fabs(gu));
        nan_seen |= !(gu == gu);
        pc -= QF_STAGE << QS;
        pg -= GF_STAGE << QS;
        if (k > 0) {
            double fl[3];
            chain_FT(fa, lv, fl);
            lv[0] = qk0 + fl[0];
            lv[1] = axis ? qk1 + fl[1] : 0.0;
            lv[2] = axis ? qk2 + fl[2] : 0.0;
        }
    }
    *ed_out = ed, *ec_out = ec, *cm_out = cm, *nan_out = nan_seen;
}

// ---- loop 2: Riccati backward sweep, all quads of the warp in step.  Lane a < 3 keeps block
// row a of P (blocks P^(ab), b = 0..2, 3x3 each) and p^(a) in registers; per stage
//   t = P^(ab) G_b,  S_ab = G_a't (+R_a),  Bm^(a)_b = F_a't,  A^(ab) = F_a'P^(ab)F_b,
//   S = L D L' (3x3, gathered by shuffles, factored redundantly),  Y^(a) = S^-1 Bm^(a)' (local),
//   P^(ab) <- Q^(ab) + A^(ab) - Bm^(a) Y^(b)   (Y^(b) from lane b through shared memory).
// Lane 3 runs the same instructions on zeros (its shuffles keep the warp converged; what it
// computes is never gathered); the decoupled scalar recursion of the yaw chain is computed by
// every lane and stored by lane 3.  R_k = 2 Q_u + zl/sl + zu/su and the barrier gradient
// r_k - mu/sl + mu/su come from the stored reciprocals.  Returns false (quad-uniform) if a pivot
// of S is not positive: the reduced Hessian has the wrong inertia and the caller regularises.
// Writes the feedback gains Kg (global) and the feed-forward kf (shared, in the du slots) of
// quads with `on`.
template <int QS>
__device__ __forceinline__ bool quad_riccati_backward(const SolveConsts &c, const QuadGmem &LG, double *__restrict__ S,
                                                      double *__restrict__ G, QuadTables *tb, int sq, int a, int N,
                                                      double delta, double mu, bool on) {
    const bool axis = a < 3;
    const int aa = axis ? a : 0; // lane 3 reads lane 0's slots (values unused)
    const Chain fa = tb->ch[a];
    const double qp2 = 2.0 * tb->qp[a][2];
    const double qua = 2.0 * tb->qu[a], quy = 2.0 * c.wgt[23];
    bool ok = true; // pivots so far positive (identical in the four lanes of a quad)
    double P[3][9], pv[3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int i = 0; i < 9; ++i)
            P[b][i] = 0.0;
#pragma unroll
    for (int b = 0; b < 3; ++b)
        if (b == a) {
            P[b][0] = 2.0 * tb->qg[a][0] + delta;
            P[b][4] = 2.0 * tb->qg[a][1] + delta;
            P[b][8] = 2.0 * tb->qg[a][2] + delta;
        }
    // q, r and the stage Hessian rows come from the global workspace: fetched one stage ahead
    const double *pq = G + ((LG.q + 10 * N + aa) << QS); // q_{k+1}, this chain (p, v, a at +0, +4, +7)
    const double *py = G + ((LG.q + 10 * N + 3) << QS);  // q_{k+1}, yaw
    const double *pr = G + ((LG.rg + GF_STAGE * (N - 1) + GF_R + a) << QS);
    const double *pr3 = G + ((LG.rg + GF_STAGE * (N - 1) + GF_R + 3) << QS);
    pv[0] = axis ? RW(pq, 0) : 0.0;
    pv[1] = axis ? RW(pq, 4) : 0.0;
    pv[2] = axis ? RW(pq, 7) : 0.0;
    // yaw chain: x+ = d1 x + g1 u, stage Hessian 2 Q_pen[3] (k >= 1), terminal 2 Q_goal[3]
    const double yd1 = c.ch[3].d1, yg1 = c.ch[3].g1;
    const double yq = 2.0 * c.wgt[13] + delta;
    double yP = 2.0 * c.wgt[3] + delta, ypv = *py;
    const double *pc = S + ((QF_STAGE * (N - 1) + a) << QS); // this lane's control block
    const double *pc3 = S + ((QF_STAGE * (N - 1) + 3) << QS); // the yaw control block
    // row aa of pp / pv / vv at hrow + {0, 9, 18} + b, column aa of pv at hcol + 3 b
    const double *hrow = G + ((LG.Hc + 27 * (N > 1 ? N - 2 : 0) + 3 * aa) << QS);
    const double *hcol = G + ((LG.Hc + 27 * (N > 1 ? N - 2 : 0) + 9 + aa) << QS);
    double n_H[12];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        n_H[4 * b + 0] = N > 1 ? RW(hrow, b) : 0.0;
        n_H[4 * b + 1] = N > 1 ? RW(hrow, 9 + b) : 0.0;
        n_H[4 * b + 2] = N > 1 ? RW(hcol, 3 * b) : 0.0;
        n_H[4 * b + 3] = N > 1 ? RW(hrow, 18 + b) : 0.0;
    }
    pq -= 10 << QS; // q_{N-1}
    py -= 10 << QS;
    double n_q[4], n_r = *pr, n_r3 = *pr3;
    n_q[0] = RW(pq, 0), n_q[1] = RW(pq, 4), n_q[2] = RW(pq, 7), n_q[3] = *py;
    double *kgp = G + ((LG.Kg + 28 * (N - 1) + (axis ? 9 * a : 27)) << QS);
    double *yx = tb->Y[sq][aa];
#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
        double Hk[12];
#pragma unroll
        for (int i = 0; i < 12; ++i)
            Hk[i] = n_H[i];
        double qk[3];
        qk[0] = n_q[0], qk[1] = n_q[1], qk[2] = n_q[2];
        const double yqk = n_q[3], rk = n_r, rk3 = n_r3;
        if (k > 0) { // prefetch stage k-1
            pr -= GF_STAGE << QS;
            pr3 -= GF_STAGE << QS;
            n_r = *pr, n_r3 = *pr3;
            if (k > 1) {
                hrow -= 27 << QS;
                hcol -= 27 << QS;
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    n_H[4 * b + 0] = RW(hrow, b);
                    n_H[4 * b + 1] = RW(hrow, 9 + b);
                    n_H[4 * b + 2] = RW(hcol, 3 * b);
                    n_H[4 * b + 3] = RW(hrow, 18 + b);
                }
                pq -= 10 << QS;
                py -= 10 << QS;
                n_q[0] = RW(pq, 0), n_q[1] = RW(pq, 4), n_q[2] = RW(pq, 7), n_q[3] = *py;
            }
        }
        const double zl = RW(pc, QF_ZL), zu = RW(pc, QF_ZU), isl = RW(pc, QF_ISL), isu = RW(pc, QF_ISU);
        const double rdk = qua + zl * isl + zu * isu;
        const double rtk = rk - mu * isl + mu * isu;
        const double yzl = RW(pc3, QF_ZL), yzu = RW(pc3, QF_ZU), yisl = RW(pc3, QF_ISL), yisu = RW(pc3, QF_ISU);
        const double yrd = quy + yzl * yisl + yzu * yisu;
        const double yrt = rk3 - mu * yisl + mu * yisu;
        // (1) products with this lane's blocks; P^(ab) is overwritten by A^(ab)
        double Srow[3], bm[3][3]; // bm[b][i] = Bm^(a)_b, component i of chain a
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const Chain &fb = c.ch[b];
            double t[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                t[i] = P[b][3 * i] * fb.g1 + P[b][3 * i + 1] * fb.g2 + P[b][3 * i + 2] * fb.g3;
            Srow[b] = fa.g1 * t[0] + fa.g2 * t[1] + fa.g3 * t[2] + (a == b ? rdk + delta : 0.0);
            chain_FT(fa, t, bm[b]);
            double M[9];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                M[3 * i] = fb.d1 * P[b][3 * i];
                M[3 * i + 1] = fb.c1 * P[b][3 * i] + fb.d2 * P[b][3 * i + 1];
                M[3 * i + 2] = fb.c2 * P[b][3 * i] + fb.c3 * P[b][3 * i + 1] + fb.c4 * P[b][3 * i + 2];
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                P[b][j] = fa.d1 * M[j];
                P[b][3 + j] = fa.c1 * M[j] + fa.d2 * M[3 + j];
                P[b][6 + j] = fa.c2 * M[j] + fa.c3 * M[3 + j] + fa.c4 * M[6 + j];
            }
        }
        const double bi = rtk + fa.g1 * pv[0] + fa.g2 * pv[1] + fa.g3 * pv[2];
        // (2) S = L D L', every lane redundantly from the same gathered entries
        const double S00 = qshfl(Srow[0], 0), S10 = qshfl(Srow[0], 1), S20 = qshfl(Srow[0], 2);
        const double S11 = qshfl(Srow[1], 1), S21 = qshfl(Srow[1], 2), S22 = qshfl(Srow[2], 2);
        const double d0 = S00, i0 = 1.0 / d0;
        const double l10 = S10 * i0, l20 = S20 * i0;
        const double d1 = S11 - l10 * l10 * d0, i1 = 1.0 / d1;
        const double l21 = (S21 - l20 * l10 * d0) * i1;
        const double d2 = S22 - l20 * l20 * d0 - l21 * l21 * d1, i2 = 1.0 / d2;
        // yaw scalar
        const double yt = yP * yg1;
        const double S33 = yg1 * yt + (yrd + delta);
        const double ybm = yd1 * yt;
        const double ybi = yrt + yg1 * ypv;
        const double i3 = 1.0 / S33;
        ok = ok && (d0 > 0.0) && (d1 > 0.0) && (d2 > 0.0) && (S33 > 0.0);
        if (!QANY(on && ok))
            break; // every quad that is factoring has hit a non-positive pivot
        const bool st = on && ok;
        // feed-forward kff = -S^-1 b
        double kff0, kff1, kff2;
        {
            const double b0 = -qshfl(bi, 0), b1 = -qshfl(bi, 1), b2 = -qshfl(bi, 2);
            const double w1 = b1 - l10 * b0, w2 = b2 - l20 * b0 - l21 * w1;
            kff2 = w2 * i2;
            kff1 = w1 * i1 - l21 * kff2;
            kff0 = b0 * i0 - l10 * kff1 - l20 * kff2;
        }
        const double kff3 = -ybi * i3;
        if (st)
            RW(const_cast<double *>(pc), QF_DU) = a == 0 ? kff0 : (a == 1 ? kff1 : (a == 2 ? kff2 : kff3));
        if (k == 0)
            break; // dx_0 = 0: no feedback gain and no P_0 needed
        pc -= QF_STAGE << QS;
        pc3 -= QF_STAGE << QS;
        // (3) Y^(a) = S^-1 Bm^(a)'  (control l x component cc of chain a): local
        double Y[9];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const double r0 = bm[0][cc], r1 = bm[1][cc], r2 = bm[2][cc];
            const double w1 = r1 - l10 * r0, w2 = r2 - l20 * r0 - l21 * w1;
            Y[6 + cc] = w2 * i2;
            Y[3 + cc] = w1 * i1 - l21 * Y[6 + cc];
            Y[cc] = r0 * i0 - l10 * Y[3 + cc] - l20 * Y[6 + cc];
        }
        __syncwarp(); // the previous stage's reads of the exchange slots are complete
        if (axis) {
            // (lanes that shadow a live quad when fewer than 8 quads are in use have on == false:
            // they must not touch the exchange slots they share with it)
            if (on) {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    yx[i] = Y[i];
            }
            if (st) {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    RW(kgp, i) = -Y[i]; // feedback gain K = -Y
            }
        } else if (st) {
            *kgp = -ybm * i3;
        }
        kgp -= 28 << QS;
        __syncwarp();
        // (4) P^(ab) <- Q^(ab) + A^(ab) - Bm^(a) Y^(b);  p^(a) <- q^(a) + F_a'p^(a) + Bm^(a) kff
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double *yb = tb->Y[sq][b];
            double Yb[9];
#pragma unroll
            for (int i = 0; i < 9; ++i)
                Yb[i] = yb[i];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    P[b][3 * i + j] -= bm[0][i] * Yb[j] + bm[1][i] * Yb[3 + j] + bm[2][i] * Yb[6 + j];
            const double dd = a == b ? delta : 0.0;
            P[b][0] += Hk[4 * b + 0] + dd;
            P[b][1] += Hk[4 * b + 1];
            P[b][3] += Hk[4 * b + 2];
            P[b][4] += Hk[4 * b + 3] + dd;
            P[b][8] += a == b ? qp2 + delta : 0.0;
        }
        if (!axis) { // lane 3's blocks stay exactly zero whatever it read
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                P[b][0] = 0.0, P[b][1] = 0.0, P[b][3] = 0.0, P[b][4] = 0.0;
            }
        }
        double fp[3];
        chain_FT(fa, pv, fp);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            pv[i] = axis ? qk[i] + fp[i] + bm[0][i] * kff0 + bm[1][i] * kff1 + bm[2][i] * kff2 : 0.0;
        yP = yd1 * (yd1 * yP) - i3 * ybm * ybm + yq;
        ypv = yqk + yd1 * ypv + kff3 * ybm;
    }
    return ok;
}

// barrier sum over a lane's controls: log of the product of four stages' slacks at a time
// (a slack is in (0, 20], so eight factors cannot overflow or underflow)
__device__ __noinline__ double log_ool(double v) { return log(v); } // rare: keep the code small
struct BarAcc {
    double sum = 0.0, prod = 1.0;
    __device__ __forceinline__ void add(int k, double sl, double su) {
        prod *= sl * su;
        if ((k & 3) == 3) {
            sum += log_ool(prod);
            prod = 1.0;
        }
    }
    __device__ __forceinline__ double finish() { return sum + log_ool(prod); }
};

// ---- loop 3, stages 0 .. N-1.  With `forward`: the forward sweep du_k = K_k dx_k + kff_k,
// dx_{k+1} = Phi dx_k + Gam du_k, the multiplier step length (IPOPT eq. (15)) and the first
// trial point (alpha = 1) in one go; without: a shorter trial along the stored step.
// (x_t = x + roll-out of the step, NOT a fresh roll-out of u_t: measured, the fresh roll-out
// re-rounds the whole trajectory, and that noise (1e-14 in x, 1e-11 in f) exceeds the Armijo
// slack near convergence -- +4 % iterations, +10 % time.)
// Trial point of the projected line search: u_t = u + clip(alpha du) per component
// (fraction-to-the-boundary rule), x_t = x + roll-out of the clipped step (by linearity),
// written to the other x buffer / the trial controls for quads with `on`; Armijo slope
// g'd = sum_k (gr_k - mu/sl + mu/su) d_k  (reduced-gradient form) and the barrier sum there.
template <int QS>
__device__ __forceinline__ void quad_step(const SolveConsts &c, const QuadGmem &LG, double *__restrict__ S,
                                          double *__restrict__ G, const QuadTables *tb, int a, int cur, bool forward,
                                          double alpha, double tau_f, double mu, bool on, double *gdt_out,
                                          double *bar_out, double *adu_out) {
    const int N = c.N;
    const Chain fa = tb->ch[a];
    const double lo = tb->lb[a], hi = tb->ub[a];
    const bool axis = a < 3;
    const int s0 = axis ? a : 3;
    const double *xc = G + ((LG.x(cur) + 10 + s0) << QS);
    double *xt = G + ((LG.x(cur ^ 1) + 10 + s0) << QS);
    const double *kgp = G + ((LG.Kg + 28 + (axis ? 9 * a : 27)) << QS);
    double *utp = G + ((LG.ut + a) << QS);
    const double *pgr = G + ((LG.rg + GF_GR + a) << QS);
    double *pc = S + (a << QS);
    double xa[3] = {0.0, 0.0, 0.0}; // dx of this chain (unclipped Newton step)
    double xd[3] = {0.0, 0.0, 0.0}; // roll-out of the clipped step
    double gdt = 0.0;
    double rn = 1.0, rd = tau_f; // smallest z / -dz seen so far as a fraction rn / rd (<= 1 / tau_f)
    BarAcc bar;
    double nk[9];
#pragma unroll
    for (int i = 0; i < 9; ++i)
        nk[i] = 0.0;
    double nx[3], ngr = *pgr;
    nx[0] = RW(xc, 0);
    nx[1] = axis ? RW(xc, 4) : 0.0;
    nx[2] = axis ? RW(xc, 7) : 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        double kg[9], xk[3];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            kg[i] = nk[i];
        xk[0] = nx[0], xk[1] = nx[1], xk[2] = nx[2];
        const double grk = ngr;
        if (k + 1 < N) { // operands of the next stage from the global workspace
            pgr += GF_STAGE << QS;
            ngr = *pgr;
            xc += 10 << QS;
            nx[0] = RW(xc, 0);
            if (axis)
                nx[1] = RW(xc, 4), nx[2] = RW(xc, 7);
            if (forward) {
                nk[0] = RW(kgp, 0);
                if (axis) {
#pragma unroll
                    for (int i = 1; i < 9; ++i)
                        nk[i] = RW(kgp, i);
                }
                kgp += 28 << QS;
            }
        }
        const double uu = RW(pc, QF_U), isl = RW(pc, QF_ISL), isu = RW(pc, QF_ISU);
        const double sl = uu - lo, su = hi - uu;
        double du = RW(pc, QF_DU); // kff_k with `forward`, else the stored step
        if (forward) {
            // contribution of this chain's dx to the three axis controls (yaw: to its own)
            double p0 = 0.0, p1 = 0.0, p2 = 0.0;
            const double py = kg[0] * xa[0];
            if (axis) {
                p0 = py + kg[1] * xa[1] + kg[2] * xa[2];
                p1 = kg[3] * xa[0] + kg[4] * xa[1] + kg[5] * xa[2];
                p2 = kg[6] * xa[0] + kg[7] * xa[1] + kg[8] * xa[2];
            }
            p0 = qsum(p0);
            p1 = qsum(p1);
            p2 = qsum(p2);
            du += a == 0 ? p0 : (a == 1 ? p1 : (a == 2 ? p2 : py));
            if (on)
                RW(pc, QF_DU) = du;
            const double y0 = fa.d1 * xa[0] + fa.c1 * xa[1] + fa.c2 * xa[2] + fa.g1 * du;
            const double y1 = fa.d2 * xa[1] + fa.c3 * xa[2] + fa.g2 * du;
            const double y2 = fa.c4 * xa[2] + fa.g3 * du;
            xa[0] = y0, xa[1] = y1, xa[2] = y2;
            // fraction to the boundary for the multipliers: alpha_z = min(1, tau min z / -dz)
            // over the components with dz < 0, kept as a fraction (no division per component)
            const double zl = RW(pc, QF_ZL), zu = RW(pc, QF_ZU);
            const double dzl = mu * isl - zl - zl * isl * du;
            const double dzu = mu * isu - zu + zu * isu * du;
            if (dzl < 0.0 && zl * rd < rn * -dzl)
                rn = zl, rd = -dzl;
            if (dzu < 0.0 && zu * rd < rn * -dzu)
                rn = zu, rd = -dzu;
        }
        double d = alpha * du;
        d = dmin(dmax(d, -tau_f * sl), tau_f * su);
        const double ut = uu + d;
        gdt += (grk - mu * isl + mu * isu) * d;
        bar.add(k, ut - lo, hi - ut);
        const double z0 = fa.d1 * xd[0] + fa.c1 * xd[1] + fa.c2 * xd[2] + fa.g1 * d;
        const double z1 = fa.d2 * xd[1] + fa.c3 * xd[2] + fa.g2 * d;
        const double z2 = fa.c4 * xd[2] + fa.g3 * d;
        xd[0] = z0, xd[1] = z1, xd[2] = z2;
        if (on) {
            *utp = ut;
            RW(xt, 0) = xk[0] + xd[0];
            if (axis) {
                RW(xt, 4) = xk[1] + xd[1];
                RW(xt, 7) = xk[2] + xd[2];
            }
        }
        xt += 10 << QS;
        utp += 4 << QS;
        pc += QF_STAGE << QS;
    }
    *gdt_out = qsum(gdt);
    *bar_out = qsum(bar.finish());
    if (forward)
        *adu_out = qmin(dmin(1.0, tau_f * rn / rd));
}

enum { QS_DONE = 0, QS_FETCH, QS_TEST, QS_FACTOR, QS_TRIAL, QS_FINISH };

// One CTA = one warp = up to Q = 1 << QS instances in flight, refilled from *counter.
template <int QS>
__global__ void __launch_bounds__(256, 1)
ipm_quad_kernel(const __grid_constant__ SolveConsts c, int B, const double *__restrict__ prefix,
                double *__restrict__ w_inout, SolveOut *__restrict__ info,
                const int32_t *__restrict__ active /* nullable: instances with 0 are skipped */,
                const int32_t *__restrict__ order /* nullable: the order instances are taken in */,
                double *__restrict__ ws, int32_t *__restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    constexpr int Q = 1 << QS;
    // a CTA is W independent warps (own shared-memory slice, own workspace, own queue slots) that
    // only meet at one barrier per pass, so that they run the same phase -- the same code -- together
    const int wid = threadIdx.x >> 5, gw = blockIdx.x * (blockDim.x >> 5) + wid;
    unsigned char *smem_raw = smem_all + (size_t)wid * quad_smem_bytes(c.N, Q);
    QuadTables *tb = reinterpret_cast<QuadTables *>(smem_raw);
    double *Sw = reinterpret_cast<double *>(smem_raw + ((sizeof(QuadTables) + 15) & ~(size_t)15));
    const int lane = threadIdx.x & 31, a = lane & 3;
    const int sq = (lane >> 2) & (Q - 1); // lanes of quads >= Q shadow a live quad, all effects off
    const bool live = (lane >> 2) < Q;
    const int N = c.N, n_w = 10 + 14 * N;
    const QuadGmem LG(N);
    double *Gw = ws + (size_t)gw * LG.total * Q;
    double *S = Sw + sq, *G = Gw + sq; // this quad's column
    if (lane < 4) {
        tb->ch[lane] = c.ch[lane];
        tb->lb[lane] = c.lb[lane];
        tb->ub[lane] = c.ub[lane];
        tb->qu[lane] = c.wgt[20 + lane];
        for (int cc = 0; cc < 3; ++cc) {
            const int si = chain_state(lane, cc);
            const bool pad = lane == 3 && cc > 0;
            tb->qg[lane][cc] = pad ? 0.0 : c.wgt[si];
            tb->qp[lane][cc] = pad ? 0.0 : c.wgt[10 + si];
            tb->gam[lane][cc] = pad ? 0.0 : c.gam[si];
        }
    }
    __syncwarp();
    const double lo = tb->lb[a], hi = tb->ub[a];
    const int s0 = chain_state(a, 0), s1 = chain_state(a, 1), s2 = chain_state(a, 2);
    const double kappa_mu = 0.2, tau_min = 0.99, eta = 1e-4;
    const double mu_min = c.tol / 10.0;

    // quad-uniform solver state
    int state = live ? QS_FETCH : QS_DONE;
    int inst = -1, cur = 0, iter = 0, n_reg = 0, n_bt = 0, ntry = 0, ls = 0, status = 1;
    double mu = 0.0, delta = 0.0, delta_last = 0.0, alpha = 1.0, f_cur = 0.0, fns_cur = 0.0, eps_at = 0.0;
    double a_du = 1.0, tau_f = 0.0, gdt = 0.0, bar = 0.0, bar_cur = 0.0, e_dual = 0.0, e_compl = 0.0;

    for (;;) {
        bool want = false;    // this quad asks for an evaluation in this pass
        bool retrial = false; // ... of a shorter trial step, to be built first
        bool acc = false;     // the trial evaluated last pass was accepted
        // ================= quad phase (warp-uniform control flow, predicated effects) =========
        { // ---- Armijo test on the barrier objective at the trial point evaluated last pass
            const bool in_trial = state == QS_TRIAL;
            if (QANY(in_trial)) {
                double ft = 0.0, fdt = 0.0;
                for (int k = a; k < N; k += 4) {
                    ft += RW(G, LG.fp + k);
                    fdt += RW(G, LG.fd + k);
                }
                ft = qsum(ft);
                fdt = qsum(fdt);
                const double phi0 = f_cur - mu * bar_cur;
                const double phit = ft - mu * bar;
                acc = in_trial && phit <= phi0 + eta * gdt + 10.0 * 2.220446049250313e-16 * fabs(phi0);
                if (acc) { // the iterate becomes the trial buffer (controls: in the adjoint loop)
                    cur ^= 1;
                    f_cur = ft;
                    fns_cur = ft + fdt;
                    bar_cur = bar;
                    ++iter;
                    state = QS_TEST;
                } else if (in_trial) {
                    alpha *= 0.5;
                    ++n_bt;
                    if (++ls >= 40) {
                        status = 2;
                        state = QS_FINISH;
                    } else {
                        retrial = true;
                    }
                }
            }
        }
        { // ---- multiplier update of an accepted trial, convergence test, barrier update
          // (IPOPT eq. (7)); also recomputes the reduced gradient after a refresh of the smoothing
            const bool in_test = state == QS_TEST;
            if (QANY(in_test)) {
                double ed, ec, cm;
                bool nan_seen;
                quad_adjoint<QS>(c, LG, S, G, tb, a, in_test, acc, mu, a_du, &ed, &ec, &cm, &nan_seen);
                ed = qmax(ed);
                ec = qmax(ec);
                double c_mu = qmax(cm);
                // a NaN in any lane must reach every lane of the quad (the max drops NaNs)
                const unsigned nb = __ballot_sync(AMPC_FULL_MASK, nan_seen);
                const bool bad = ((nb >> (lane & ~3)) & 0xFu) != 0u || !(f_cur == f_cur);
                bool cont = false;
                if (in_test && ntry < 0) {
                    // second visit after the smoothing was refreshed: only the reduced gradient
                    // (just rewritten) was needed
                    cont = true;
                } else if (in_test) {
                    e_dual = ed;
                    e_compl = ec;
                    if (bad) {
                        status = 3;
                        state = QS_FINISH;
                    } else if (dmax(e_dual, e_compl) <= c.tol) {
                        status = 0;
                        state = QS_FINISH;
                    } else if (iter >= c.max_iter) {
                        status = 1;
                        state = QS_FINISH;
                    } else {
                        cont = true;
                    }
                }
                bool lp = cont && ntry >= 0 && mu > mu_min && dmax(e_dual, c_mu) <= c.kappa_eps * mu;
                while (QANY(lp)) {
                    const double mun = dmax(mu_min, dmin(kappa_mu * mu, mu * sqrt(mu))); // theta_mu = 1.5
                    double c2 = 0.0;
                    const double *pc = S + (a << QS);
                    for (int k = 0; k < N; ++k) {
                        const double uu = RW(pc, QF_U);
                        c2 = dmax(c2, dmax(fabs((uu - lo) * RW(pc, QF_ZL) - mun), fabs((hi - uu) * RW(pc, QF_ZU) - mun)));
                        pc += QF_STAGE << QS;
                    }
                    c2 = qmax(c2);
                    if (lp) {
                        mu = mun;
                        c_mu = c2;
                    }
                    lp = lp && mu > mu_min && dmax(e_dual, c_mu) <= c.kappa_eps * mu;
                }
                if (cont) {
                    const double eps = dmax(c.eps_min, c.eps_scale * mu);
                    if (eps != eps_at) { // the smoothing moved with mu: refresh f, q, Hessian first
                        eps_at = eps;
                        want = true;
                        ntry = -1; // come back here for the reduced gradient, then factor
                        if (a == 0)
                            tb->box[sq].buf = (int16_t)cur, tb->box[sq].utrial = 0, tb->box[sq].eps = eps;
                    } else {
                        tau_f = dmax(tau_min, 1.0 - mu);
                        delta = 0.0;
                        ntry = 0;
                        state = QS_FACTOR;
                    }
                }
            }
        }
        __syncwarp(); // lane 3's multipliers and reciprocals are read by the whole quad in the sweep
        if (state == QS_FINISH) {
            // results: w = [X_0,U_0,...,X_N]; objective without smoothing
            double *wo = w_inout + (size_t)inst * n_w;
            for (int e = a; e < 10 * (N + 1); e += 4) {
                const int k = e / 10, i = e - 10 * k;
                wo[14 * k + i] = RW(G, LG.x(cur) + e);
            }
            for (int k = 0; k < N; ++k)
                wo[14 * k + 10 + a] = RW(S, QF_STAGE * k + QF_U + a);
            if (a == 0) {
                SolveOut o;
                o.cost = fns_cur, o.kkt_dual = e_dual, o.kkt_compl = e_compl, o.mu = mu;
                o.iters = iter, o.status = status, o.n_reg = n_reg, o.n_backtrack = n_bt;
                info[inst] = o;
            }
            state = QS_FETCH;
        }
        { // ---- next instance from the queue
            const bool in_fetch = state == QS_FETCH;
            if (QANY(in_fetch)) {
                // the lanes of an idle quad ran this pass's sweeps with their effects off: order their
                // (discarded) reads of the multiplier rows before the new instance's values land there
                __syncwarp();
                int b = B;
                if (in_fetch && a == 0) {
                    for (;;) {
                        const int t = atomicAdd(counter, 1);
                        b = t < B ? (order ? order[t] : t) : B;
                        if (b >= B || !active || active[b] != 0)
                            break;
                    }
                }
                b = __shfl_sync(AMPC_FULL_MASK, b, 0, 4);
                if (in_fetch && b >= B) {
                    state = QS_DONE;
                } else if (in_fetch) {
                    inst = b;
                    const double *pf = prefix + (size_t)b * c.n_prefix;
                    const double *wi = w_inout + (size_t)b * n_w;
                    mu = c.mu_init, delta_last = 0.0, cur = 0, iter = 0, n_reg = 0, n_bt = 0, status = 1, ntry = 0;
                    // controls from the warm start pushed into the box interior
                    const double pl = dmin(c.bound_push * dmax(1.0, fabs(lo)), c.bound_frac * (hi - lo));
                    const double pu = dmin(c.bound_push * dmax(1.0, fabs(hi)), c.bound_frac * (hi - lo));
                    const Chain fa = tb->ch[a];
                    double xa[3];
                    xa[0] = pf[s0], xa[1] = a < 3 ? pf[s1] : 0.0, xa[2] = a < 3 ? pf[s2] : 0.0;
                    RW(G, LG.x0 + s0) = xa[0], RW(G, LG.x1 + s0) = xa[0];
                    if (a < 3) {
                        RW(G, LG.x0 + s1) = xa[1], RW(G, LG.x1 + s1) = xa[1];
                        RW(G, LG.x0 + s2) = xa[2], RW(G, LG.x1 + s2) = xa[2];
                    }
                    const double g0 = tb->gam[a][0], g1 = tb->gam[a][1], g2 = tb->gam[a][2];
                    BarAcc b0;
                    double *pc = S + (a << QS);
                    for (int k = 0; k < N; ++k) {
                        double uu = wi[14 * k + 10 + a];
                        uu = dmin(dmax(uu, lo + pl), hi - pu);
                        const double isl = 1.0 / (uu - lo), isu = 1.0 / (hi - uu);
                        RW(pc, QF_U) = uu;
                        RW(pc, QF_ISL) = isl;
                        RW(pc, QF_ISU) = isu;
                        RW(pc, QF_ZL) = mu * isl;
                        RW(pc, QF_ZU) = mu * isu;
                        pc += QF_STAGE << QS;
                        b0.add(k, uu - lo, hi - uu);
                        // roll-out X_{k+1} = Phi X_k + Gam U_k + gam, chain by chain
                        const double y0 = (g0 + (fa.d1 * xa[0] + fa.c1 * xa[1] + fa.c2 * xa[2])) + fa.g1 * uu;
                        const double y1 = (g1 + (fa.d2 * xa[1] + fa.c3 * xa[2])) + fa.g2 * uu;
                        const double y2 = (g2 + fa.c4 * xa[2]) + fa.g3 * uu;
                        xa[0] = y0, xa[1] = y1, xa[2] = y2;
                        RW(G, LG.x0 + 10 * (k + 1) + s0) = xa[0];
                        if (a < 3) {
                            RW(G, LG.x0 + 10 * (k + 1) + s1) = xa[1];
                            RW(G, LG.x0 + 10 * (k + 1) + s2) = xa[2];
                        }
                    }
                    bar_cur = b0.finish(); // summed over the quad below
                    for (int k = a; k < N; k += 4) { // cos / sin of the reference yaw
                        const double yaw = pf[10 + 10 * k + 3];
                        RW(G, LG.cs + 2 * k) = cos(yaw);
                        RW(G, LG.cs + 2 * k + 1) = sin(-yaw);
                    }
                    eps_at = dmax(c.eps_min, c.eps_scale * mu);
                    want = true;
                    if (a == 0) {
                        tb->box[sq].buf = 0, tb->box[sq].utrial = 0, tb->box[sq].eps = eps_at, tb->box[sq].inst = b;
                        tb->heavy[sq] = ~0ull; // nothing known yet about the new instance's stages
                    }
                    state = QS_TEST;
                }
                const double bsum = qsum(bar_cur);
                if (in_fetch && state == QS_TEST)
                    bar_cur = bsum;
            }
        }
        bool fwd = false;
        { // ---- Newton system, ONE sweep per pass; inertia correction (IPOPT Algorithm IC ladder)
            const bool fac = state == QS_FACTOR && !want;
            if (QANY(fac)) {
                const bool ok = quad_riccati_backward<QS>(c, LG, S, G, tb, sq, a, N, delta, mu, fac);
                fwd = fac && ok;
                if (fac && !ok) {
                    if (delta == 0.0)
                        delta = (delta_last == 0.0) ? 1e-4 : dmax(1e-20, delta_last / 3.0);
                    else
                        delta *= (delta_last == 0.0) ? 100.0 : 8.0;
                    if (++ntry > 60 || delta > 1e40) {
                        status = 3;
                        state = QS_FINISH; // written out in the next pass
                    }
                }
            }
        }
        // ---- forward sweep + first trial point of quads whose sweep succeeded; shorter trial of
        // quads whose last one was rejected (two rounds of the same loop only if both occur)
        for (int round = 0; round < 2; ++round) {
            const bool mine = round == 0 ? fwd : retrial;
            if (!QANY(mine))
                continue;
            double g_t, b_t, ad = 1.0;
            quad_step<QS>(c, LG, S, G, tb, a, cur, round == 0, round == 0 ? 1.0 : alpha, tau_f, mu, mine, &g_t, &b_t, &ad);
            if (mine) {
                gdt = g_t, bar = b_t;
                if (round == 0) {
                    a_du = ad;
                    if (delta > 0.0) {
                        delta_last = delta;
                        ++n_reg;
                    }
                    alpha = 1.0;
                    ls = 0;
                    state = QS_TRIAL;
                }
                want = true;
                if (a == 0)
                    tb->box[sq].buf = (int16_t)(cur ^ 1), tb->box[sq].utrial = 1, tb->box[sq].eps = eps_at;
            }
        }
        // ================= pooled evaluation =================
        __syncwarp();
        const unsigned req = __ballot_sync(AMPC_FULL_MASK, want && a == 0 && live);
        const unsigned alive = __ballot_sync(AMPC_FULL_MASK, state != QS_DONE);
        if (blockDim.x == 32) {
            if (alive == 0u)
                break;
        } else if (__syncthreads_and(alive == 0u)) { // (a finished warp keeps meeting the others: its passes are empty)
            break;
        }
        const int m = __popc(req);
        const int n_items = m * N;
        // item i = (stage i / m, i-th requesting quad); stage-major so that neighbouring lanes read
        // neighbouring columns.  Stages whose collision terms were all out of range last time (23 %
        // of the benchmark's) cost a twentieth of the others: deal the heavy items first.
        const bool sorted = N <= 64 && n_items > 32;
        if (sorted) {
            const unsigned lt = (1u << lane) - 1u;
            int nh = 0;
            for (int base = 0; base < n_items; base += 32) {
                const int i = base + lane, kc = i / m;
                const bool hv = i < n_items && kc < N - 1 && ((tb->heavy[__fns(req, 0, i - kc * m + 1) >> 2] >> kc) & 1ull);
                nh += __popc(__ballot_sync(AMPC_FULL_MASK, hv));
            }
            int ph = 0, pl = nh;
            for (int base = 0; base < n_items; base += 32) {
                const int i = base + lane, kc = i / m;
                const bool valid = i < n_items;
                const bool hv = valid && kc < N - 1 && ((tb->heavy[__fns(req, 0, i - kc * m + 1) >> 2] >> kc) & 1ull);
                const unsigned bh = __ballot_sync(AMPC_FULL_MASK, hv), bl = __ballot_sync(AMPC_FULL_MASK, valid && !hv);
                if (hv)
                    tb->order[ph + __popc(bh & lt)] = (unsigned short)i;
                else if (valid)
                    tb->order[pl + __popc(bl & lt)] = (unsigned short)i;
                ph += __popc(bh), pl += __popc(bl);
            }
            __syncwarp();
        }
        for (int j = lane; j < n_items; j += 32) {
            const int i = sorted ? tb->order[j] : j;
            const int kc = i / m, rnk = i - kc * m;
            const int q = __fns(req, 0, rnk + 1) >> 2;
            const QuadBox bx = tb->box[q];
            eval_item<QS>(c, LG, Sw + q, Gw + q, kc, prefix + (size_t)bx.inst * c.n_prefix, bx.buf, bx.utrial, bx.eps,
                          &tb->heavy[q]);
        }
        __syncwarp();
        // an evaluation of the CURRENT point (first one of an instance, or the refresh after the
        // smoothing moved) updates f; a trial's evaluation is consumed by the Armijo test
        if (QANY(want && state == QS_TEST)) {
            double ft = 0.0, fdt = 0.0;
            for (int k = a; k < N; k += 4) {
                ft += RW(G, LG.fp + k);
                fdt += RW(G, LG.fd + k);
            }
            ft = qsum(ft);
            fdt = qsum(fdt);
            if (want && state == QS_TEST) {
                f_cur = ft;
                fns_cur = ft + fdt;
            }
        }
    }
}

#undef RW
#undef QANY

// ---- longest-first scheduling of the queue.  A call ends when its slowest instance ends, and
// the slowest instances are the ones whose reference path runs closest to (or through) the
// obstacles: instances whose path keeps > 0.5 m of clearance need <= 12 iterations on the
// benchmark scenes, the others up to the cap.  Instances are therefore taken in order of
// increasing clearance (counting sort into 32 classes; the order inside a class is arbitrary).
// Scheduling only: results do not depend on it.
#define AMPC_ORDER_CLASSES 32
__global__ void solve_order_class_kernel(int B, int N, int K, int n_prefix, double radius,
                                         const double *__restrict__ prefix, const int32_t *__restrict__ active,
                                         int32_t *__restrict__ cls, int32_t *__restrict__ hist) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B)
        return;
    const double *p = prefix + (size_t)b * n_prefix;
    double m = 1e300;
    for (int t = lane; t < (N - 1) * K; t += 32) {
        const int k = t / K;
        const double *r = p + 10 + 10 * k, *o = p + 10 + 10 * N + 3 * t;
        const double d0 = o[0] - r[0], d1 = o[1] - r[1], d2 = o[2] - r[2];
        m = fmin(m, d0 * d0 + d1 * d1 + d2 * d2);
    }
    for (int o = 16; o > 0; o >>= 1)
        m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        const double clear = sqrt(m) - radius; // < 0: the path crosses an inflated obstacle
        int c = (int)floor((clear + 0.5) * 16.0);
        c = c < 0 ? 0 : (c > AMPC_ORDER_CLASSES - 1 ? AMPC_ORDER_CLASSES - 1 : c);
        if (active && active[b] == 0)
            c = AMPC_ORDER_CLASSES - 1; // skipped anyway
        cls[b] = c;
        atomicAdd(&hist[c], 1);
    }
}
__global__ void solve_order_scatter_kernel(int B, const int32_t *__restrict__ cls, int32_t *__restrict__ hist,
                                           int32_t *__restrict__ order) {
    // hist[0..31] counts -> exclusive offsets in hist[32..63] (every block recomputes them: 32 adds)
    __shared__ int off[AMPC_ORDER_CLASSES];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < AMPC_ORDER_CLASSES; ++c) {
            off[c] = acc;
            acc += hist[c];
        }
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const int c = cls[b];
        order[off[c] + atomicAdd(&hist[AMPC_ORDER_CLASSES + c], 1)] = b;
    }
}

} // namespace ampc
