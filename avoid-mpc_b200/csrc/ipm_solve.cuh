// Batched interior-point solve of the quadrotor collision-avoidance NLP,
// one warp per MPC instance (sm_100a, FP64).
//
// Replaces ObstacleAvoidanceMPC::Solve -> casadi::nlpsol("ipopt") and the
// CasADi-generated nlp_f / nlp_grad_f / nlp_hess_l of tools/mpc_obstacle_casadi.py
// (src/HighLvlMpc.cpp:93-137; tools/mpc_obstacle_casadi.py:51-242,338-357).
//
// NLP (tools/mpc_obstacle_casadi.py:156-220): w = [X_0,U_0,...,U_{N-1},X_N],
//   min  sum_k (U_k-u_ref)'Q_u(U_k-u_ref) + l_k(X_{k+1})
//   s.t. X_0 = x0,  X_{k+1} = Phi X_k + Gam U_k + gam  (RK4x4 of an affine ODE),
//        lb <= U_k <= ub                                 (src/HighLvlMpc.cpp:70-92)
//   l_k = path quadratic in the yaw-rotated error + sum_j lambda*softplus(-32(r-R))*|v.n|
//   (k < N-1),  terminal quadratic (k = N-1).
//
// Method (see DESIGN.md "Solver"): primal-dual log-barrier on the control box with
// IPOPT's monotone mu rule, fraction-to-the-boundary rule, inertia-correction schedule and
// multiplier safeguard; iterates stay on the dynamics manifold (X = roll-out of U); the
// Newton system of the multiple-shooting NLP is solved exactly by a stage-wise Riccati
// sweep; projected Armijo line search on the barrier objective.  |v.n| is smoothed inside
// the solver as sqrt(s^2+eps^2)-eps with eps = max(eps_min, mu).
//
// Lane mapping.  Cost/gradient/Hessian evaluation: lane = stage (all N stages in parallel,
// the K obstacle terms accumulated in registers).  Value-only evaluation (line-search
// trials): collision terms spread over all 32 lanes.  Riccati sweep: the dynamics split
// into four chains (x, y, z axis: p,v,a; yaw); lane 4i+j owns the 3x3 block P^(ij) coupling
// chain i and chain j IN REGISTERS, S = R + G'PG is gathered by shuffles and factored as
// LDL' redundantly; no matrix lives in shared memory.  Warp-private shared memory holds the
// iterate, the step, gradients, compact (p,v) Hessian blocks and the feedback gains.
#pragma once
#include "common.cuh"
#include "ipm_quad.cuh" // SolveOut

namespace ampc {
namespace v1 { // round-1 kernel, kept for A/B measurements (AMPC_SOLVE_KERNEL=warp)

// per-warp shared-memory layout, in doubles
struct WarpLayout {
    int x, u, zl, zu, dx, du, dxt, dut, q, r, rt, rdiag, Hc, Kg, kf, cs, total;
    __host__ __device__ explicit WarpLayout(int N) {
        int o = 0;
        x = o, o += (N + 1) * 10;
        u = o, o += N * 4;
        zl = o, o += N * 4;
        zu = o, o += N * 4;
        dx = o, o += (N + 1) * 10;
        du = o, o += N * 4;
        dxt = o, o += (N + 1) * 10; // trial displacement of the line search
        dut = o, o += N * 4;
        q = o, o += (N + 1) * 10;
        r = o, o += N * 4;
        rt = o, o += N * 4;
        rdiag = o, o += N * 4;
        Hc = o, o += N * 21; // stage k (1..N-1) at (k-1)*21: pp(6) pv(9) vv(6)
        Kg = o, o += N * 48; // feedback gains: [k][component 0..2][chain-pair lane 0..15]
        kf = o, o += N * 4;
        cs = o, o += N * 2;
        total = (o + 1) & ~1;
    }
};

__host__ __device__ inline size_t solve_smem_bytes(int N, int warps) {
    return sizeof(SolveConsts) + 128 + (size_t)warps * WarpLayout(N).total * sizeof(double);
}

// ---- chain structure.  State order [p(0..2), yaw(3), v(4..6), a(7..9)]; the affine
// dynamics (mpc_obstacle_casadi.py:106-122) decouple into four chains: axis i < 3 with
// components (p_i, v_i, a_i) = state indices (i, 4+i, 7+i) driven by control i, and yaw
// (component 0 = state 3; components 1,2 are padding that stays exactly zero).  Chain i
// advances by the upper-triangular 3x3 matrix F_i and the 3-vector G_i below.
__device__ __forceinline__ Chain load_chain(const double *Phi, const double *Gam, int i) {
    Chain c;
    if (i < 3) {
        c.d1 = Phi[i * 10 + i], c.c1 = Phi[i * 10 + 4 + i], c.c2 = Phi[i * 10 + 7 + i];
        c.d2 = Phi[(4 + i) * 10 + 4 + i], c.c3 = Phi[(4 + i) * 10 + 7 + i];
        c.c4 = Phi[(7 + i) * 10 + 7 + i];
        c.g1 = Gam[i * 4 + i], c.g2 = Gam[(4 + i) * 4 + i], c.g3 = Gam[(7 + i) * 4 + i];
    } else {
        c.d1 = Phi[33], c.c1 = c.c2 = c.d2 = c.c3 = c.c4 = 0.0;
        c.g1 = Gam[15], c.g2 = c.g3 = 0.0;
    }
    return c;
}
// state index of component c of chain i (-1: padding)
__device__ __forceinline__ int chain_state(int i, int c) {
    return i < 3 ? (c == 0 ? i : (c == 1 ? 4 + i : 7 + i)) : (c == 0 ? 3 : -1);
}
// y = F' x   and   y = F x
__device__ __forceinline__ void chain_FT(const Chain &f, const double x[3], double y[3]) {
    y[0] = f.d1 * x[0];
    y[1] = f.c1 * x[0] + f.d2 * x[1];
    y[2] = f.c2 * x[0] + f.c3 * x[1] + f.c4 * x[2];
}
__device__ __forceinline__ void chain_F(const Chain &f, const double x[3], double u, double y[3]) {
    y[0] = f.d1 * x[0] + f.c1 * x[1] + f.c2 * x[2] + f.g1 * u;
    y[1] = f.d2 * x[1] + f.c3 * x[2] + f.g2 * u;
    y[2] = f.c4 * x[2] + f.g3 * u;
}

// sum_l Phi[i][l] * v[l]   (row i of Phi: cols {i} + {i+4, i+7} for p, {i+3} for v)
__device__ __forceinline__ double phi_row_dot(const double *Phi, int i, const double *v) {
    double a = Phi[i * 10 + i] * v[i];
    if (i < 3) {
        a += Phi[i * 10 + i + 4] * v[i + 4];
        a += Phi[i * 10 + i + 7] * v[i + 7];
    } else if (i >= 4 && i <= 6) {
        a += Phi[i * 10 + i + 3] * v[i + 3];
    }
    return a;
}
// sum_l Gam[i][l] * u[l]   (row i of Gam has one non-zero)
__device__ __forceinline__ double gam_row_dot(const double *Gam, int i, const double *u) {
    const int j = (i < 3) ? i : (i == 3 ? 3 : (i < 7 ? i - 4 : i - 7));
    return Gam[i * 4 + j] * u[j];
}

__device__ __forceinline__ int sym3(int a, int b) { // a <= b in 0..2 -> 0..5
    return a * 3 - (a * (a - 1)) / 2 + (b - a);
}

struct WarpCtx {
    const SolveConsts *c; // in shared memory
    double *s;            // warp-private shared memory
    WarpLayout L;
    const double *prefix; // this instance's [x0 | ref | obst | target]
    int lane;
    __device__ WarpCtx(const SolveConsts *c_, double *s_, const double *p_, int lane_)
        : c(c_), s(s_), L(c_->N), prefix(p_), lane(lane_) {}
};

// Evaluate the objective at the trial point (x + dxt, u + dut) or at (x, u).
// lane = stage.  full: also writes q (grad x), r (grad u) and Hc (Hessian).
// Returns the warp-wide objective value (identical in all lanes).
template <bool FULL, bool TRIAL>
__device__ __noinline__ double eval_cost(const WarpCtx &w, double eps) {
    const SolveConsts &c = *w.c;
    const int N = c.N, K = c.K;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const double *qg = c.wgt, *qp = c.wgt + 10, *qu = c.wgt + 20;
    const double lam = c.wgt[24];
    double acc = 0.0;
    for (int kc = w.lane; kc < N; kc += 32) { // cost stage kc acts on U_kc and X_{kc+1}
        const int k = kc + 1;
        // control term (mpc_obstacle_casadi.py:209-210)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double uu = s[L.u + 4 * kc + i];
            if (TRIAL)
                uu += s[L.dut + 4 * kc + i];
            const double d = uu - (i == 2 ? AMPC_GZ : 0.0);
            acc += qu[i] * d * d;
            if (FULL)
                s[L.r + 4 * kc + i] = 2.0 * qu[i] * d;
        }
        double x[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            x[i] = s[L.x + 10 * k + i];
            if (TRIAL)
                x[i] += s[L.dxt + 10 * k + i];
        }
        if (kc == N - 1) { // terminal (mpc_obstacle_casadi.py:168-170)
            const double *tg = w.prefix + 10 + 10 * N + 3 * K * N;
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const double d = x[i] - tg[i];
                acc += qg[i] * d * d;
                if (FULL)
                    s[L.q + 10 * k + i] = 2.0 * qg[i] * d;
            }
            continue;
        }
        // path term in the yaw-rotated frame (mpc_obstacle_casadi.py:172-185,206-208)
        const double *ref = w.prefix + 10 + 10 * kc;
        const double cy = s[L.cs + 2 * kc], sy = s[L.cs + 2 * kc + 1]; // cos(yaw), sin(-yaw)
        double dl[10];
#pragma unroll
        for (int i = 0; i < 10; ++i)
            dl[i] = x[i] - ref[i];
        const double rd0 = cy * dl[0] - sy * dl[1], rd1 = sy * dl[0] + cy * dl[1];
        const double rd4 = cy * dl[4] - sy * dl[5], rd5 = sy * dl[4] + cy * dl[5];
        acc += qp[0] * rd0 * rd0 + qp[1] * rd1 * rd1 + qp[4] * rd4 * rd4 + qp[5] * rd5 * rd5;
        acc += qp[2] * dl[2] * dl[2] + qp[3] * dl[3] * dl[3] + qp[6] * dl[6] * dl[6] +
               qp[7] * dl[7] * dl[7] + qp[8] * dl[8] * dl[8] + qp[9] * dl[9] * dl[9];
        double g[10], hpp[6], hpv[9], hvv[6];
        if (FULL) {
            g[0] = 2.0 * (cy * qp[0] * rd0 + sy * qp[1] * rd1);
            g[1] = 2.0 * (-sy * qp[0] * rd0 + cy * qp[1] * rd1);
            g[4] = 2.0 * (cy * qp[4] * rd4 + sy * qp[5] * rd5);
            g[5] = 2.0 * (-sy * qp[4] * rd4 + cy * qp[5] * rd5);
            g[2] = 2.0 * qp[2] * dl[2];
            g[3] = 2.0 * qp[3] * dl[3];
            g[6] = 2.0 * qp[6] * dl[6];
            g[7] = 2.0 * qp[7] * dl[7];
            g[8] = 2.0 * qp[8] * dl[8];
            g[9] = 2.0 * qp[9] * dl[9];
            hpp[0] = 2.0 * (cy * cy * qp[0] + sy * sy * qp[1]);
            hpp[1] = 2.0 * (cy * sy * (qp[1] - qp[0]));
            hpp[2] = 0.0;
            hpp[3] = 2.0 * (sy * sy * qp[0] + cy * cy * qp[1]);
            hpp[4] = 0.0;
            hpp[5] = 2.0 * qp[2];
            hvv[0] = 2.0 * (cy * cy * qp[4] + sy * sy * qp[5]);
            hvv[1] = 2.0 * (cy * sy * (qp[5] - qp[4]));
            hvv[2] = 0.0;
            hvv[3] = 2.0 * (sy * sy * qp[4] + cy * cy * qp[5]);
            hvv[4] = 0.0;
            hvv[5] = 2.0 * qp[6];
#pragma unroll
            for (int i = 0; i < 9; ++i)
                hpv[i] = 0.0;
        }
        // collision terms (mpc_obstacle_casadi.py:186-204)
        const double *ob = w.prefix + 10 + 10 * N + 3 * K * kc;
        const double far2 = (c.radius + 1.25) * (c.radius + 1.25);
        for (int j = 0; j < K; ++j) {
            const double d0 = ob[3 * j] - x[0], d1 = ob[3 * j + 1] - x[1], d2 = ob[3 * j + 2] - x[2];
            const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
            // softplus argument below -40 (clearance > 1.25 m): exp < 4.3e-18, so the term is
            // < 1e-16*|v.n| in value and < 3e-15 in any derivative -- below the rounding of the
            // sums it would be added to (this also covers the (1e4,1e4,1e4) padding points)
            if (r2 > far2)
                continue;
            const double rr = sqrt(r2);
            const double ir = 1.0 / rr;
            const double n0 = d0 * ir, n1 = d1 * ir, n2 = d2 * ir;
            const double sv = x[4] * n0 + x[5] * n1 + x[6] * n2;
            const double e = exp((rr - c.radius) * -32.0);
            // log(1+e) and e/(1+e) equal e to within e^2 < 6e-17 when e < 2^-27: same accuracy
            // as the reference's un-stabilised log(1+exp(x)), whose 1+e rounds at 1.1e-16
            const bool tiny = e < 7.450580596923828e-09;
            const double sp = tiny ? e : log(1.0 + e);
            const double hyp = sqrt(sv * sv + eps * eps);
            const double psi = hyp - eps;
            acc += lam * sp * psi;
            if (FULL) {
                const double ih = 1.0 / hyp;
                const double dpsi = sv * ih;
                const double ddpsi = eps * eps * ih * ih * ih;
                const double sig = tiny ? e : e / (1.0 + e);
                const double w0 = (x[4] - sv * n0) * ir, w1 = (x[5] - sv * n1) * ir,
                             w2 = (x[6] - sv * n2) * ir;
                const double nn[3] = {n0, n1, n2}, ww[3] = {w0, w1, w2};
                const double a_n = lam * 32.0 * sig * psi; // grad p along n
                const double a_w = lam * sp * dpsi;        // grad p along -w, grad v along n
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    g[a] += a_n * nn[a] - a_w * ww[a];
                    g[4 + a] += a_w * nn[a];
                }
                // Hessian coefficients (SURVEY.md 8a closed forms, generalised to psi)
                const double c_nn = lam * 1024.0 * sig * (1.0 - sig) * psi;
                const double c_nw = lam * 32.0 * sig * dpsi;
                const double c_pi = lam * 32.0 * sig * psi * ir;
                const double c_ww = lam * sp * ddpsi;
                const double c_sp = lam * sp * dpsi * ir;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double Pi = (a == b ? 1.0 : 0.0) - nn[a] * nn[b];
                        if (b >= a)
                            hpp[sym3(a, b)] += c_nn * nn[a] * nn[b] -
                                               c_nw * (nn[a] * ww[b] + ww[a] * nn[b]) - c_pi * Pi +
                                               c_ww * ww[a] * ww[b] -
                                               c_sp * (nn[a] * ww[b] + sv * Pi * ir + ww[a] * nn[b]);
                        hpv[a * 3 + b] += c_nw * nn[a] * nn[b] - c_sp * Pi - c_ww * ww[a] * nn[b];
                        if (b >= a)
                            hvv[sym3(a, b)] += c_ww * nn[a] * nn[b];
                    }
            }
        }
        if (FULL) {
#pragma unroll
            for (int i = 0; i < 10; ++i)
                s[L.q + 10 * k + i] = g[i];
            double *Hc = s + L.Hc + 21 * kc;
#pragma unroll
            for (int i = 0; i < 6; ++i)
                Hc[i] = hpp[i];
#pragma unroll
            for (int i = 0; i < 9; ++i)
                Hc[6 + i] = hpv[i];
#pragma unroll
            for (int i = 0; i < 6; ++i)
                Hc[15 + i] = hvv[i];
        }
    }
    __syncwarp();
    return warp_sum(acc);
}

// Objective value only (line-search trials, final cost): the (N-1)*K collision terms are
// spread over all 32 lanes (term t -> lane t mod 32; consecutive lanes read consecutive
// obstacle points), the N control/path/terminal terms over lanes 0..N-1.
__device__ __noinline__ double eval_value(const WarpCtx &w, double eps, const bool TRIAL) {
    const SolveConsts &c = *w.c;
    const int N = c.N, K = c.K;
    const double *s = w.s;
    const WarpLayout &L = w.L;
    const double *qg = c.wgt, *qp = c.wgt + 10, *qu = c.wgt + 20;
    const double lam = c.wgt[24];
    double acc = 0.0;
    for (int kc = w.lane; kc < N; kc += 32) {
        const int k = kc + 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double uu = s[L.u + 4 * kc + i];
            if (TRIAL)
                uu += s[L.dut + 4 * kc + i];
            const double d = uu - (i == 2 ? AMPC_GZ : 0.0);
            acc += qu[i] * d * d;
        }
        double x[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            x[i] = s[L.x + 10 * k + i];
            if (TRIAL)
                x[i] += s[L.dxt + 10 * k + i];
        }
        if (kc == N - 1) {
            const double *tg = w.prefix + 10 + 10 * N + 3 * K * N;
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                const double d = x[i] - tg[i];
                acc += qg[i] * d * d;
            }
        } else {
            const double *ref = w.prefix + 10 + 10 * kc;
            const double cy = s[L.cs + 2 * kc], sy = s[L.cs + 2 * kc + 1];
            double dl[10];
#pragma unroll
            for (int i = 0; i < 10; ++i)
                dl[i] = x[i] - ref[i];
            const double rd0 = cy * dl[0] - sy * dl[1], rd1 = sy * dl[0] + cy * dl[1];
            const double rd4 = cy * dl[4] - sy * dl[5], rd5 = sy * dl[4] + cy * dl[5];
            acc += qp[0] * rd0 * rd0 + qp[1] * rd1 * rd1 + qp[4] * rd4 * rd4 + qp[5] * rd5 * rd5;
            acc += qp[2] * dl[2] * dl[2] + qp[3] * dl[3] * dl[3] + qp[6] * dl[6] * dl[6] +
                   qp[7] * dl[7] * dl[7] + qp[8] * dl[8] * dl[8] + qp[9] * dl[9] * dl[9];
        }
    }
    const double far2 = (c.radius + 1.25) * (c.radius + 1.25);
    const double *ob = w.prefix + 10 + 10 * N;
    const int n_terms = (N - 1) * K;
    int kc = 0, j = w.lane;
    while (j >= K) {
        j -= K;
        ++kc;
    }
    for (int t = w.lane; t < n_terms; t += 32) {
        const int k = kc + 1;
        double px = s[L.x + 10 * k], py = s[L.x + 10 * k + 1], pz = s[L.x + 10 * k + 2];
        double vx = s[L.x + 10 * k + 4], vy = s[L.x + 10 * k + 5], vz = s[L.x + 10 * k + 6];
        if (TRIAL) {
            px += s[L.dxt + 10 * k], py += s[L.dxt + 10 * k + 1], pz += s[L.dxt + 10 * k + 2];
            vx += s[L.dxt + 10 * k + 4], vy += s[L.dxt + 10 * k + 5], vz += s[L.dxt + 10 * k + 6];
        }
        const double d0 = ob[3 * t] - px, d1 = ob[3 * t + 1] - py, d2 = ob[3 * t + 2] - pz;
        const double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        if (r2 <= far2) { // same far-term rule as eval_cost
            const double rr = sqrt(r2);
            const double sv = (vx * d0 + vy * d1 + vz * d2) / rr;
            const double e = exp((rr - c.radius) * -32.0);
            const double sp = e < 7.450580596923828e-09 ? e : log(1.0 + e);
            acc += lam * sp * (sqrt(sv * sv + eps * eps) - eps);
        }
        j += 32;
        while (j >= K) {
            j -= K;
            ++kc;
        }
    }
    return warp_sum(acc);
}

// Riccati backward sweep + adjoint.  Lane 4i+j (i,j = chains, lanes 16..31 mirror 0..15)
// keeps the 3x3 block P^(ij) coupling chain i and chain j in registers; per stage
//   S_ij = G_i' P^(ij) G_j (+R_i),  Bm^(i)_j = F_i' P^(ij) G_j,  A^(ij) = F_i' P^(ij) F_j,
//   P^(ij) <- Q^(ij) + A^(ij) - Bm^(i) S^-1 Bm^(j)',
// with S gathered by shuffles (it is 3x3 + the decoupled yaw scalar) and factored as LDL'
// redundantly in every lane.  Returns false if a pivot of S is not positive (the reduced
// Hessian has the wrong inertia) -> the caller regularises.  Also returns the dual
// infeasibility |r_k + G'lam_{k+1} - zl + zu|_inf.
// Dual infeasibility |grad_U L|_inf of the current iterate: the costate recursion
// lam_k = q_k + Phi' lam_{k+1}, grad_{U_k} L = r_k - zl_k + zu_k + Gam' lam_{k+1}, chain by
// chain (lane -> chain (lane >> 2) & 3, computed redundantly).  This is all the convergence
// test and the mu rule need, so they run BEFORE the Newton system is factored: one Riccati
// sweep per iteration whether or not mu changes, none for the final test.
__device__ __noinline__ double adjoint_dual_inf(const WarpCtx &w) {
    const SolveConsts &c = *w.c;
    const int N = c.N;
    const double *s = w.s;
    const WarpLayout &L = w.L;
    const int ci = (w.lane & 15) >> 2;
    const Chain fi = load_chain(c.Phi, c.Gam, ci);
    const int si0 = chain_state(ci, 0), si1 = chain_state(ci, 1), si2 = chain_state(ci, 2);
    const int q1 = ci < 3 ? si1 : si0, q2 = ci < 3 ? si2 : si0; // padding reads a valid slot,
    const double m_ax = ci < 3 ? 1.0 : 0.0;                     // masked to zero
    double lv[3];
    lv[0] = s[L.q + 10 * N + si0];
    lv[1] = m_ax * s[L.q + 10 * N + q1];
    lv[2] = m_ax * s[L.q + 10 * N + q2];
    double e_dual = 0.0;
#pragma unroll 2
    for (int k = N - 1; k >= 0; --k) {
        const double guk = s[L.r + 4 * k + ci] - s[L.zl + 4 * k + ci] + s[L.zu + 4 * k + ci];
        const double gu = guk + fi.g1 * lv[0] + fi.g2 * lv[1] + fi.g3 * lv[2];
        e_dual = fmax(e_dual, fabs(gu));
        if (k == 0)
            break;
        const double qk0 = s[L.q + 10 * k + si0];
        const double qk1 = m_ax * s[L.q + 10 * k + q1];
        const double qk2 = m_ax * s[L.q + 10 * k + q2];
        double fl[3];
        v1::chain_FT(fi, lv, fl);
        lv[0] = qk0 + fl[0], lv[1] = qk1 + fl[1], lv[2] = qk2 + fl[2];
    }
    __syncwarp(); // q, r may be rewritten by the next evaluation
    return warp_max(e_dual);
}

__device__ __noinline__ bool riccati_backward(const WarpCtx &w, double delta) {
    const SolveConsts &c = *w.c;
    const int N = c.N;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const int pl = w.lane & 15, ci = pl >> 2, cj = pl & 3;
    const Chain fi = load_chain(c.Phi, c.Gam, ci), fj = load_chain(c.Phi, c.Gam, cj);
    const double *qg = c.wgt, *qp = c.wgt + 10;
    const int si0 = chain_state(ci, 0), si1 = chain_state(ci, 1), si2 = chain_state(ci, 2);
    const bool diag = ci == cj, real = ci < 3 && cj < 3, yaw = pl == 15;
    const int lo = ci < cj ? ci : cj, hi = ci < cj ? cj : ci;
    // branch-free access to this lane's entries of the stage Hessian: offsets into the
    // compact store and 0/1 masks (constant over the sweep)
    const double m_real = real ? 1.0 : 0.0, m_diag = diag ? 1.0 : 0.0;
    const int o_pp = real ? sym3(lo, hi) : 0, o_pv = real ? 6 + ci * 3 + cj : 0;
    const int o_vp = real ? 6 + cj * 3 + ci : 0, o_vv = real ? 15 + sym3(lo, hi) : 0;
    const double q00c = yaw ? 2.0 * qp[3] + delta : (real && diag ? delta : 0.0);
    const double q11c = real && diag ? delta : 0.0;
    const double q22c = real && diag ? 2.0 * qp[si2] + delta : 0.0;
    const int q1 = ci < 3 ? si1 : si0, q2 = ci < 3 ? si2 : si0; // padding reads a valid slot,
    const double m_ax = ci < 3 ? 1.0 : 0.0;                     // masked to zero
    // terminal: P_N = diag(2 Q_goal) + delta, p_N = q_N
    double P[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double pv[3];
    if (diag) {
        P[0] = 2.0 * qg[si0] + delta;
        if (ci < 3) {
            P[4] = 2.0 * qg[si1] + delta;
            P[8] = 2.0 * qg[si2] + delta;
        }
    }
    pv[0] = s[L.q + 10 * N + si0];
    pv[1] = m_ax * s[L.q + 10 * N + q1];
    pv[2] = m_ax * s[L.q + 10 * N + q2];
    bool ok = true;
    for (int k = N - 1; k >= 0; --k) {
        // operands from shared memory first (latency overlaps the products)
        const double rdk = s[L.rdiag + 4 * k + ci], rtk = s[L.rt + 4 * k + ci];
        double Qb0 = q00c, Qb1 = 0.0, Qb3 = 0.0, Qb4 = q11c, qk0 = 0.0, qk1 = 0.0, qk2 = 0.0;
        if (k > 0) {
            const double *Hc = s + L.Hc + 21 * (k - 1);
            Qb0 += m_real * Hc[o_pp];
            Qb1 = m_real * Hc[o_pv];
            Qb3 = m_real * Hc[o_vp];
            Qb4 += m_real * Hc[o_vv];
            qk0 = s[L.q + 10 * k + si0];
            qk1 = m_ax * s[L.q + 10 * k + q1];
            qk2 = m_ax * s[L.q + 10 * k + q2];
        }
        // (1) products with this lane's block
        double t[3], bm[3], M[9], A[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            t[a] = P[3 * a] * fj.g1 + P[3 * a + 1] * fj.g2 + P[3 * a + 2] * fj.g3;
        const double Sij = fi.g1 * t[0] + fi.g2 * t[1] + fi.g3 * t[2] + m_diag * (rdk + delta);
        v1::chain_FT(fi, t, bm);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            M[3 * a] = fj.d1 * P[3 * a];
            M[3 * a + 1] = fj.c1 * P[3 * a] + fj.d2 * P[3 * a + 1];
            M[3 * a + 2] = fj.c2 * P[3 * a] + fj.c3 * P[3 * a + 1] + fj.c4 * P[3 * a + 2];
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            A[b] = fi.d1 * M[b];
            A[3 + b] = fi.c1 * M[b] + fi.d2 * M[3 + b];
            A[6 + b] = fi.c2 * M[b] + fi.c3 * M[3 + b] + fi.c4 * M[6 + b];
        }
        const double bi = rtk + fi.g1 * pv[0] + fi.g2 * pv[1] + fi.g3 * pv[2];
        // (2) S = L D L' (3x3 for the axes; yaw is a decoupled scalar), every lane redundantly
        const double S00 = __shfl_sync(AMPC_FULL_MASK, Sij, 0), S10 = __shfl_sync(AMPC_FULL_MASK, Sij, 4);
        const double S20 = __shfl_sync(AMPC_FULL_MASK, Sij, 8), S11 = __shfl_sync(AMPC_FULL_MASK, Sij, 5);
        const double S21 = __shfl_sync(AMPC_FULL_MASK, Sij, 9), S22 = __shfl_sync(AMPC_FULL_MASK, Sij, 10);
        const double S33 = __shfl_sync(AMPC_FULL_MASK, Sij, 15);
        const double d0 = S00, i0 = 1.0 / d0, i3 = 1.0 / S33;
        const double l10 = S10 * i0, l20 = S20 * i0;
        const double d1 = S11 - l10 * l10 * d0, i1 = 1.0 / d1;
        const double l21 = (S21 - l20 * l10 * d0) * i1;
        const double d2 = S22 - l20 * l20 * d0 - l21 * l21 * d1, i2 = 1.0 / d2;
        if (!(d0 > 0.0) || !(d1 > 0.0) || !(d2 > 0.0) || !(S33 > 0.0)) {
            ok = false;
            break; // uniform: every lane computed the same pivots
        }
        // feed-forward kff = -S^-1 b  (b_l lives in lane 5l; the yaw entry stays in lane 15)
        double kff0, kff1, kff2;
        const double kff3 = -bi * i3; // meaningful in lane 15 only
        {
            const double b0 = -__shfl_sync(AMPC_FULL_MASK, bi, 0), b1 = -__shfl_sync(AMPC_FULL_MASK, bi, 5);
            const double b2 = -__shfl_sync(AMPC_FULL_MASK, bi, 10);
            const double w1 = b1 - l10 * b0, w2 = b2 - l20 * b0 - l21 * w1;
            kff2 = w2 * i2;
            kff1 = w1 * i1 - l21 * kff2;
            kff0 = b0 * i0 - l10 * kff1 - l20 * kff2;
        }
        if (w.lane == 0) {
            s[L.kf + 4 * k + 0] = kff0;
            s[L.kf + 4 * k + 1] = kff1;
            s[L.kf + 4 * k + 2] = kff2;
        }
        if (w.lane == 15)
            s[L.kf + 4 * k + 3] = kff3;
        if (k == 0)
            break; // dx_0 = 0: no feedback gain and no P_0 needed
        // (3) Y = S^-1 Bm^(j)' (axis controls x components of chain j) and Bm^(i).  All
        // couplings between the yaw chain and the axes are exactly zero, so only the three
        // axis controls are gathered; the yaw block uses its own values.
        double Y[3][3], Bi[3][3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            const double r0 = __shfl_sync(AMPC_FULL_MASK, bm[cc], 4 * cj + 0);
            const double r1 = __shfl_sync(AMPC_FULL_MASK, bm[cc], 4 * cj + 1);
            const double r2 = __shfl_sync(AMPC_FULL_MASK, bm[cc], 4 * cj + 2);
            const double w1 = r1 - l10 * r0, w2 = r2 - l20 * r0 - l21 * w1;
            Y[2][cc] = w2 * i2;
            Y[1][cc] = w1 * i1 - l21 * Y[2][cc];
            Y[0][cc] = r0 * i0 - l10 * Y[1][cc] - l20 * Y[2][cc];
#pragma unroll
            for (int l = 0; l < 3; ++l)
                Bi[cc][l] = __shfl_sync(AMPC_FULL_MASK, bm[cc], 4 * ci + l);
        }
        // feedback gains K = -Y: Kg[k][chain j][control l][component], written by lanes (0, j);
        // the yaw gain -bm[0]/S33 by lane 15
        if (w.lane < 3) {
            double *Kg = s + L.Kg + 48 * k + 12 * cj;
#pragma unroll
            for (int l = 0; l < 3; ++l)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    Kg[3 * l + cc] = -Y[l][cc];
        }
        if (w.lane == 15)
            s[L.Kg + 48 * k + 36 + 9] = -bm[0] * i3; // the never-written slots stay zero (cleared at start)
        // (4) P^(ij) <- Q^(ij) + A - Bm^(i) Y ;  p^(i) <- q^(i) + F_i' p^(i) + Bm^(i) kff
        const double yy = yaw ? i3 : 0.0; // yaw block: - bm bm' / S33
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                P[3 * a + b] = A[3 * a + b] - (Bi[a][0] * Y[0][b] + Bi[a][1] * Y[1][b] + Bi[a][2] * Y[2][b]) -
                               yy * bm[a] * bm[b];
        P[0] += Qb0, P[1] += Qb1, P[3] += Qb3, P[4] += Qb4, P[8] += q22c;
        double fp[3];
        v1::chain_FT(fi, pv, fp);
        const double kq[3] = {qk0, qk1, qk2};
        const double ky = yaw ? kff3 : 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            pv[a] = kq[a] + fp[a] + Bi[a][0] * kff0 + Bi[a][1] * kff1 + Bi[a][2] * kff2 + ky * bm[a];
        }
    }
    __syncwarp();
    return __all_sync(AMPC_FULL_MASK, ok);
}

// forward sweep: du_k = K_k dx_k + kff_k, dx_{k+1} = Phi dx_k + Gam du_k, dx_0 = 0.
// Lane 4i+j carries dx of chain i and of chain j.
__device__ __noinline__ void riccati_forward(const WarpCtx &w) {
    const SolveConsts &c = *w.c;
    const int N = c.N;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const int pl = w.lane & 15, ci = pl >> 2, cj = pl & 3;
    const Chain fi = load_chain(c.Phi, c.Gam, ci), fj = load_chain(c.Phi, c.Gam, cj);
    const int si0 = chain_state(ci, 0), si1 = chain_state(ci, 1), si2 = chain_state(ci, 2);
    double xi[3] = {0, 0, 0}, xj[3] = {0, 0, 0};
    if (w.lane < 10)
        s[L.dx + w.lane] = 0.0;
    for (int k = 0; k < N; ++k) {
        double part = 0.0;
        if (k > 0) { // Kg[k][chain j][control i][component]; axis/yaw cross gains are zero
            const double *Kg = s + L.Kg + 48 * k + 12 * cj + 3 * ci;
            const double m = ((ci < 3) == (cj < 3)) ? 1.0 : 0.0;
            part = m * (Kg[0] * xj[0] + Kg[1] * xj[1] + Kg[2] * xj[2]);
        }
        part += __shfl_xor_sync(AMPC_FULL_MASK, part, 1);
        part += __shfl_xor_sync(AMPC_FULL_MASK, part, 2);
        const double ui = s[L.kf + 4 * k + ci] + part;
        const double uj = __shfl_sync(AMPC_FULL_MASK, ui, 4 * cj);
        double yi[3], yj[3];
        chain_F(fi, xi, ui, yi);
        chain_F(fj, xj, uj, yj);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            xi[a] = yi[a];
            xj[a] = yj[a];
        }
        if (w.lane < 16 && cj == 0) {
            s[L.du + 4 * k + ci] = ui;
            s[L.dx + 10 * (k + 1) + si0] = xi[0];
            if (ci < 3) {
                s[L.dx + 10 * (k + 1) + si1] = xi[1];
                s[L.dx + 10 * (k + 1) + si2] = xi[2];
            }
        }
    }
    __syncwarp();
}

// dxt = linear roll-out of dut (dxt_0 = 0, dxt_{k+1} = Phi dxt_k + Gam dut_k); lane i < 4 = chain i
__device__ __noinline__ void rollout_delta(const WarpCtx &w) {
    const SolveConsts &c = *w.c;
    const int N = c.N;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const int ci = w.lane & 3;
    const Chain fi = load_chain(c.Phi, c.Gam, ci);
    const int si0 = chain_state(ci, 0), si1 = chain_state(ci, 1), si2 = chain_state(ci, 2);
    double xi[3] = {0, 0, 0};
    if (w.lane < 10)
        s[L.dxt + w.lane] = 0.0;
    for (int k = 0; k < N; ++k) {
        double y[3];
        chain_F(fi, xi, s[L.dut + 4 * k + ci], y);
        xi[0] = y[0], xi[1] = y[1], xi[2] = y[2];
        if (w.lane < 4) {
            s[L.dxt + 10 * (k + 1) + si0] = xi[0];
            if (ci < 3) {
                s[L.dxt + 10 * (k + 1) + si1] = xi[1];
                s[L.dxt + 10 * (k + 1) + si2] = xi[2];
            }
        }
    }
    __syncwarp();
}

// has_work = false: a warp of the CTA without an instance; it only keeps the CTA's per-iteration
// barrier company (every warp of a CTA executes that ONE barrier the same number of times)
template <bool SYNC>
__device__ void solve_instance(const WarpCtx &w, double *w_inout, SolveOut *out, bool has_work) {
    const SolveConsts &c = *w.c;
    const int N = c.N, lane = w.lane, nu = 4 * N;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const double kappa_eps = c.kappa_eps, kappa_mu = 0.2, tau_min = 0.99;
    const double eta = 1e-4, kappa_sigma = 1e10;
    const double mu_min = c.tol / 10.0;
    double mu = c.mu_init, delta_last = 0.0;
    int n_reg = 0, n_bt = 0;

    // controls from the warm start pushed into the box interior; cos/sin of the ref yaw
    for (int e = lane; has_work && e < nu; e += 32) {
        const int k = e >> 2, i = e & 3;
        const double lo = c.lb[i], hi = c.ub[i];
        double pl = fmin(c.bound_push * fmax(1.0, fabs(lo)), c.bound_frac * (hi - lo));
        double pu = fmin(c.bound_push * fmax(1.0, fabs(hi)), c.bound_frac * (hi - lo));
        double uu = w_inout[14 * k + 10 + i];
        uu = fmin(fmax(uu, lo + pl), hi - pu);
        s[L.u + e] = uu;
        s[L.zl + e] = mu / (uu - lo);
        s[L.zu + e] = mu / (hi - uu);
        s[L.du + e] = 0.0;
    }
    for (int e = lane; e < 48 * N; e += 32)
        s[L.Kg + e] = 0.0;
    for (int k = lane; has_work && k < N; k += 32) {
        const double yaw = w.prefix[10 + 10 * k + 3];
        s[L.cs + 2 * k] = cos(yaw);
        s[L.cs + 2 * k + 1] = sin(-yaw);
    }
    if (lane < 10 && has_work)
        s[L.x + lane] = w.prefix[lane];
    __syncwarp();
    for (int k = 0; has_work && k < N; ++k) { // roll-out
        if (lane < 10)
            s[L.x + 10 * (k + 1) + lane] = c.gam[lane] + phi_row_dot(c.Phi, lane, s + L.x + 10 * k) +
                                           gam_row_dot(c.Gam, lane, s + L.u + 4 * k);
        __syncwarp();
    }

    int status = 1, iter = 0;
    double e_dual = 0.0, e_compl = 0.0;
    bool done = !has_work;
    for (iter = 0;;) {
        // the warps of a CTA start every iteration together, so that they run the same code
        // region at the same time and share the instruction cache (the kernel's code is far
        // larger than it).  A warp that has finished (or never had an instance) keeps arriving
        // at this one barrier, voting "done", until every warp of the CTA votes "done".
        if (SYNC) {
            if (!__syncthreads_or(done ? 0 : 1))
                break;
        } else if (done) {
            break;
        }
        if (done)
            continue;
        const bool finished = [&]() -> bool {
        double eps = 0.0, f = 0.0, c_mu = 0.0, delta = 0.0;
        bool stop = false;
        // pass 0: evaluate, test convergence, update mu (IPOPT eq. (7)); if mu changed, the
        // barrier terms changed, and so did the smoothing unless it sits at its floor: pass 1
        // refreshes what depends on them.  Then ONE Newton system per iteration.
        double eps_at = -1.0;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            eps = fmax(c.eps_min, c.eps_scale * mu);
            if (eps != eps_at) {
                f = eval_cost<true, false>(w, eps);
                eps_at = eps;
            }
            double ec = 0.0, cm = 0.0;
#pragma unroll 1
            for (int e = lane; e < nu; e += 32) {
                const int i = e & 3;
                const double sl = s[L.u + e] - c.lb[i], su = c.ub[i] - s[L.u + e];
                const double isl = 1.0 / sl, isu = 1.0 / su;
                const double zl = s[L.zl + e], zu = s[L.zu + e];
                ec = fmax(ec, fmax(sl * zl, su * zu));
                cm = fmax(cm, fmax(fabs(sl * zl - mu), fabs(su * zu - mu)));
                s[L.rdiag + e] = 2.0 * c.wgt[20 + i] + zl * isl + zu * isu;
                s[L.rt + e] = s[L.r + e] - mu * isl + mu * isu;
            }
            __syncwarp();
            if (pass == 1)
                break;
            e_dual = adjoint_dual_inf(w);
            e_compl = warp_max(ec);
            c_mu = warp_max(cm);
            if (!(f == f) || !(e_dual == e_dual)) {
                status = 3;
                stop = true;
            } else if (fmax(e_dual, e_compl) <= c.tol) {
                status = 0;
                stop = true;
            } else if (iter >= c.max_iter) {
                status = 1;
                stop = true;
            }
            if (stop)
                break;
            bool mu_changed = false;
#pragma unroll 1
            while (mu > mu_min && fmax(e_dual, c_mu) <= kappa_eps * mu) {
                mu = fmax(mu_min, fmin(kappa_mu * mu, mu * sqrt(mu))); // theta_mu = 1.5
                mu_changed = true;
                double c2 = 0.0;
#pragma unroll 1
                for (int e = lane; e < nu; e += 32) {
                    const int i = e & 3;
                    const double sl = s[L.u + e] - c.lb[i], su = c.ub[i] - s[L.u + e];
                    c2 = fmax(c2, fmax(fabs(sl * s[L.zl + e] - mu), fabs(su * s[L.zu + e] - mu)));
                }
                c_mu = warp_max(c2);
            }
            if (!mu_changed)
                break;
        }
        if (stop)
            return true;
        // Newton system with inertia correction (IPOPT Algorithm IC schedule)
        {
            int ntry = 0;
#pragma unroll 1
            while (!riccati_backward(w, delta)) {
                if (delta == 0.0)
                    delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
                else
                    delta *= (delta_last == 0.0) ? 100.0 : 8.0;
                if (++ntry > 60 || delta > 1e40) {
                    status = 3;
                    stop = true;
                    break;
                }
            }
        }
        if (stop)
            return true;
        if (delta > 0.0) {
            delta_last = delta;
            ++n_reg;
        }
        riccati_forward(w);
        const double tau_f = fmax(tau_min, 1.0 - mu);
        // fraction to the boundary (IPOPT eq. (15)), barrier value, slope
        double a_du = 1.0, bar0 = 0.0;
#pragma unroll 1
        for (int e = lane; e < nu; e += 32) {
            const int i = e & 3;
            const double uu = s[L.u + e], du = s[L.du + e];
            const double sl = uu - c.lb[i], su = c.ub[i] - uu;
            const double zl = s[L.zl + e], zu = s[L.zu + e];
            const double isl = 1.0 / sl, isu = 1.0 / su;
            const double dzl = mu * isl - zl - zl * isl * du;
            const double dzu = mu * isu - zu + zu * isu * du;
            if (dzl < 0.0)
                a_du = fmin(a_du, -tau_f * zl / dzl);
            if (dzu < 0.0)
                a_du = fmin(a_du, -tau_f * zu / dzu);
            bar0 += log(sl) + log(su);
        }
        a_du = warp_min(a_du);
        bar0 = warp_sum(bar0);
        const double phi0 = f - mu * bar0;
        // Projected Armijo line search on the barrier objective: the step is limited PER
        // COMPONENT by the fraction-to-the-boundary rule (a control that would cross its
        // bound stops at (1-tau)*slack from it) instead of scaling the whole step by the most
        // restrictive component; the states follow by linearity.  Every bound that wants to
        // become active is reached in one iteration.
        double alpha = 1.0;
        bool accepted = false;
#pragma unroll 1
        for (int ls = 0; ls < 40; ++ls) {
            double bar = 0.0, gdt = 0.0;
            __syncwarp(); // the previous trial's reads of dut / dxt are complete
#pragma unroll 1
            for (int e = lane; e < nu; e += 32) {
                const int i = e & 3;
                const double uu = s[L.u + e];
                const double sl = uu - c.lb[i], su = c.ub[i] - uu;
                double d = alpha * s[L.du + e];
                d = fmin(fmax(d, -tau_f * sl), tau_f * su);
                s[L.dut + e] = d;
                gdt += s[L.rt + e] * d;
                bar += log(sl + d) + log(su - d);
            }
            __syncwarp();
            rollout_delta(w);
            for (int e = 10 + lane; e < 10 * (N + 1); e += 32)
                gdt += s[L.q + e] * s[L.dxt + e];
            bar = warp_sum(bar);
            gdt = warp_sum(gdt);
            const double phit = eval_value(w, eps, true) - mu * bar;
            if (phit <= phi0 + eta * gdt + 10.0 * 2.220446049250313e-16 * fabs(phi0)) {
                accepted = true;
                break;
            }
            alpha *= 0.5;
            ++n_bt;
        }
        if (!accepted) {
            status = 2;
            return true;
        }
        for (int e = lane; e < nu; e += 32) {
            const int i = e & 3;
            const double u0 = s[L.u + e], du = s[L.du + e];
            const double sl0 = u0 - c.lb[i], su0 = c.ub[i] - u0;
            double zl = s[L.zl + e], zu = s[L.zu + e];
            const double isl0 = 1.0 / sl0, isu0 = 1.0 / su0;
            const double dzl = mu * isl0 - zl - zl * isl0 * du;
            const double dzu = mu * isu0 - zu + zu * isu0 * du;
            const double un = u0 + s[L.dut + e];
            const double sl = un - c.lb[i], su = c.ub[i] - un;
            const double mil = mu / sl, miu = mu / su;
            zl += a_du * dzl;
            zu += a_du * dzu;
            zl = fmax(fmin(zl, kappa_sigma * mil), mil * (1.0 / kappa_sigma)); // IPOPT eq. (16)
            zu = fmax(fmin(zu, kappa_sigma * miu), miu * (1.0 / kappa_sigma));
            s[L.u + e] = un;
            s[L.zl + e] = zl;
            s[L.zu + e] = zu;
        }
        for (int e = 10 + lane; e < 10 * (N + 1); e += 32)
            s[L.x + e] += s[L.dxt + e];
        __syncwarp();
        return false;
        }();
        if (finished)
            done = true;
        else
            ++iter;
    }
    if (!has_work)
        return;
    // results: w = [X_0,U_0,...,X_N]; objective without smoothing
    const double cost = eval_value(w, 0.0, false);
    for (int e = lane; e < 10 * (N + 1); e += 32) {
        const int k = e / 10, i = e - 10 * k;
        w_inout[14 * k + i] = s[L.x + e];
    }
    for (int e = lane; e < nu; e += 32)
        w_inout[14 * (e >> 2) + 10 + (e & 3)] = s[L.u + e];
    if (lane == 0) {
        out->cost = cost;
        out->kkt_dual = e_dual;
        out->kkt_compl = e_compl;
        out->mu = mu;
        out->iters = iter;
        out->status = status;
        out->n_reg = n_reg;
        out->n_backtrack = n_bt;
    }
}

template <int WARPS>
#ifndef AMPC_SOLVE_MIN_BLOCKS
#define AMPC_SOLVE_MIN_BLOCKS 1 // occupancy is bounded by the ~246 registers the Riccati blocks need
#endif
__global__ void __launch_bounds__(WARPS * 32, AMPC_SOLVE_MIN_BLOCKS)
ipm_solve_kernel(const __grid_constant__ SolveConsts consts, int B, const double *__restrict__ prefix,
                 double *__restrict__ w_inout, SolveOut *__restrict__ info,
                 const int32_t *__restrict__ active /* nullable: instances with 0 are skipped */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SolveConsts *sc = reinterpret_cast<SolveConsts *>(smem_raw);
    double *wbase = reinterpret_cast<double *>(smem_raw + ((sizeof(SolveConsts) + 127) / 128) * 128);
    {
        const double *src = reinterpret_cast<const double *>(&consts);
        double *dst = reinterpret_cast<double *>(sc);
        for (int i = threadIdx.x; i < (int)(sizeof(SolveConsts) / 8); i += WARPS * 32)
            dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * WARPS + warp;
    const bool has_work = !(b >= B || (active && active[b] == 0));
    if (WARPS == 1 && !has_work)
        return;
    const int bb = has_work ? b : 0; // a warp without work only attends the CTA's barrier
    const WarpLayout L(sc->N);
    WarpCtx ctx(sc, wbase + (size_t)warp * L.total, prefix + (size_t)bb * sc->n_prefix, lane);
    solve_instance<(WARPS > 1)>(ctx, w_inout + (size_t)bb * (10 + 14 * sc->N), info + bb, has_work);
}

} // namespace v1
} // namespace ampc
