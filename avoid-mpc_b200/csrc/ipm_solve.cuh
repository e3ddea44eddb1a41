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
// Method (see DESIGN.md "Solver"): primal-dual log-barrier on the control box
// with IPOPT's monotone mu rule, fraction-to-the-boundary rule, inertia-correction
// schedule and multiplier safeguard; iterates stay on the dynamics manifold
// (X = roll-out of U), the Newton system of the multiple-shooting NLP is solved
// exactly by a stage-wise Riccati sweep (nx = 10, nu = 4) exploiting the chain
// sparsity of Phi/Gam; Armijo backtracking on the barrier objective.  |v.n| is
// smoothed inside the solver as sqrt(s^2+eps^2)-eps with eps = max(eps_min, mu).
//
// Lane mapping.  Cost/gradient/Hessian evaluation: lane = stage (all N stages
// in parallel, K obstacle terms accumulated in registers, no reduction except
// the objective value).  Riccati sweep: the 32 lanes share the entries of the
// 10x10 / 10x4 / 4x4 stage matrices, which live in warp-private shared memory.
#pragma once
#include "common.cuh"

namespace ampc {

#define AMPC_GZ 9.81 // tools/mpc_obstacle_casadi.py:39

// per-warp shared-memory layout, in doubles
struct WarpLayout {
    int x, u, zl, zu, dx, du, q, r, rt, rdiag, Hc, Kg, kf, P, PA, PB, Bm, S, pv, lam, b, cs, total;
    __host__ __device__ explicit WarpLayout(int N) {
        int o = 0;
        x = o, o += (N + 1) * 10;
        u = o, o += N * 4;
        zl = o, o += N * 4;
        zu = o, o += N * 4;
        dx = o, o += (N + 1) * 10;
        du = o, o += N * 4;
        q = o, o += (N + 1) * 10;
        r = o, o += N * 4;
        rt = o, o += N * 4;
        rdiag = o, o += N * 4;
        Hc = o, o += N * 21; // stage k (1..N-1) at (k-1)*21: pp(6) pv(9) vv(6)
        Kg = o, o += N * 40;
        kf = o, o += N * 4;
        P = o, o += 100;
        PA = o, o += 100;
        PB = o, o += 40;
        Bm = o, o += 40;
        S = o, o += 16;
        pv = o, o += 10;
        lam = o, o += 10;
        b = o, o += 4;
        cs = o, o += N * 2;
        total = (o + 1) & ~1;
    }
};

__host__ __device__ inline size_t solve_smem_bytes(int N, int warps) {
    return sizeof(SolveConsts) + 128 + (size_t)warps * WarpLayout(N).total * sizeof(double);
}

// ---- sparse products with the chain-structured Phi (10x10) and Gam (10x4).
// State order [p(0..2), yaw(3), v(4..6), a(7..9)].  Column j of Phi has
// non-zeros at rows {j} (+ {j-4} for v, + {j-3, j-7} for a); column j of Gam at
// rows {j, 4+j, 7+j} (j < 3) or {3} (j = 3).  (mpc_obstacle_casadi.py:106-122)

// sum_l Phi[l][i] * v[l*stride]
__device__ __forceinline__ double phiT_dot(const double *Phi, int i, const double *v, int stride) {
    double a = Phi[i * 10 + i] * v[i * stride];
    if (i >= 4 && i <= 6)
        a += Phi[(i - 4) * 10 + i] * v[(i - 4) * stride];
    if (i >= 7) {
        a += Phi[(i - 3) * 10 + i] * v[(i - 3) * stride];
        a += Phi[(i - 7) * 10 + i] * v[(i - 7) * stride];
    }
    return a;
}
// sum_l row[l] * Phi[l][j]
__device__ __forceinline__ double dot_phi(const double *Phi, const double *row, int j) {
    double a = row[j] * Phi[j * 10 + j];
    if (j >= 4 && j <= 6)
        a += row[j - 4] * Phi[(j - 4) * 10 + j];
    if (j >= 7) {
        a += row[j - 3] * Phi[(j - 3) * 10 + j];
        a += row[j - 7] * Phi[(j - 7) * 10 + j];
    }
    return a;
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
// sum_l Gam[l][j] * v[l*stride]
__device__ __forceinline__ double gamT_dot(const double *Gam, int j, const double *v, int stride) {
    if (j == 3)
        return Gam[3 * 4 + 3] * v[3 * stride];
    return Gam[j * 4 + j] * v[j * stride] + Gam[(4 + j) * 4 + j] * v[(4 + j) * stride] +
           Gam[(7 + j) * 4 + j] * v[(7 + j) * stride];
}
// sum_l Gam[i][l] * u[l]   (row i of Gam has one non-zero)
__device__ __forceinline__ double gam_row_dot(const double *Gam, int i, const double *u) {
    const int j = (i < 3) ? i : (i == 3 ? 3 : (i < 7 ? i - 4 : i - 7));
    return Gam[i * 4 + j] * u[j];
}

__device__ __forceinline__ int sym3(int a, int b) { // a <= b in 0..2 -> 0..5
    return a * 3 - (a * (a - 1)) / 2 + (b - a);
}

// Hessian entry (i <= j) of stage k's cost from the compact (p,v)-block store.
__device__ __forceinline__ double stage_hess(const double *Hc, const double *qp, int i, int j) {
    const bool ip = i < 3, iv = (i >= 4 && i <= 6);
    const bool jp = j < 3, jv = (j >= 4 && j <= 6);
    if (ip && jp)
        return Hc[sym3(i, j)];
    if (ip && jv)
        return Hc[6 + i * 3 + (j - 4)];
    if (iv && jv)
        return Hc[15 + sym3(i - 4, j - 4)];
    return (i == j) ? 2.0 * qp[i] : 0.0;
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

// Evaluate the objective at (x + alpha*dx, u + alpha*du) [trial] or at (x, u).
// lane = stage.  full: also writes q (grad x), r (grad u) and Hc (Hessian).
// Returns the warp-wide objective value (identical in all lanes).
template <bool FULL, bool TRIAL>
__device__ double eval_cost(const WarpCtx &w, double alpha, double eps) {
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
                uu += alpha * s[L.du + 4 * kc + i];
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
                x[i] += alpha * s[L.dx + 10 * k + i];
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
        for (int j = 0; j < K; ++j) {
            const double d0 = ob[3 * j] - x[0], d1 = ob[3 * j + 1] - x[1], d2 = ob[3 * j + 2] - x[2];
            const double rr = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double ir = 1.0 / rr;
            const double n0 = d0 * ir, n1 = d1 * ir, n2 = d2 * ir;
            const double sv = x[4] * n0 + x[5] * n1 + x[6] * n2;
            const double e = exp((rr - c.radius) * -32.0);
            const double sp = log(1.0 + e);
            const double hyp = sqrt(sv * sv + eps * eps);
            const double psi = hyp - eps;
            acc += lam * sp * psi;
            if (FULL) {
                const double ih = 1.0 / hyp;
                const double dpsi = sv * ih;
                const double ddpsi = eps * eps * ih * ih * ih;
                const double sig = e / (1.0 + e);
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

// Riccati backward sweep + adjoint; returns false if some S_k is not positive
// definite (the reduced Hessian has the wrong inertia) -> caller regularises.
// Also returns the dual infeasibility |r_k + Gam'lam_{k+1} - zl + zu|_inf.
__device__ bool riccati_backward(const WarpCtx &w, double delta, double *e_dual_out) {
    const SolveConsts &c = *w.c;
    const int N = c.N, lane = w.lane;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const double *Phi = c.Phi, *Gam = c.Gam;
    const double *qg = c.wgt, *qp = c.wgt + 10;
    double *P = s + L.P, *PA = s + L.PA, *PB = s + L.PB, *Bm = s + L.Bm, *S = s + L.S;
    double *pv = s + L.pv, *lam = s + L.lam, *bb = s + L.b;
    // terminal: P_N = diag(2 Q_goal) + delta, p_N = lam_N = q_N
    for (int e = lane; e < 100; e += 32) {
        const int i = e / 10, j = e - 10 * i;
        P[e] = (i == j) ? 2.0 * qg[i] + delta : 0.0;
    }
    if (lane < 10) {
        pv[lane] = s[L.q + 10 * N + lane];
        lam[lane] = pv[lane];
    }
    __syncwarp();
    double e_dual = 0.0;
    bool ok = true;
    for (int k = N - 1; k >= 0; --k) {
        // (1) PA = P Phi, PB = P Gam
        for (int e = lane; e < 100; e += 32) {
            const int i = e / 10, j = e - 10 * i;
            PA[e] = dot_phi(Phi, P + 10 * i, j);
        }
        for (int e = lane; e < 40; e += 32) {
            const int i = e >> 2, j = e & 3;
            const double *row = P + 10 * i;
            PB[e] = (j == 3) ? row[3] * Gam[15]
                             : row[j] * Gam[j * 4 + j] + row[4 + j] * Gam[(4 + j) * 4 + j] +
                                   row[7 + j] * Gam[(7 + j) * 4 + j];
        }
        __syncwarp();
        // (2) Bm = Phi' PB (10x4), S = R + Gam' PB (4x4), b = rt + Gam' pv ;
        //     reduced gradient and adjoint for the KKT error
        for (int e = lane; e < 40; e += 32) {
            const int i = e >> 2, j = e & 3;
            Bm[e] = phiT_dot(Phi, i, PB + j, 4);
        }
        if (lane < 16) {
            const int i = lane >> 2, j = lane & 3;
            double a = gamT_dot(Gam, i, PB + j, 4);
            if (i == j)
                a += s[L.rdiag + 4 * k + i] + delta;
            S[lane] = a;
        } else if (lane < 20) {
            const int i = lane - 16;
            bb[i] = s[L.rt + 4 * k + i] + gamT_dot(Gam, i, pv, 1);
        } else if (lane < 24) {
            const int i = lane - 20;
            const double gu = s[L.r + 4 * k + i] + gamT_dot(Gam, i, lam, 1);
            e_dual = fmax(e_dual, fabs(gu - s[L.zl + 4 * k + i] + s[L.zu + 4 * k + i]));
        }
        __syncwarp();
        // (3) S = L D L' in registers (all lanes), inertia check, solve for gains
        const double d0 = S[0];
        const double i0 = 1.0 / d0;
        const double l10 = S[4] * i0, l20 = S[8] * i0, l30 = S[12] * i0;
        const double d1 = S[5] - l10 * l10 * d0;
        const double i1 = 1.0 / d1;
        const double l21 = (S[9] - l20 * l10 * d0) * i1, l31 = (S[13] - l30 * l10 * d0) * i1;
        const double d2 = S[10] - l20 * l20 * d0 - l21 * l21 * d1;
        const double i2 = 1.0 / d2;
        const double l32 = (S[14] - l30 * l20 * d0 - l31 * l21 * d1) * i2;
        const double d3 = S[15] - l30 * l30 * d0 - l31 * l31 * d1 - l32 * l32 * d2;
        const double i3 = 1.0 / d3;
        if (!(d0 > 0.0) || !(d1 > 0.0) || !(d2 > 0.0) || !(d3 > 0.0)) {
            ok = false;
            break; // uniform: every lane computed the same pivots
        }
        if (lane <= 10) { // columns 0..9 of -Bm' and column 10 = -b
            double b0, b1, b2, b3;
            if (lane < 10) {
                b0 = -Bm[lane * 4 + 0], b1 = -Bm[lane * 4 + 1], b2 = -Bm[lane * 4 + 2],
                b3 = -Bm[lane * 4 + 3];
            } else {
                b0 = -bb[0], b1 = -bb[1], b2 = -bb[2], b3 = -bb[3];
            }
            const double w0 = b0;
            const double w1 = b1 - l10 * w0;
            const double w2 = b2 - l20 * w0 - l21 * w1;
            const double w3 = b3 - l30 * w0 - l31 * w1 - l32 * w2;
            const double y3 = w3 * i3;
            const double y2 = w2 * i2 - l32 * y3;
            const double y1 = w1 * i1 - l21 * y2 - l31 * y3;
            const double y0 = w0 * i0 - l10 * y1 - l20 * y2 - l30 * y3;
            if (lane < 10) {
                double *Kg = s + L.Kg + 40 * k;
                Kg[0 * 10 + lane] = y0;
                Kg[1 * 10 + lane] = y1;
                Kg[2 * 10 + lane] = y2;
                Kg[3 * 10 + lane] = y3;
            } else {
                double *kf = s + L.kf + 4 * k;
                kf[0] = y0, kf[1] = y1, kf[2] = y2, kf[3] = y3;
            }
        }
        __syncwarp();
        if (k == 0)
            break;
        // (4) P_k = Q_k + delta I + Phi' PA + Bm Kg (upper triangle, mirrored),
        //     p_k = q_k + Phi' pv + Bm kf,  lam_k = q_k + Phi' lam_{k+1}
        const double *Kg = s + L.Kg + 40 * k, *kf = s + L.kf + 4 * k;
        const double *Hc = s + L.Hc + 21 * (k - 1);
        double pn = 0.0, ln = 0.0;
        if (lane < 10) {
            const double qk = s[L.q + 10 * k + lane];
            pn = qk + phiT_dot(Phi, lane, pv, 1) + Bm[lane * 4 + 0] * kf[0] +
                 Bm[lane * 4 + 1] * kf[1] + Bm[lane * 4 + 2] * kf[2] + Bm[lane * 4 + 3] * kf[3];
            ln = qk + phiT_dot(Phi, lane, lam, 1);
        }
        double pe[2];
        int pi_[2], pj_[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int e = lane + 32 * m; // 55 upper-triangular entries
            pi_[m] = -1;
            if (e < 55) {
                int i = 0, rem = e;
                while (rem >= 10 - i) {
                    rem -= 10 - i;
                    ++i;
                }
                const int j = i + rem;
                double a = stage_hess(Hc, qp, i, j) + (i == j ? delta : 0.0);
                a += phiT_dot(Phi, i, PA + j, 10);
                a += Bm[i * 4 + 0] * Kg[0 * 10 + j] + Bm[i * 4 + 1] * Kg[1 * 10 + j] +
                     Bm[i * 4 + 2] * Kg[2 * 10 + j] + Bm[i * 4 + 3] * Kg[3 * 10 + j];
                pe[m] = a;
                pi_[m] = i;
                pj_[m] = j;
            }
        }
        __syncwarp(); // all reads of P, pv, lam done
#pragma unroll
        for (int m = 0; m < 2; ++m)
            if (pi_[m] >= 0) {
                P[pi_[m] * 10 + pj_[m]] = pe[m];
                P[pj_[m] * 10 + pi_[m]] = pe[m];
            }
        if (lane < 10) {
            pv[lane] = pn;
            lam[lane] = ln;
        }
        __syncwarp();
    }
    ok = __all_sync(AMPC_FULL_MASK, ok);
    *e_dual_out = warp_max(e_dual);
    return ok;
}

// forward sweep: du_k = Kg_k dx_k + kf_k, dx_{k+1} = Phi dx_k + Gam du_k, dx_0 = 0
__device__ void riccati_forward(const WarpCtx &w) {
    const SolveConsts &c = *w.c;
    const int N = c.N, lane = w.lane;
    double *s = w.s;
    const WarpLayout &L = w.L;
    if (lane < 10)
        s[L.dx + lane] = 0.0;
    __syncwarp();
    for (int k = 0; k < N; ++k) {
        if (lane < 4) {
            double a = s[L.kf + 4 * k + lane];
            if (k > 0) {
                const double *Kg = s + L.Kg + 40 * k + 10 * lane;
                const double *dx = s + L.dx + 10 * k;
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    a += Kg[j] * dx[j];
            }
            s[L.du + 4 * k + lane] = a;
        }
        __syncwarp();
        if (lane < 10)
            s[L.dx + 10 * (k + 1) + lane] = phi_row_dot(c.Phi, lane, s + L.dx + 10 * k) +
                                            gam_row_dot(c.Gam, lane, s + L.du + 4 * k);
        __syncwarp();
    }
}

struct SolveOut {
    double cost, kkt_dual, kkt_compl, mu;
    int32_t iters, status, n_reg, n_backtrack;
};

__device__ void solve_instance(const WarpCtx &w, double *w_inout, SolveOut *out) {
    const SolveConsts &c = *w.c;
    const int N = c.N, lane = w.lane, nu = 4 * N;
    double *s = w.s;
    const WarpLayout &L = w.L;
    const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99;
    const double eta = 1e-4, kappa_sigma = 1e10;
    const double mu_min = c.tol / 10.0;
    double mu = c.mu_init, delta_last = 0.0;
    int n_reg = 0, n_bt = 0;

    // controls from the warm start pushed into the box interior; cos/sin of the ref yaw
    for (int e = lane; e < nu; e += 32) {
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
    for (int k = lane; k < N; k += 32) {
        const double yaw = w.prefix[10 + 10 * k + 3];
        s[L.cs + 2 * k] = cos(yaw);
        s[L.cs + 2 * k + 1] = sin(-yaw);
    }
    if (lane < 10)
        s[L.x + lane] = w.prefix[lane];
    __syncwarp();
    for (int k = 0; k < N; ++k) { // roll-out
        if (lane < 10)
            s[L.x + 10 * (k + 1) + lane] = c.gam[lane] + phi_row_dot(c.Phi, lane, s + L.x + 10 * k) +
                                           gam_row_dot(c.Gam, lane, s + L.u + 4 * k);
        __syncwarp();
    }

    int status = 1, iter = 0;
    double e_dual = 0.0, e_compl = 0.0;
    for (iter = 0;; ++iter) {
        double eps = fmax(c.eps_min, c.eps_scale * mu);
        double f = eval_cost<true, false>(w, 0.0, eps);
        // barrier quantities at the current mu
        double c_mu = 0.0, ec = 0.0;
        for (int e = lane; e < nu; e += 32) {
            const int i = e & 3;
            const double sl = s[L.u + e] - c.lb[i], su = c.ub[i] - s[L.u + e];
            const double zl = s[L.zl + e], zu = s[L.zu + e];
            ec = fmax(ec, fmax(sl * zl, su * zu));
            c_mu = fmax(c_mu, fmax(fabs(sl * zl - mu), fabs(su * zu - mu)));
            s[L.rdiag + e] = 2.0 * c.wgt[20 + i] + zl / sl + zu / su;
            s[L.rt + e] = s[L.r + e] - mu / sl + mu / su;
        }
        e_compl = warp_max(ec);
        c_mu = warp_max(c_mu);
        __syncwarp();
        double delta = 0.0;
        int ntry = 0;
        bool bad = false;
        for (;;) { // Newton system with inertia correction (IPOPT Algorithm IC schedule)
            // (the adjoint / dual-infeasibility part of the sweep does not depend on delta)
            if (riccati_backward(w, delta, &e_dual))
                break;
            if (delta == 0.0)
                delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else
                delta *= (delta_last == 0.0) ? 100.0 : 8.0;
            if (++ntry > 60 || delta > 1e40) {
                bad = true;
                break;
            }
        }
        if (bad || !(f == f) || !(e_dual == e_dual)) {
            status = 3;
            break;
        }
        if (fmax(e_dual, e_compl) <= c.tol) {
            status = 0;
            break;
        }
        if (iter >= c.max_iter) {
            status = 1;
            break;
        }
        // monotone barrier update (IPOPT eq. (7)); a change of mu changes the
        // smoothing and the barrier gradient, so the sweep is redone
        bool mu_changed = false;
        while (mu > mu_min && fmax(e_dual, c_mu) <= kappa_eps * mu) {
            mu = fmax(mu_min, fmin(kappa_mu * mu, pow(mu, theta_mu)));
            mu_changed = true;
            double cm = 0.0;
            for (int e = lane; e < nu; e += 32) {
                const int i = e & 3;
                const double sl = s[L.u + e] - c.lb[i], su = c.ub[i] - s[L.u + e];
                cm = fmax(cm, fmax(fabs(sl * s[L.zl + e] - mu), fabs(su * s[L.zu + e] - mu)));
            }
            c_mu = warp_max(cm);
        }
        if (mu_changed) {
            eps = fmax(c.eps_min, c.eps_scale * mu);
            f = eval_cost<true, false>(w, 0.0, eps);
            for (int e = lane; e < nu; e += 32) {
                const int i = e & 3;
                const double sl = s[L.u + e] - c.lb[i], su = c.ub[i] - s[L.u + e];
                s[L.rt + e] = s[L.r + e] - mu / sl + mu / su;
            }
            __syncwarp();
            delta = 0.0;
            ntry = 0;
            for (;;) {
                double ed;
                if (riccati_backward(w, delta, &ed))
                    break;
                if (delta == 0.0)
                    delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
                else
                    delta *= (delta_last == 0.0) ? 100.0 : 8.0;
                if (++ntry > 60 || delta > 1e40) {
                    bad = true;
                    break;
                }
            }
            if (bad) {
                status = 3;
                break;
            }
        }
        if (delta > 0.0) {
            delta_last = delta;
            ++n_reg;
        }
        riccati_forward(w);
        const double tau_f = fmax(tau_min, 1.0 - mu);
        // fraction to the boundary (IPOPT eq. (15)), barrier value, slope
        double a_pri = 1.0, a_du = 1.0, bar0 = 0.0, gdw = 0.0;
        for (int e = lane; e < nu; e += 32) {
            const int i = e & 3;
            const double uu = s[L.u + e], du = s[L.du + e];
            const double sl = uu - c.lb[i], su = c.ub[i] - uu;
            const double zl = s[L.zl + e], zu = s[L.zu + e];
            const double dzl = mu / sl - zl - zl / sl * du;
            const double dzu = mu / su - zu + zu / su * du;
            if (du < 0.0)
                a_pri = fmin(a_pri, -tau_f * sl / du);
            if (du > 0.0)
                a_pri = fmin(a_pri, tau_f * su / du);
            if (dzl < 0.0)
                a_du = fmin(a_du, -tau_f * zl / dzl);
            if (dzu < 0.0)
                a_du = fmin(a_du, -tau_f * zu / dzu);
            bar0 += log(sl) + log(su);
            gdw += s[L.rt + e] * du;
        }
        for (int e = 10 + lane; e < 10 * (N + 1); e += 32)
            gdw += s[L.q + e] * s[L.dx + e];
        a_pri = warp_min(a_pri);
        a_du = warp_min(a_du);
        bar0 = warp_sum(bar0);
        gdw = warp_sum(gdw);
        const double phi0 = f - mu * bar0;
        // Armijo backtracking on the barrier objective
        double alpha = a_pri;
        bool accepted = false;
        for (int ls = 0; ls < 40; ++ls) {
            double bar = 0.0;
            for (int e = lane; e < nu; e += 32) {
                const int i = e & 3;
                const double ut = s[L.u + e] + alpha * s[L.du + e];
                bar += log(ut - c.lb[i]) + log(c.ub[i] - ut);
            }
            bar = warp_sum(bar);
            const double phit = eval_cost<false, true>(w, alpha, eps) - mu * bar;
            if (phit <= phi0 + eta * alpha * gdw + 10.0 * 2.220446049250313e-16 * fabs(phi0)) {
                accepted = true;
                break;
            }
            alpha *= 0.5;
            ++n_bt;
        }
        if (!accepted) {
            status = 2;
            break;
        }
        for (int e = lane; e < nu; e += 32) {
            const int i = e & 3;
            const double u0 = s[L.u + e], du = s[L.du + e];
            const double sl0 = u0 - c.lb[i], su0 = c.ub[i] - u0;
            double zl = s[L.zl + e], zu = s[L.zu + e];
            const double dzl = mu / sl0 - zl - zl / sl0 * du;
            const double dzu = mu / su0 - zu + zu / su0 * du;
            const double un = u0 + alpha * du;
            const double sl = un - c.lb[i], su = c.ub[i] - un;
            zl += a_du * dzl;
            zu += a_du * dzu;
            zl = fmax(fmin(zl, kappa_sigma * mu / sl), mu / (kappa_sigma * sl)); // IPOPT eq. (16)
            zu = fmax(fmin(zu, kappa_sigma * mu / su), mu / (kappa_sigma * su));
            s[L.u + e] = un;
            s[L.zl + e] = zl;
            s[L.zu + e] = zu;
        }
        for (int e = 10 + lane; e < 10 * (N + 1); e += 32)
            s[L.x + e] += alpha * s[L.dx + e];
        __syncwarp();
    }
    // results: w = [X_0,U_0,...,X_N]; objective without smoothing
    const double cost = eval_cost<false, false>(w, 0.0, 0.0);
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
__global__ void __launch_bounds__(WARPS * 32)
ipm_solve_kernel(const __grid_constant__ SolveConsts consts, int B, const double *__restrict__ prefix,
                 double *__restrict__ w_inout, SolveOut *__restrict__ info) {
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
    if (b >= B)
        return;
    const WarpLayout L(sc->N);
    WarpCtx ctx(sc, wbase + (size_t)warp * L.total, prefix + (size_t)b * sc->n_prefix, lane);
    solve_instance(ctx, w_inout + (size_t)b * (10 + 14 * sc->N), info + b);
}

} // namespace ampc
