/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement, in plain C (double), of
 * the reference's quadrotor collision-avoidance NLP and of the interior-point
 * iteration this repo's CUDA solver implements.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (avoid-mpc_b200/) never does.
 *
 * Parity status: PARITY UNPINNED against the reference's third-party solver.
 * The NLP *functions* (f, grad f, g, Hessian) restate
 * roswrapper/ros/src/avoid_mpc/tools/mpc_obstacle_casadi.py line by line and
 * are cross-checked here against torch.autograd (float64) re-derivations of
 * that script (tests/test_oracle_nlp.py).  The *solver* in the reference is
 * CasADi 3.6.4 -> IPOPT -> MUMPS (README.md:41-43), none of which exist in this
 * image, and the reference ships no golden vector for it; so the solve is
 * checked at convergence against an independent interior-point solver of the
 * same class (scipy.optimize trust-constr on the reference's full-space
 * formulation: w = [X, U], g(w) = 0, bounds on U, exact Hessian; 52 golden
 * optima in tests/golden/solve_golden.npz, minted by
 * tests/golden/make_solve_golden.py), not against IPOPT.
 *
 * Reference (paths under roswrapper/ros/src/avoid_mpc/):
 *   tools/mpc_obstacle_casadi.py:36-48    dimensions (s_dim 10, u_dim 4, weights 25)
 *   tools/mpc_obstacle_casadi.py:76-94    parameter vector P layout (gain, tau at the tail)
 *   tools/mpc_obstacle_casadi.py:106-122  ODE (drag off)
 *   tools/mpc_obstacle_casadi.py:134-149  P slices: x0, ref, obstacles, target, weights, radius
 *   tools/mpc_obstacle_casadi.py:156-220  decision vector, cost, constraints
 *   tools/mpc_obstacle_casadi.py:250-251  softplus = log(1 + exp(x)), un-stabilised
 *   tools/mpc_obstacle_casadi.py:338-357  RK4 with M = 4 sub-steps
 *   src/HighLvlMpc.cpp:25-49,70-92        bounds (X free, U boxed), zero cold start
 *   src/HighLvlMpc.cpp:93-137             p tail packing [gains, tau, weights, radius]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NX 10
#define NU 4
#define NXU 14
#define GZ 9.81 /* mpc_obstacle_casadi.py:39 */

/* ---- layout helpers (mpc_obstacle_casadi.py:76-94,134-149) ---- */
int nlp_oracle_nw(int N) { return NX + NXU * N; }
int nlp_oracle_ng(int N) { return NX * (N + 1); }
int nlp_oracle_np(int N, int K) { return NX + NX * N + 3 * K * N + NX + 2 * NU + 25 + 1; }

typedef struct {
    const double *x0, *ref, *obst, *target, *gain, *tau, *wgt;
    double radius;
} pview;

static pview p_view(int N, int K, const double *p) {
    pview v;
    const int np = nlp_oracle_np(N, K);
    v.x0 = p;
    v.ref = p + NX;
    v.obst = p + NX + NX * N;
    v.target = v.obst + 3 * K * N;
    v.gain = p + np - 34; /* P[-34:-30] */
    v.tau = p + np - 30;  /* P[-30:-26] */
    v.wgt = p + np - 26;  /* P[-26:-1]  */
    v.radius = p[np - 1];
    return v;
}

/* ---- dynamics (mpc_obstacle_casadi.py:106-122) ---- */
static void ode(const double x[NX], const double u[NU], const double tau[4], double xd[NX]) {
    xd[0] = x[4];
    xd[1] = x[5];
    xd[2] = x[6];
    xd[3] = u[3];
    xd[4] = x[7];
    xd[5] = x[8];
    xd[6] = x[9];
    xd[7] = (u[0] - x[7]) * tau[0];
    xd[8] = (u[1] - x[8]) * tau[1];
    xd[9] = (u[2] - GZ - x[9]) * tau[2];
}

/* sys_dynamics (mpc_obstacle_casadi.py:338-357): F(x,u) = 4 RK4 sub-steps of dt/4. */
void nlp_oracle_F(const double x[NX], const double u[NU], const double tau[4], double dt,
                  double xn[NX]) {
    const int M = 4;
    const double DT = dt / M;
    double X[NX], k1[NX], k2[NX], k3[NX], k4[NX], t[NX];
    memcpy(X, x, sizeof X);
    for (int m = 0; m < M; ++m) {
        ode(X, u, tau, k1);
        for (int i = 0; i < NX; ++i) {
            k1[i] *= DT;
            t[i] = X[i] + 0.5 * k1[i];
        }
        ode(t, u, tau, k2);
        for (int i = 0; i < NX; ++i) {
            k2[i] *= DT;
            t[i] = X[i] + 0.5 * k2[i];
        }
        ode(t, u, tau, k3);
        for (int i = 0; i < NX; ++i) {
            k3[i] *= DT;
            t[i] = X[i] + k3[i];
        }
        ode(t, u, tau, k4);
        for (int i = 0; i < NX; ++i) {
            k4[i] *= DT;
            X[i] = X[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6;
        }
    }
    memcpy(xn, X, sizeof X);
}

/* F is affine in (x,u): F(x,u) = Phi x + Gam u + gam.  Extracted column by
 * column from F itself, so it is the RK4 polynomial, not expm. */
void nlp_oracle_dyn_matrices(const double tau[4], double dt, double Phi[NX * NX],
                             double Gam[NX * NU], double gam[NX]) {
    double z[NX] = {0}, zu[NU] = {0}, col[NX];
    nlp_oracle_F(z, zu, tau, dt, gam);
    for (int j = 0; j < NX; ++j) {
        double e[NX] = {0};
        e[j] = 1.0;
        nlp_oracle_F(e, zu, tau, dt, col);
        for (int i = 0; i < NX; ++i)
            Phi[i * NX + j] = col[i] - gam[i];
    }
    for (int j = 0; j < NU; ++j) {
        double e[NU] = {0};
        e[j] = 1.0;
        nlp_oracle_F(z, e, tau, dt, col);
        for (int i = 0; i < NX; ++i)
            Gam[i * NU + j] = col[i] - gam[i];
    }
}

/* ---- constraints g (mpc_obstacle_casadi.py:160,219) ---- */
void nlp_oracle_g(int N, int K, const double *w, const double *p, double dt, double *g) {
    const pview v = p_view(N, K, p);
    for (int i = 0; i < NX; ++i)
        g[i] = w[i] - v.x0[i];
    for (int k = 0; k < N; ++k) {
        double xn[NX];
        nlp_oracle_F(w + NXU * k, w + NXU * k + NX, v.tau, dt, xn);
        for (int i = 0; i < NX; ++i)
            g[NX * (k + 1) + i] = xn[i] - w[NXU * (k + 1) + i];
    }
}

static inline double sgn(double s) { return (s > 0) - (s < 0); }

/* One stage cost on X_{k+1} (mpc_obstacle_casadi.py:165-208): value, and
 * optionally gradient gx[10] and Hessian Hx[10x10] (both ACCUMULATED into). */
static double stage_cost(int N, int K, int k, const pview *v, const double *x, double *gx,
                         double *Hx, double eps) {
    double c = 0.0;
    if (k >= N - 1) { /* terminal: (X_N - target)' Q_goal (.) */
        for (int i = 0; i < NX; ++i) {
            const double d = x[i] - v->target[i];
            c += v->wgt[i] * d * d;
            if (gx)
                gx[i] += 2 * v->wgt[i] * d;
            if (Hx)
                Hx[i * NX + i] += 2 * v->wgt[i];
        }
        return c;
    }
    const double *ref = v->ref + NX * k;
    const double *qp = v->wgt + NX;
    const double cy = cos(ref[3]);
    const double sy = sin(-ref[3]); /* script: sin_yaw = sin(-yaw) */
    /* rot[0,0]=cy rot[0,1]=-sy rot[1,0]=sy rot[1,1]=cy (same for rows 4,5) */
    double rot[NX][NX];
    memset(rot, 0, sizeof rot);
    for (int i = 0; i < NX; ++i)
        rot[i][i] = 1.0;
    rot[0][0] = cy, rot[0][1] = -sy, rot[1][0] = sy, rot[1][1] = cy;
    rot[4][4] = cy, rot[4][5] = -sy, rot[5][4] = sy, rot[5][5] = cy;
    double dl[NX], rd[NX];
    for (int i = 0; i < NX; ++i)
        dl[i] = x[i] - ref[i];
    for (int i = 0; i < NX; ++i) {
        rd[i] = 0;
        for (int j = 0; j < NX; ++j)
            rd[i] += rot[i][j] * dl[j];
    }
    for (int i = 0; i < NX; ++i)
        c += qp[i] * rd[i] * rd[i];
    if (gx)
        for (int j = 0; j < NX; ++j) {
            double s = 0;
            for (int i = 0; i < NX; ++i)
                s += rot[i][j] * qp[i] * rd[i];
            gx[j] += 2 * s;
        }
    if (Hx)
        for (int a = 0; a < NX; ++a)
            for (int b = 0; b < NX; ++b) {
                double s = 0;
                for (int i = 0; i < NX; ++i)
                    s += rot[i][a] * qp[i] * rot[i][b];
                Hx[a * NX + b] += 2 * s;
            }
    /* collision terms (mpc_obstacle_casadi.py:186-204) */
    const double lam = v->wgt[24];
    const double *vel = x + 4;
    for (int j = 0; j < K; ++j) {
        const double *o = v->obst + 3 * (K * k + j);
        const double d[3] = {o[0] - x[0], o[1] - x[1], o[2] - x[2]};
        const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double n[3] = {d[0] / r, d[1] / r, d[2] / r};
        const double s = vel[0] * n[0] + vel[1] * n[1] + vel[2] * n[2];
        const double e = exp((r - v->radius) * -32.0);
        const double sp = log(1 + e);
        /* |s| (CasADi norm_2 of a scalar, script line 199).  eps = 0 is the
         * reference's function: psi = |s|, psi' = sign(s) (sign(0) = 0), psi'' = 0.
         * eps > 0 is used ONLY inside the solver: psi = sqrt(s^2+eps^2) - eps. */
        double psi, dpsi, ddpsi;
        if (eps > 0) {
            const double hyp = sqrt(s * s + eps * eps);
            psi = hyp - eps;
            dpsi = s / hyp;
            ddpsi = eps * eps / (hyp * hyp * hyp);
        } else {
            psi = fabs(s);
            dpsi = sgn(s);
            ddpsi = 0.0;
        }
        c += lam * sp * psi;
        if (!gx && !Hx)
            continue;
        const double sig = e / (1 + e);
        const double wv[3] = {(vel[0] - s * n[0]) / r, (vel[1] - s * n[1]) / r,
                              (vel[2] - s * n[2]) / r};
        if (gx)
            for (int a = 0; a < 3; ++a) {
                gx[a] += lam * (32 * sig * psi * n[a] - sp * dpsi * wv[a]);
                gx[4 + a] += lam * sp * dpsi * n[a];
            }
        if (Hx)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    const double Pi = (a == b ? 1.0 : 0.0) - n[a] * n[b];
                    const double hpp =
                        lam * (1024 * sig * (1 - sig) * psi * n[a] * n[b] -
                               32 * sig * dpsi * (n[a] * wv[b] + wv[a] * n[b]) -
                               32 * sig * psi * Pi / r + sp * ddpsi * wv[a] * wv[b] -
                               sp * dpsi * ((n[a] * wv[b] + s * Pi / r) / r + wv[a] * n[b] / r));
                    const double hpv = lam * (32 * sig * dpsi * n[a] * n[b] - sp * dpsi * Pi / r -
                                              sp * ddpsi * wv[a] * n[b]);
                    Hx[a * NX + b] += hpp;
                    Hx[a * NX + 4 + b] += hpv;   /* d2c / dp_a dv_b */
                    Hx[(4 + b) * NX + a] += hpv; /* symmetric entry */
                    Hx[(4 + a) * NX + 4 + b] += lam * sp * ddpsi * n[a] * n[b];
                }
    }
    return c;
}

/* objective (mpc_obstacle_casadi.py:162-214) */
double nlp_oracle_f(int N, int K, const double *w, const double *p) {
    const pview v = p_view(N, K, p);
    const double *qu = v.wgt + 2 * NX;
    const double uref[NU] = {0, 0, GZ, 0};
    double f = 0.0;
    for (int k = 0; k < N; ++k) {
        const double *u = w + NXU * k + NX;
        double cu = 0;
        for (int i = 0; i < NU; ++i)
            cu += qu[i] * (u[i] - uref[i]) * (u[i] - uref[i]);
        f += stage_cost(N, K, k, &v, w + NXU * (k + 1), NULL, NULL, 0.0) + cu;
    }
    return f;
}

void nlp_oracle_grad_f(int N, int K, const double *w, const double *p, double *grad) {
    const pview v = p_view(N, K, p);
    const double *qu = v.wgt + 2 * NX;
    const double uref[NU] = {0, 0, GZ, 0};
    memset(grad, 0, sizeof(double) * (size_t)nlp_oracle_nw(N));
    for (int k = 0; k < N; ++k) {
        const double *u = w + NXU * k + NX;
        for (int i = 0; i < NU; ++i)
            grad[NXU * k + NX + i] = 2 * qu[i] * (u[i] - uref[i]);
        stage_cost(N, K, k, &v, w + NXU * (k + 1), grad + NXU * (k + 1), NULL, 0.0);
    }
}

/* Hessian of f: block diagonal.  Hx: N blocks of 10x10 (row-major) for
 * X_1..X_N; Hu: 4 diagonal entries (identical at every stage). */
void nlp_oracle_hess_f(int N, int K, const double *w, const double *p, double *Hx, double *Hu) {
    const pview v = p_view(N, K, p);
    memset(Hx, 0, sizeof(double) * (size_t)(N * NX * NX));
    for (int i = 0; i < NU; ++i)
        Hu[i] = 2 * v.wgt[2 * NX + i];
    for (int k = 0; k < N; ++k) {
        double g[NX] = {0};
        stage_cost(N, K, k, &v, w + NXU * (k + 1), g, Hx + (size_t)k * NX * NX, 0.0);
    }
}

/* ======================================================================== */
/* Interior-point solve.  This is the algorithm the CUDA solver implements   */
/* (avoid-mpc_b200/csrc/ipm_solve.cu); see DESIGN.md "Solver".  The          */
/* reference hands the NLP to IPOPT (HighLvlMpc.cpp:50-52,116-122); IPOPT's  */
/* filter line search / MUMPS pivoting are not reproducible, so parity is    */
/* defined at convergence of the same NLP from the same starting controls.   */
/*                                                                           */
/* Because F is affine, every iterate is kept on the dynamics manifold: the  */
/* controls U are the free variables, X = roll-out(x0, U).  The X part of    */
/* the warm start is therefore replaced by the roll-out of its U part.  The  */
/* Newton system of the multiple-shooting NLP (variables X and U, equality   */
/* multipliers lam) is solved exactly, stage by stage, by a Riccati sweep;   */
/* bounds on U are handled by a primal-dual log barrier with IPOPT's         */
/* monotone mu rule, fraction-to-the-boundary rule, inertia-correction       */
/* schedule and multiplier safeguard (Waechter & Biegler 2006, eqs. 7,15,16  */
/* and Algorithm IC) and an Armijo backtracking line search on the barrier   */
/* objective (no filter is needed: there is no constraint violation).        */
/* The reference's |v.n| term is non-smooth and its minimisers frequently    */
/* sit exactly on the kink (a stage at closest approach has v.n = 0), where  */
/* Newton iterations zig-zag.  Inside the solver |s| is therefore replaced   */
/* by sqrt(s^2 + eps^2) - eps with eps = mu (driven to tol/10 together with  */
/* the barrier), the smoothing an interior-point method applies to the       */
/* epigraph form t >= |s|.  nlp_oracle_f/grad_f/hess_f above stay the        */
/* reference's exact functions; the reported cost is the un-smoothed f.      */
/* ======================================================================== */

typedef struct {
    double tol;        /* KKT tolerance (role of ipopt.tol) */
    int32_t max_iter;
    double mu_init;    /* 0.1 (IPOPT default) */
    double bound_push; /* 1e-2 (IPOPT default bound_push) */
    double bound_frac; /* 1e-2 (IPOPT default bound_frac) */
    double eps_min;    /* floor of the |s| smoothing, m/s */
    double eps_scale;  /* eps = max(eps_min, eps_scale * mu) */
    double tau_min;    /* fraction-to-the-boundary floor (IPOPT tau_min, 0.99) */
    double kappa_eps;  /* barrier sub-problem tolerance factor (IPOPT barrier_tol_factor is 10; 100 saves ~15% iterations) */
    double kappa_mu;   /* linear mu decrease (IPOPT mu_linear_decrease_factor, 0.2) */
    double theta_mu;   /* superlinear mu decrease (IPOPT mu_superlinear_decrease_power, 1.5) */
    int32_t proj_step; /* 1: per-component (projected) step limiting in the line search */
} nlp_oracle_opts;

typedef struct {
    double cost;    /* f at the returned point */
    int32_t iters;
    int32_t status; /* 0 converged, 1 max_iter, 2 line-search stall, 3 numerical failure */
    double kkt_dual, kkt_primal, kkt_compl;
    double mu;
    double reg_last; /* last inertia-correcting regularisation used */
    int32_t n_reg;   /* iterations that needed regularisation */
    int32_t n_backtrack;
} nlp_oracle_info;

void nlp_oracle_default_opts(nlp_oracle_opts *o) {
    o->tol = 1e-8;
    o->max_iter = 100;
    o->mu_init = 0.1;
    o->bound_push = 1e-2;
    o->bound_frac = 1e-2;
    o->eps_min = 1e-5;
    o->eps_scale = 1.0;
    o->tau_min = 0.99;
    o->kappa_eps = 100.0;
    o->kappa_mu = 0.2;
    o->theta_mu = 1.5;
    o->proj_step = 1;
}

typedef struct {
    int N, K;
    double Phi[NX * NX], Gam[NX * NU], gam[NX];
    double lb[NU], ub[NU];
    pview v;
    double *x, *u, *zl, *zu;   /* iterate: x (N+1)*10, u/zl/zu N*4 */
    double *q, *Q, *r;         /* grad x, Hess x blocks (index k = 0..N; k = 0 unused), grad u */
    double *dx, *du, *Kg, *kf; /* Newton step and Riccati gains */
} ipm;

static double eval_all(ipm *s, const double *x, const double *u, int need_derivs, double eps) {
    const int N = s->N, K = s->K;
    const double *qu = s->v.wgt + 2 * NX;
    const double uref[NU] = {0, 0, GZ, 0};
    double f = 0;
    if (need_derivs) {
        memset(s->q, 0, sizeof(double) * (size_t)((N + 1) * NX));
        memset(s->Q, 0, sizeof(double) * (size_t)((N + 1) * NX * NX));
    }
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < NU; ++i) {
            const double du = u[NU * k + i] - uref[i];
            f += qu[i] * du * du;
            if (need_derivs)
                s->r[NU * k + i] = 2 * qu[i] * du;
        }
        f += stage_cost(N, K, k, &s->v, x + NX * (k + 1), need_derivs ? s->q + NX * (k + 1) : NULL,
                        need_derivs ? s->Q + (size_t)(k + 1) * NX * NX : NULL, eps);
    }
    return f;
}

static void rollout(const ipm *s, const double *u, double *x) {
    for (int k = 0; k < s->N; ++k)
        for (int i = 0; i < NX; ++i) {
            double a = s->gam[i];
            for (int j = 0; j < NX; ++j)
                a += s->Phi[i * NX + j] * x[NX * k + j];
            for (int j = 0; j < NU; ++j)
                a += s->Gam[i * NU + j] * u[NU * k + j];
            x[NX * (k + 1) + i] = a;
        }
}

/* Riccati sweep for  min 1/2 dw'H dw + grad'dw  s.t. dx_{k+1} = Phi dx_k + Gam du_k,
 * dx_0 = 0;  Hessian blocks Q_k + delta I (k = 1..N) and diag(rdiag_k) + delta I.
 * Returns 0 on success, 1 if some S_k = R_k + Gam'P_{k+1}Gam is not positive
 * definite (== the reduced Hessian is not positive definite: wrong inertia). */
static int riccati(ipm *s, const double *rdiag, const double *rt, double delta) {
    const int N = s->N;
    double P[NX * NX], pv[NX];
    for (int i = 0; i < NX * NX; ++i)
        P[i] = s->Q[(size_t)N * NX * NX + i];
    for (int i = 0; i < NX; ++i) {
        P[i * NX + i] += delta;
        pv[i] = s->q[NX * N + i];
    }
    for (int k = N - 1; k >= 0; --k) {
        double PA[NX * NX], PB[NX * NU];
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) {
                double a = 0;
                for (int l = 0; l < NX; ++l)
                    a += P[i * NX + l] * s->Phi[l * NX + j];
                PA[i * NX + j] = a;
            }
            for (int j = 0; j < NU; ++j) {
                double a = 0;
                for (int l = 0; l < NX; ++l)
                    a += P[i * NX + l] * s->Gam[l * NU + j];
                PB[i * NU + j] = a;
            }
        }
        double S[NU * NU], Bm[NX * NU], b[NU];
        for (int i = 0; i < NU; ++i) {
            for (int j = 0; j < NU; ++j) {
                double a = (i == j) ? rdiag[NU * k + i] + delta : 0.0;
                for (int l = 0; l < NX; ++l)
                    a += s->Gam[l * NU + i] * PB[l * NU + j];
                S[i * NU + j] = a;
            }
            double a = rt[NU * k + i];
            for (int l = 0; l < NX; ++l)
                a += s->Gam[l * NU + i] * pv[l];
            b[i] = a;
        }
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NU; ++j) { /* Bm = Phi' P Gam  (10x4) */
                double a = 0;
                for (int l = 0; l < NX; ++l)
                    a += s->Phi[l * NX + i] * PB[l * NU + j];
                Bm[i * NU + j] = a;
            }
        /* Cholesky S = L L' */
        double L[NU * NU] = {0};
        for (int i = 0; i < NU; ++i)
            for (int j = 0; j <= i; ++j) {
                double a = S[i * NU + j];
                for (int l = 0; l < j; ++l)
                    a -= L[i * NU + l] * L[j * NU + l];
                if (i == j) {
                    if (!(a > 0))
                        return 1;
                    L[i * NU + i] = sqrt(a);
                } else
                    L[i * NU + j] = a / L[j * NU + j];
            }
        /* solve S [Kg | kf] = -[Bm' | b] */
        double *Kg = s->Kg + (size_t)k * NU * NX, *kf = s->kf + NU * k;
        for (int c = 0; c <= NX; ++c) {
            double y[NU];
            for (int i = 0; i < NU; ++i) {
                double a = (c < NX) ? -Bm[c * NU + i] : -b[i];
                for (int l = 0; l < i; ++l)
                    a -= L[i * NU + l] * y[l];
                y[i] = a / L[i * NU + i];
            }
            for (int i = NU - 1; i >= 0; --i) {
                double a = y[i];
                for (int l = i + 1; l < NU; ++l)
                    a -= L[l * NU + i] * y[l];
                y[i] = a / L[i * NU + i];
            }
            for (int i = 0; i < NU; ++i) {
                if (c < NX)
                    Kg[i * NX + c] = y[i];
                else
                    kf[i] = y[i];
            }
        }
        if (k == 0)
            break; /* P_0, p_0 are not needed: dx_0 = 0 */
        /* P_k = Q_k + delta I + Phi' PA + Bm Kg ; p_k = q_k + Phi' p_{k+1} + Bm kf */
        double Pn[NX * NX], pn[NX];
        for (int i = 0; i < NX; ++i) {
            for (int j = 0; j < NX; ++j) {
                double a = s->Q[(size_t)k * NX * NX + i * NX + j] + (i == j ? delta : 0.0);
                for (int l = 0; l < NX; ++l)
                    a += s->Phi[l * NX + i] * PA[l * NX + j];
                for (int l = 0; l < NU; ++l)
                    a += Bm[i * NU + l] * Kg[l * NX + j];
                Pn[i * NX + j] = a;
            }
            double a = s->q[NX * k + i];
            for (int l = 0; l < NX; ++l)
                a += s->Phi[l * NX + i] * pv[l];
            for (int l = 0; l < NU; ++l)
                a += Bm[i * NU + l] * kf[l];
            pn[i] = a;
        }
        for (int i = 0; i < NX; ++i) { /* symmetrise */
            for (int j = 0; j < NX; ++j)
                P[i * NX + j] = 0.5 * (Pn[i * NX + j] + Pn[j * NX + i]);
            pv[i] = pn[i];
        }
    }
    /* forward sweep */
    for (int i = 0; i < NX; ++i)
        s->dx[i] = 0.0;
    for (int k = 0; k < N; ++k) {
        const double *Kg = s->Kg + (size_t)k * NU * NX;
        for (int i = 0; i < NU; ++i) {
            double a = s->kf[NU * k + i];
            for (int j = 0; j < NX; ++j)
                a += Kg[i * NX + j] * s->dx[NX * k + j];
            s->du[NU * k + i] = a;
        }
        for (int i = 0; i < NX; ++i) {
            double a = 0;
            for (int j = 0; j < NX; ++j)
                a += s->Phi[i * NX + j] * s->dx[NX * k + j];
            for (int j = 0; j < NU; ++j)
                a += s->Gam[i * NU + j] * s->du[NU * k + j];
            s->dx[NX * (k + 1) + i] = a;
        }
    }
    return 0;
}

/* Solve one instance.  p: full parameter vector (n_p); w: n_w, warm start in /
 * solution out (layout [X_0,U_0,...,U_{N-1},X_N]); lbu/ubu: control box;
 * lam_g_out (n_g, may be NULL): multipliers of g with L = f + lam'g. */
int nlp_oracle_solve(int N, int K, double dt, const double *p, double *w, const double lbu[NU],
                     const double ubu[NU], const nlp_oracle_opts *opt, nlp_oracle_info *info,
                     double *lam_g_out) {
    ipm s;
    memset(&s, 0, sizeof s);
    s.N = N;
    s.K = K;
    s.v = p_view(N, K, p);
    nlp_oracle_dyn_matrices(s.v.tau, dt, s.Phi, s.Gam, s.gam);
    memcpy(s.lb, lbu, sizeof s.lb);
    memcpy(s.ub, ubu, sizeof s.ub);
    const size_t nxs = (size_t)(N + 1) * NX, nus = (size_t)N * NU;
    double *buf = (double *)calloc(nxs * 5 + nus * 11 + (size_t)(N + 1) * NX * NX + nus * NX, sizeof(double));
    double *b = buf;
    s.x = b, b += nxs;
    s.q = b, b += nxs;
    s.dx = b, b += nxs;
    double *xt = b;
    b += nxs;
    s.u = b, b += nus;
    s.zl = b, b += nus;
    s.zu = b, b += nus;
    s.r = b, b += nus;
    s.du = b, b += nus;
    s.kf = b, b += nus;
    double *sig = b;
    b += nus;
    double *rt = b;
    b += nus;
    double *ut = b;
    b += nus;
    double *rdiag = b;
    b += nus;
    double *dut = b;
    b += nus;
    double *dxt = b;
    b += nxs;
    s.Q = b, b += (size_t)(N + 1) * NX * NX;
    s.Kg = b, b += nus * NX;

    const double kappa_eps = opt->kappa_eps, kappa_mu = opt->kappa_mu, theta_mu = opt->theta_mu, tau_min = opt->tau_min;
    const double eta = 1e-4, kappa_sigma = 1e10;
    const double mu_min = opt->tol / 10.0;
#define EPS_OF(m) fmax(opt->eps_min, opt->eps_scale * (m))
    double mu = opt->mu_init, delta_last = 0.0;
    memset(info, 0, sizeof *info);

    /* controls from the warm start, pushed into the interior of their box;
     * states from the roll-out */
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i) {
            const double lo = s.lb[i], hi = s.ub[i];
            double pl = opt->bound_push * fmax(1.0, fabs(lo)), pu = opt->bound_push * fmax(1.0, fabs(hi));
            pl = fmin(pl, opt->bound_frac * (hi - lo));
            pu = fmin(pu, opt->bound_frac * (hi - lo));
            double uu = w[NXU * k + NX + i];
            uu = fmax(uu, lo + pl);
            uu = fmin(uu, hi - pu);
            s.u[NU * k + i] = uu;
            s.zl[NU * k + i] = mu / (uu - lo);
            s.zu[NU * k + i] = mu / (hi - uu);
        }
    for (int i = 0; i < NX; ++i)
        s.x[i] = s.v.x0[i];
    rollout(&s, s.u, s.x);

    int status = 1, iter = 0;
    double f = 0, e_dual = 0, e_compl = 0;
    for (iter = 0;; ++iter) {
        f = eval_all(&s, s.x, s.u, 1, EPS_OF(mu));
        /* adjoint multipliers (x-stationarity exact) and reduced gradient */
        double lam[NX], e_du = 0;
        for (int i = 0; i < NX; ++i)
            lam[i] = s.q[NX * N + i];
        for (int k = N - 1; k >= 0; --k) {
            for (int i = 0; i < NU; ++i) {
                double a = s.r[NU * k + i];
                for (int l = 0; l < NX; ++l)
                    a += s.Gam[l * NU + i] * lam[l];
                e_du = fmax(e_du, fabs(a - s.zl[NU * k + i] + s.zu[NU * k + i]));
            }
            if (lam_g_out) /* L = f + lam'g with g_{k+1} = F(X_k,U_k) - X_{k+1}  =>  lam_g = +grad_x cost-to-go */
                for (int i = 0; i < NX; ++i)
                    lam_g_out[NX * (k + 1) + i] = lam[i];
            double ln[NX];
            for (int i = 0; i < NX; ++i) {
                double a = s.q[NX * k + i];
                for (int l = 0; l < NX; ++l)
                    a += s.Phi[l * NX + i] * lam[l];
                ln[i] = a;
            }
            memcpy(lam, ln, sizeof lam);
        }
        if (lam_g_out) /* g_0 = X_0 - x0: stationarity wrt X_0 */
            for (int i = 0; i < NX; ++i)
                lam_g_out[i] = -lam[i];
        e_dual = e_du;
        double c_mu = 0;
        e_compl = 0;
        for (size_t i = 0; i < nus; ++i) {
            const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
            e_compl = fmax(e_compl, fmax(sl * s.zl[i], su * s.zu[i]));
            c_mu = fmax(c_mu, fmax(fabs(sl * s.zl[i] - mu), fabs(su * s.zu[i] - mu)));
        }
        if (!(f == f) || !(e_dual == e_dual)) {
            status = 3;
            break;
        }
        if (fmax(e_dual, e_compl) <= opt->tol) {
            status = 0;
            break;
        }
        if (iter >= opt->max_iter) {
            status = 1;
            break;
        }
        /* monotone barrier update (Fiacco-McCormick; IPOPT eq. (7)) */
        const double mu_before = mu;
        while (mu > mu_min && fmax(e_dual, c_mu) <= kappa_eps * mu) {
            mu = fmax(mu_min, fmin(kappa_mu * mu, pow(mu, theta_mu)));
            c_mu = 0;
            for (size_t i = 0; i < nus; ++i) {
                const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
                c_mu = fmax(c_mu, fmax(fabs(sl * s.zl[i] - mu), fabs(su * s.zu[i] - mu)));
            }
        }
        if (mu != mu_before) /* the smoothing changed with mu: refresh f, q, Q */
            f = eval_all(&s, s.x, s.u, 1, EPS_OF(mu));
        const double tau_f = fmax(tau_min, 1.0 - mu);
        /* primal-dual barrier system in the controls */
        double phi0 = f;
        for (size_t i = 0; i < nus; ++i) {
            const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
            sig[i] = s.zl[i] / sl + s.zu[i] / su;
            rt[i] = s.r[i] - mu / sl + mu / su;
            rdiag[i] = 2 * s.v.wgt[2 * NX + (i % NU)] + sig[i];
            phi0 -= mu * (log(sl) + log(su));
        }
        /* Newton step with inertia correction (IPOPT Algorithm IC schedule) */
        double delta = 0.0;
        int ntry = 0, bad = 0;
        while (riccati(&s, rdiag, rt, delta)) {
            if (delta == 0.0)
                delta = (delta_last == 0.0) ? 1e-4 : fmax(1e-20, delta_last / 3.0);
            else
                delta *= (delta_last == 0.0) ? 100.0 : 8.0;
            if (++ntry > 60 || delta > 1e40) {
                bad = 1;
                break;
            }
        }
        if (bad) {
            status = 3;
            break;
        }
        if (delta > 0) {
            delta_last = delta;
            info->n_reg++;
        }
        info->reg_last = delta;
        /* fraction to the boundary (IPOPT eq. (15)) */
        double a_pri = 1.0, a_du = 1.0;
        for (size_t i = 0; i < nus; ++i) {
            const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
            const double dzl = mu / sl - s.zl[i] - s.zl[i] / sl * s.du[i];
            const double dzu = mu / su - s.zu[i] + s.zu[i] / su * s.du[i];
            if (s.du[i] < 0)
                a_pri = fmin(a_pri, -tau_f * sl / s.du[i]);
            if (s.du[i] > 0)
                a_pri = fmin(a_pri, tau_f * su / s.du[i]);
            if (dzl < 0)
                a_du = fmin(a_du, -tau_f * s.zl[i] / dzl);
            if (dzu < 0)
                a_du = fmin(a_du, -tau_f * s.zu[i] / dzu);
        }
        /* Armijo backtracking on the barrier objective phi_mu along (dx, du) */
        double gdw = 0;
        for (size_t i = NX; i < nxs; ++i)
            gdw += s.q[i] * s.dx[i];
        for (size_t i = 0; i < nus; ++i)
            gdw += rt[i] * s.du[i];
        /* Projected line search: the step is limited PER COMPONENT by the
         * fraction-to-the-boundary rule (a control that would cross its bound stops
         * at distance (1-tau)*slack from it) instead of scaling the whole step by
         * the most restrictive component; the states follow by linearity
         * (dx_t = rollout of du_t).  Every bound that wants to become active is
         * reached in one iteration.  The Armijo test uses the actual displacement. */
        double alpha = opt->proj_step ? 1.0 : a_pri;
        int accepted = 0;
        for (int ls = 0; ls < 40; ++ls) {
            double bar = 0, gdt = 0;
            for (size_t i = 0; i < nus; ++i) {
                const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
                double d = alpha * s.du[i];
                if (opt->proj_step) {
                    if (d < -tau_f * sl)
                        d = -tau_f * sl;
                    if (d > tau_f * su)
                        d = tau_f * su;
                }
                dut[i] = d;
                ut[i] = s.u[i] + d;
                gdt += rt[i] * d;
                bar += log(ut[i] - s.lb[i % NU]) + log(s.ub[i % NU] - ut[i]);
            }
            if (opt->proj_step) { /* dx_t by linearity */
                for (int i = 0; i < NX; ++i)
                    dxt[i] = 0.0;
                for (int k = 0; k < N; ++k)
                    for (int i = 0; i < NX; ++i) {
                        double a = 0;
                        for (int j = 0; j < NX; ++j)
                            a += s.Phi[i * NX + j] * dxt[NX * k + j];
                        for (int j = 0; j < NU; ++j)
                            a += s.Gam[i * NU + j] * dut[NU * k + j];
                        dxt[NX * (k + 1) + i] = a;
                    }
                for (size_t i = 0; i < nxs; ++i) {
                    xt[i] = s.x[i] + dxt[i];
                    gdt += s.q[i] * dxt[i];
                }
            } else {
                for (size_t i = 0; i < nxs; ++i)
                    xt[i] = s.x[i] + alpha * s.dx[i];
                gdt = alpha * gdw;
            }
            const double phit = eval_all(&s, xt, ut, 0, EPS_OF(mu)) - mu * bar;
            /* rounding-level slack so steps at the noise floor of phi are not rejected */
            if (phit <= phi0 + eta * gdt + 10 * 2.220446049250313e-16 * fabs(phi0)) {
                accepted = 1;
                break;
            }
            alpha *= 0.5;
            info->n_backtrack++;
        }
        if (getenv("NLP_ORACLE_DEBUG"))
            fprintf(stderr,
                    "it %d f %.10g mu %.2e ed %.2e ec %.2e delta %.1e apri %.3g adu %.3g alpha %.3g gdw %.3e acc %d\n",
                    iter, f, mu, e_dual, e_compl, delta, a_pri, a_du, alpha, gdw, accepted);
        if (!accepted) {
            status = 2;
            break;
        }
        memcpy(s.x, xt, sizeof(double) * nxs);
        for (size_t i = 0; i < nus; ++i) {
            const double sl0 = s.u[i] - s.lb[i % NU], su0 = s.ub[i % NU] - s.u[i];
            const double dzl = mu / sl0 - s.zl[i] - s.zl[i] / sl0 * s.du[i];
            const double dzu = mu / su0 - s.zu[i] + s.zu[i] / su0 * s.du[i];
            s.u[i] = ut[i];
            double zl = s.zl[i] + a_du * dzl, zu = s.zu[i] + a_du * dzu;
            const double sl = s.u[i] - s.lb[i % NU], su = s.ub[i % NU] - s.u[i];
            /* IPOPT eq. (16) safeguard */
            zl = fmax(fmin(zl, kappa_sigma * mu / sl), mu / (kappa_sigma * sl));
            zu = fmax(fmin(zu, kappa_sigma * mu / su), mu / (kappa_sigma * su));
            s.zl[i] = zl;
            s.zu[i] = zu;
        }
    }
    /* primal infeasibility of the returned point (drift of x += alpha dx) */
    double e_pri = 0;
    memcpy(xt, s.x, sizeof(double) * nxs);
    rollout(&s, s.u, xt);
    for (size_t i = 0; i < nxs; ++i)
        e_pri = fmax(e_pri, fabs(xt[i] - s.x[i]));
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < NX; ++i)
            w[NXU * k + i] = s.x[NX * k + i];
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < NU; ++i)
            w[NXU * k + NX + i] = s.u[NU * k + i];
    info->cost = eval_all(&s, s.x, s.u, 0, 0.0); /* the reference's objective (|s|, no smoothing) */
    info->iters = iter;
    info->status = status;
    info->kkt_dual = e_dual;
    info->kkt_primal = e_pri;
    info->kkt_compl = e_compl;
    info->mu = mu;
    free(buf);
    return status;
}

/* Batch wrapper used by the CPU baseline: instances are independent. */
void nlp_oracle_solve_batch(int B, int N, int K, double dt, const double *p, double *w,
                            const double lbu[NU], const double ubu[NU], const nlp_oracle_opts *opt,
                            nlp_oracle_info *info) {
    const size_t np = (size_t)nlp_oracle_np(N, K), nw = (size_t)nlp_oracle_nw(N);
    for (int b = 0; b < B; ++b)
        nlp_oracle_solve(N, K, dt, p + b * np, w + b * nw, lbu, ubu, opt, info + b, NULL);
}
