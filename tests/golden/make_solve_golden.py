#!/usr/bin/env python
"""Mint tests/golden/solve_golden.npz: converged optima of the reference NLP computed by an
INDEPENDENT solver in the reference's own formulation.

CasADi/IPOPT cannot be installed here, so the closest thing available is used:
scipy.optimize.minimize(method="trust-constr") -- an interior-point trust-region SQP
(Byrd-Hribar-Nocedal), the same class of method as IPOPT -- on the FULL multiple-shooting problem
exactly as tools/mpc_obstacle_casadi.py:156-220 poses it: variables w = [X_0,U_0,...,X_N], the
210 equality constraints g(w) = 0, bounds on the U entries only (src/HighLvlMpc.cpp:70-92), exact
gradient and exact Hessian of the un-smoothed objective.  Nothing of this repository's solver is
involved: only the oracle's function evaluations f / grad f / hess f, which are themselves pinned to
an autograd re-derivation of the script (tests/golden/nlp_golden.npz).

Instances: synthetic forest scenes at the three (N, K) shapes the tests use.
Instances on which trust-constr does not reach optimality <= 1e-6 are kept and flagged
(converged = 0): they are the scenes whose minimiser sits on the |v.n| kink.

usage: OMP_NUM_THREADS=1 python tests/golden/make_solve_golden.py   (a few minutes on 8 cores)
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

JOBS = ([(s, 20, 16, 10000, "ref") for s in range(40)] + [(s, 20, 8, 10000, "cold") for s in range(40, 48)] +
        [(s, 30, 3, 3072, "cold") for s in range(48, 52)])


def run(job):
    from scipy.optimize import Bounds, LinearConstraint, minimize

    import avoid_mpc_b200 as A
    from oracle import oracle as O
    D, S = A.defaults, A.synth
    sid, N, K, npts, warm = job
    dt = 0.05 if N == 20 else 1.0 / N
    lb, ub = D.u_bounds()
    c, _ = S.forest_cloud(sid, npts)
    x0, ref, tgt = S.states(sid, N, dt)
    idx, _, _ = O.knn_bruteforce(c, ref[:, :3], K)
    p = S.full_params(S.pack_prefix(x0, ref, c[idx][:, :, :3].astype(np.float64), tgt))
    w0 = S.warm_start(warm, x0, ref, N)
    Phi, Gam, gam = O.dyn_matrices(D.TAU, dt)
    nw, ng = O.nw(N), O.ng(N)
    J, rhs = np.zeros((ng, nw)), np.zeros(ng)
    J[0:10, 0:10] = np.eye(10)
    rhs[0:10] = x0
    for k in range(N):  # g_{k+1} = Phi X_k + Gam U_k + gam - X_{k+1}
        r = 10 * (k + 1)
        J[r:r + 10, 14 * k:14 * k + 10] = Phi
        J[r:r + 10, 14 * k + 10:14 * k + 14] = Gam
        J[r:r + 10, 14 * (k + 1):14 * (k + 1) + 10] = -np.eye(10)
        rhs[r:r + 10] = -gam
    lbw, ubw = np.full(nw, -np.inf), np.full(nw, np.inf)
    for k in range(N):
        lbw[14 * k + 10:14 * k + 14] = lb
        ubw[14 * k + 10:14 * k + 14] = ub

    def hess(z):
        Hx, Hu = O.hess_f(N, K, z, p)
        H = np.zeros((nw, nw))
        for k in range(N):
            a, b = 14 * (k + 1), 14 * k + 10
            H[a:a + 10, a:a + 10] = Hx[k]
            H[b:b + 4, b:b + 4] = np.diag(Hu)
        return H

    z0 = w0.copy()  # strictly inside the bounds, as IPOPT's bound_push does
    for k in range(N):
        z0[14 * k + 10:14 * k + 14] = np.clip(z0[14 * k + 10:14 * k + 14], lb + 1e-2 * (ub - lb), ub - 1e-2 * (ub - lb))
    t = time.time()
    res = minimize(lambda z: O.f(N, K, z, p), z0, jac=lambda z: O.grad_f(N, K, z, p), hess=hess, method="trust-constr",
                   constraints=[LinearConstraint(J, rhs, rhs)], bounds=Bounds(lbw, ubw),
                   options=dict(gtol=1e-9, xtol=1e-12, barrier_tol=1e-10, maxiter=800, initial_barrier_parameter=0.1))
    ok = res.optimality <= 1e-6 and res.constr_violation <= 1e-10
    return dict(job=job, p=p, w0=w0, w=res.x, cost=float(res.fun), conv=int(ok), opt=float(res.optimality),
                nit=int(res.nit), secs=time.time() - t)


def main():
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        out = pool.map(run, JOBS)
    d = {"n": np.array(len(out))}
    for i, r in enumerate(out):
        sid, N, K, npts, warm = r["job"]
        d[f"i{i}_meta"] = np.array([sid, N, K, npts, 1 if warm == "ref" else 0])
        d[f"i{i}_p"], d[f"i{i}_w0"], d[f"i{i}_w"] = r["p"], r["w0"], r["w"]
        d[f"i{i}_cost"] = np.array(r["cost"])
        d[f"i{i}_conv"] = np.array(r["conv"])
        d[f"i{i}_opt"] = np.array(r["opt"])
        print(r["job"], "conv", r["conv"], "opt %.1e" % r["opt"], "nit", r["nit"], "%.0fs" % r["secs"], flush=True)
    np.savez_compressed(os.path.join(HERE, "solve_golden.npz"), **d)
    print("converged:", sum(r["conv"] for r in out), "of", len(out))


if __name__ == "__main__":
    main()
