#!/usr/bin/env python
"""Mint the golden fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted).

knn_golden.npz  inputs + results of the REFERENCE's own KDTreeTwo<double>/nanoflann
                (oracle/_ref/libampc_ref_kdtree.so, compiled from /root/reference) on
                seeded clouds: the pin for both the C restatement and the CUDA path.
nlp_golden.npz  the fixture of the reference's only executable exercise of the NLP
                (tools/mpc_obstacle_casadi.py:448-498: 100-point cylinder at x=1, start
                (0,0,1), goal (5,0.1,1), N=30, K=3) plus a random-yaw instance; values of
                f / grad f / Hessian from an independent torch.autograd (float64)
                re-derivation of the script (tests/torch_nlp.py), and the converged optimum
                cross-checked with scipy trust-constr.  CasADi/IPOPT are not installable
                here, so these are NOT outputs of the reference solver (parity unpinned).
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import avoid_mpc_b200 as A  # noqa: E402
from oracle import oracle as O  # noqa: E402
import torch_nlp  # noqa: E402

D, S = A.defaults, A.synth


def knn_golden():
    assert O.ref_available(), "needs /root/reference to build oracle/_ref"
    out = {}
    cases = [("forest10k_k8", S.forest_cloud(1, 10000)[0], 8), ("forest3072_k3", S.forest_cloud(2, 3072)[0], 3),
             ("random5k_k16", S.random_cloud(5, 5000), 16), ("random300_k1", S.random_cloud(6, 300), 1)]
    nanc = S.random_cloud(7, 700)
    nanc[::5, 0] = np.nan
    cases.append(("nan700_k4", nanc, 4))
    for name, c, k in cases:
        _, ref, _ = S.states(11, 20)
        q = ref[:, :3].copy()
        if name.startswith("random"):
            q = np.random.default_rng(3).uniform([0, -3, 0.5], [20, 3, 3], (20, 3))
        idx, d2, cnt = O.RefTree(c).search(q, k)
        assert O.is_tie_free(O.filter_nan(c), q, k)
        out[name + "_cloud"] = c
        out[name + "_q"] = q
        out[name + "_idx"] = idx
        out[name + "_d2"] = d2
        out[name + "_cnt"] = cnt
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **out)


def cylinder_fixture():
    """tools/mpc_obstacle_casadi.py:448-498 with the shipped yaml (N=30, K=3)."""
    N, K = 30, 3
    p_init = np.array([0.0, 0.0, 1.0, 0, 0, 0, 0, 0, 0, 0])
    obstacles = []
    for obs_z in np.linspace(0, 3, 10):
        for theta in np.linspace(0, 2 * 3.14, 10):
            obstacles.append([0.1 * math.cos(theta) + 1.0, 0.1 * math.sin(theta), obs_z])
    obstacles = np.array(obstacles)
    p_goal = np.array([5.0, 0.1, 1.0, 0, 0, 0, 0, 0, 0, 0])
    dp = (p_goal - p_init) / N
    ref = np.stack([p_init + i * dp for i in range(N)])
    ob = np.zeros((N, K, 3))
    for i in range(N):
        d = np.linalg.norm(obstacles - ref[i, :3], axis=1)
        ob[i] = obstacles[np.argsort(d, kind="stable")[:K]]
    prefix = S.pack_prefix(p_init, ref, ob, p_goal)
    return N, K, 0.033, S.full_params(prefix), obstacles


def nlp_golden():
    import torch
    out = {}
    N, K, dt, p, obstacles = cylinder_fixture()
    rng = np.random.default_rng(0)
    w = rng.normal(0, 0.5, 10 + 14 * N)
    w[0:3] = [0, 0, 1]
    f, g, H = torch_nlp.f_grad_hess(N, K, w, p)
    out.update(cyl_N=N, cyl_K=K, cyl_dt=dt, cyl_p=p, cyl_w=w, cyl_f=f, cyl_grad=g, cyl_hess=H,
               cyl_obstacles=obstacles)
    # the reference harness' bounds (script lines 459-465): a_z in [-20, 20], others +-10
    lb, ub = np.array([-10, -10, -20, -10.0]), np.array([10, 10, 20, 10.0])
    w0 = np.zeros(10 + 14 * N)
    for k in range(N):
        w0[14 * k + 10:14 * k + 14] = [0, 0, 9.81, 0]
    ws, info = O.solve(N, K, dt, p, w0, lb, ub)
    assert info.status == 0
    out.update(cyl_lb=lb, cyl_ub=ub, cyl_w0=w0, cyl_wstar=ws, cyl_cost=info.cost)
    # a bench-shaped instance with non-zero reference yaw
    N2, K2 = 20, 8
    x0, ref, tgt = S.states(3, N2)
    ref[:, 3] = rng.uniform(-1, 1, N2)
    ob = ref[:, None, :3] + rng.normal(0, 0.6, (N2, K2, 3))
    p2 = S.full_params(S.pack_prefix(x0, ref, ob, tgt))
    w2 = S.warm_start("ref", x0, ref, N2) + rng.normal(0, 0.3, 10 + 14 * N2)
    f2, g2, H2 = torch_nlp.f_grad_hess(N2, K2, w2, p2)
    out.update(yaw_N=N2, yaw_K=K2, yaw_p=p2, yaw_w=w2, yaw_f=f2, yaw_grad=g2, yaw_hess=H2)
    np.savez_compressed(os.path.join(HERE, "nlp_golden.npz"), **out)


if __name__ == "__main__":
    knn_golden()
    nlp_golden()
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
