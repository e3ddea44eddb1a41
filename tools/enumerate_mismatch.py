#!/usr/bin/env python
"""Enumerate the instances of tests/test_gpu_solve.py::test_solve_matches_oracle on which the CUDA
path and the oracle disagree (status or trajectory), per solve kernel, with what each side did."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import avoid_mpc_b200 as A  # noqa: E402
from helpers import make_instances, oracle_solve_batch  # noqa: E402

S = A.synth
for N, K, npts, kind in [(20, 16, 10000, "cold"), (20, 16, 10000, "ref"), (20, 8, 10000, "ref"), (30, 3, 3072, "cold")]:
    B = 48
    dt = 1.0 / N if N != 20 else 0.05
    inst = make_instances(range(200, 200 + B), N, K, npts)
    W0 = np.stack([S.warm_start(kind, inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    oW, ost, oit, ocost = oracle_solve_batch(N, K, dt, inst["params"], W0)
    for kern in ("warp", "quad"):
        os.environ["AMPC_SOLVE_KERNEL"] = kern
        h = A.Handle(N=N, K=K, dt=dt, max_batch=B, max_points=16)
        W, info = h.solve(inst["prefix"], W0)
        h.close()
        err = np.abs(W - oW).max(axis=1)
        for b in range(B):
            if info["status"][b] != ost[b] or err[b] >= 1e-6:
                print(f"N={N} K={K} {kind} {kern} scene {200 + b}: gpu st {info['status'][b]} it {info['iters'][b]} cost {info['cost'][b]:.9g} | "
                      f"oracle st {ost[b]} it {oit[b]} cost {ocost[b]:.9g} | linf {err[b]:.2e}")
print("done")
