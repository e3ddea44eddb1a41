#!/usr/bin/env python
"""A/B of the two solve kernels against the CPU oracle on seeded forest instances (a debugging
tool, not the benchmark): per-instance status / iterations / cost / trajectory error."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import avoid_mpc_b200 as A  # noqa: E402
from helpers import make_instances, oracle_solve_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--N", type=int, default=20)
    ap.add_argument("--K", type=int, default=16)
    ap.add_argument("--warm", default="ref")
    ap.add_argument("--max-iter", type=int, default=100)
    ap.add_argument("--first", type=int, default=200)
    a = ap.parse_args()
    S = A.synth
    N, K, B = a.N, a.K, a.batch
    dt = 1.0 / N if N != 20 else 0.05
    inst = make_instances(range(a.first, a.first + B), N, K, 10000 if N == 20 else 3072)
    W0 = np.stack([S.warm_start(a.warm, inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    oW, ost, oit, ocost = oracle_solve_batch(N, K, dt, inst["params"], W0, max_iter=a.max_iter)
    res = {}
    for kern in ("quad", "warp"):  # AMPC_SOLVE_KERNEL forces one of the two
        os.environ["AMPC_SOLVE_KERNEL"] = kern
        h = A.Handle(N=N, K=K, dt=dt, max_batch=B, max_points=16)
        h.set_solver_opts(max_iter=a.max_iter)
        W, info = h.solve(inst["prefix"], W0)
        h.close()
        res[kern] = (W, info)
    print("inst | oracle st it cost | quad st it nreg nbt cost err | warp st it nreg nbt err")
    for b in range(B):
        qW, qi = res["quad"]
        wW, wi = res["warp"]
        print(f"{b:3d} | {ost[b]} {oit[b]:3d} {ocost[b]:.9g} | {qi['status'][b]} {qi['iters'][b]:3d} {qi['n_reg'][b]:2d} "
              f"{qi['n_backtrack'][b]:3d} {qi['cost'][b]:.9g} {np.abs(qW[b] - oW[b]).max():.2e} | "
              f"{wi['status'][b]} {wi['iters'][b]:3d} {wi['n_reg'][b]:2d} {wi['n_backtrack'][b]:3d} "
              f"{np.abs(wW[b] - oW[b]).max():.2e}")
    qW, qi = res["quad"]
    print("quad: converged", (qi["status"] == 0).mean(), "max err vs oracle (both converged)",
          np.abs(qW - oW).max(axis=1)[(qi["status"] == 0) & (ost == 0)].max() if ((qi["status"] == 0) & (ost == 0)).any() else None)


if __name__ == "__main__":
    main()
