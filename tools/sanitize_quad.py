#!/usr/bin/env python
"""The quad solve kernel in its production shape -- 8 instances per warp, two warps per CTA, queue
refill -- on a problem small enough for compute-sanitizer: 2 resident warps per SM (296 warps, 2368
slots) for 2400 instances.  Prints how many converged; run under
  compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize_quad.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.update(AMPC_SOLVE_KERNEL="quad", AMPC_QUADS_PER_WARP="8", AMPC_QUAD_WARPS_PER_SM="2", AMPC_QUAD_CTA_WARPS="2")
import avoid_mpc_b200 as A  # noqa: E402
from helpers import make_instances  # noqa: E402

S = A.synth
N, K, B = 20, 16, 2400
inst = make_instances([500 + (b % 40) for b in range(40)], N, K, 10000)
rng = np.random.default_rng(7)
P = np.stack([inst["prefix"][b % 40] for b in range(B)])
W0 = np.stack([S.warm_start("ref", inst["x0"][b % 40], inst["ref"][b % 40], N) for b in range(B)])
for b in range(B):
    W0[b].reshape(-1)[10:] += 0.0
    for k in range(N):
        W0[b, 14 * k + 10:14 * k + 14] += rng.normal(0, 0.3, 4) * (b // 40) / 20.0
h = A.Handle(N=N, K=K, max_batch=B, max_points=16)
W, info = h.solve(P, W0)
h.close()
print("converged", float((info["status"] == 0).mean()), "iters max", int(info["iters"].max()), "launches ok")
