#!/usr/bin/env python
"""The command the ncu captures of round 2 are taken on: ONE call of the benchmark's step
(`--batch` instances, default 32 768 = one lane of bench.py's default step: index build + k-NN +
quad solve) after one warm-up call, the measured call between cudaProfilerStart/Stop
(run ncu with --profile-from-start off).  A tool, not the benchmark; nothing timed here."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avoid_mpc_b200 as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--npts", type=int, default=50000)
    a = ap.parse_args()
    D, S = A.defaults, A.synth
    N, K, B = 20, 16, a.batch
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    h = A.Handle(N=N, K=K, dt=D.BENCH_DT, max_batch=B, max_points=a.npts)
    h.set_solver_opts(tol=1e-8, max_iter=50)
    h.cloud_set_layout(S.image_shape(a.npts)[0])
    ids = list(range(B))
    for s0 in range(0, B, 1024):
        c = S.forest_clouds_torch(ids[s0:s0 + 1024], a.npts, dev)
        h.cloud_set_batch_dev(c, first_scene=s0, stream=st)
        torch.cuda.synchronize()
        del c
    x0_np, ref_np, _ = S.states_batch(ids, N, D.BENCH_DT)
    w0 = torch.tensor(np.stack([S.warm_start("ref", x0_np[b], ref_np[b], N) for b in range(B)]), device=dev)
    x0, ref = torch.tensor(x0_np, device=dev), torch.tensor(ref_np, device=dev)
    w = torch.empty_like(w0)
    for rep in range(2):
        w.copy_(w0)
        torch.cuda.synchronize()
        if rep == 1:
            torch.cuda.profiler.start()
        h.cloud_index_dev(0, B, stream=st)
        h.round_dev(B, x0, ref, w, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE, stream=st)
        torch.cuda.synchronize()
        if rep == 1:
            torch.cuda.profiler.stop()
    print("profile_step done, launches", h.launch_count())
    h.close()


if __name__ == "__main__":
    main()
