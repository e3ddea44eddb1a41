#!/usr/bin/env python
"""The bench's C1 step (index + round on 1024 forest scenes) a few times, nothing else: the
command ncu captures are taken on.  Prints per-stage device times.  A tool, not the benchmark."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avoid_mpc_b200 as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--npts", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--max-iter", type=int, default=50)
    a = ap.parse_args()
    D, S = A.defaults, A.synth
    N, K, B = 20, 16, a.batch
    dev = torch.device("cuda", 0)
    ids = list(range(B))
    clouds = S.forest_clouds_torch(ids, a.npts, dev)
    x0_np, ref_np, _ = S.states_batch(ids, N, D.BENCH_DT)
    w0 = torch.tensor(np.stack([S.warm_start("ref", x0_np[b], ref_np[b], N) for b in range(B)]), device=dev)
    x0, ref = torch.tensor(x0_np, device=dev), torch.tensor(ref_np, device=dev)
    lanes = []
    for _ in range(a.streams):
        h = A.Handle(N=N, K=K, dt=D.BENCH_DT, max_batch=B, max_points=a.npts)
        h.set_solver_opts(tol=1e-8, max_iter=a.max_iter)
        h.cloud_set_layout(S.image_shape(a.npts)[0])
        h.cloud_set_batch_dev(clouds, stream=torch.cuda.current_stream().cuda_stream)
        lanes.append((h, torch.cuda.Stream(), torch.empty_like(w0), torch.zeros((B, 48), dtype=torch.uint8, device=dev)))
    torch.cuda.synchronize()
    lanes[0][0].profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        if rep == 1:
            torch.cuda.synchronize()
            e0.record()
        for i in range(a.steps * a.streams):
            h, st, w, info = lanes[i % a.streams]
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                w.copy_(w0, non_blocking=True)
                h.cloud_index_dev(0, B, stream=st.cuda_stream)
                h.round_dev(B, x0, ref, w, info_dev=info, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE,
                            stream=st.cuda_stream)
        for _, st, _, _ in lanes:
            torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    prof = lanes[0][0].profile_get()
    inf = lanes[0][3].cpu().numpy().view(A.capi.INFO_DTYPE).reshape(B)
    it = inf["iters"]
    print(json.dumps({"streams": a.streams, "ms_per_step": e0.elapsed_time(e1) / (a.steps * a.streams),
                      "solves_per_s": B * a.steps * a.streams / e0.elapsed_time(e1) * 1e3,
                      "stage_ms": {k: v[0] / max(v[1], 1) for k, v in prof.items()},
                      "iters": {"mean": float(it.mean()), "p50": float(np.median(it)), "p90": float(np.percentile(it, 90)),
                                "max": int(it.max())}, "converged": float((inf["status"] == 0).mean())}))


if __name__ == "__main__":
    main()
