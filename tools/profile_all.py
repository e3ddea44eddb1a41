#!/usr/bin/env python
"""The command the ncu full captures of a round are taken on: one short pass over every
kernel family of the library at the benchmark's sizes.

  part 1  the C1 step (1024 forest scenes x 50 000 points): cloud_index, knn_search, ipm_solve
  part 2  1024 depth frames 200x250 -> Obstacle + Edge clouds: depth_obstacle, edge_grad, edge_cloud
  part 3  the same clouds in shuffled storage order with AMPC_LAYOUT_SORT: cloud_sort (+ index, search)

Each part runs once for warm-up (allocation, attribute set-up) and once more between
cudaProfilerStart/Stop: run ncu with `--profile-from-start off` to capture exactly the second
pass.  A tool, not the benchmark; under ncu no timing here means anything."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avoid_mpc_b200 as A  # noqa: E402


_capturing = False


def capture(on):
    global _capturing
    torch.cuda.synchronize()
    if on and not _capturing:
        torch.cuda.profiler.start()
    elif not on and _capturing:
        torch.cuda.profiler.stop()
    _capturing = on


def main():
    parts = sys.argv[1] if len(sys.argv) > 1 else "123"
    D, S = A.defaults, A.synth
    N, K, B, npts = 20, 16, 1024, 50000
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    ids = list(range(B))
    clouds = S.forest_clouds_torch(ids, npts, dev)
    x0_np, ref_np, _ = S.states_batch(ids, N, D.BENCH_DT)
    w0 = torch.tensor(np.stack([S.warm_start("ref", x0_np[b], ref_np[b], N) for b in range(B)]), device=dev)
    x0, ref = torch.tensor(x0_np, device=dev), torch.tensor(ref_np, device=dev)
    w = torch.empty_like(w0)
    # part 1
    h = A.Handle(N=N, K=K, dt=D.BENCH_DT, max_batch=B, max_points=npts)
    h.set_solver_opts(tol=1e-8, max_iter=50)
    h.cloud_set_layout(S.image_shape(npts)[0])
    h.cloud_set_batch_dev(clouds, stream=st)
    for rep in range(2):
        w.copy_(w0)
        capture(rep == 1)
        h.cloud_index_dev(0, B, stream=st)
        h.round_dev(B, x0, ref, w, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE, stream=st)
        capture(False)
    h.close()
    if "2" not in parts and "3" not in parts:
        print("profile_all done (part 1)")
        return
    # part 2
    rows, cols, distinct = 200, 250, 64
    hd = A.Handle(N=N, K=K, dt=D.BENCH_DT, max_batch=B, max_points=rows * cols, max_edge_points=rows * cols // 4)
    hd.set_camera(fx=cols / 2, fy=cols / 2, cx=cols / 2, cy=rows / 2, resize_scale=1.0)
    base = np.stack([S.forest_depth(s, rows, cols, sky=False) for s in range(distinct)])
    depth = torch.from_numpy(base).to(dev).repeat(B // distinct, 1, 1).contiguous()
    Twb = np.eye(4)
    Twb[2, 3] = D.HEIGHT
    T = torch.from_numpy(np.tile((Twb @ D.T_B_C).reshape(1, 16), (B, 1))).to(dev)
    for rep in range(2):
        capture(rep == 1)
        hd.depth_set_batch_dev(depth, T, None, stream=st)
        capture(False)
    hd.close()
    del depth
    # part 3
    perm = torch.randperm(npts, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    shuffled = clouds[:, perm].contiguous()
    hs = A.Handle(N=N, K=K, dt=D.BENCH_DT, max_batch=B, max_points=npts)
    hs.cloud_set_layout(A.capi.LAYOUT_SORT)
    hs.cloud_set_batch_dev(shuffled, stream=st)
    q = torch.tensor(np.ascontiguousarray(ref_np[:, :, :3]), device=dev)
    idx = torch.empty((B, N, K), dtype=torch.int32, device=dev)
    d2 = torch.empty((B, N, K), dtype=torch.float64, device=dev)
    cnt = torch.empty((B, N), dtype=torch.int32, device=dev)
    for rep in range(2):
        capture(rep == 1)
        hs.cloud_index_dev(0, B, stream=st)
        hs.knn_dev(q, K, idx, d2, None, cnt, stream=st)
        capture(False)
    hs.close()
    print("profile_all done")


if __name__ == "__main__":
    main()
