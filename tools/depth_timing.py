#!/usr/bin/env python
"""Times ampc_depth_set_batch_dev (depth images resident in HBM -> both clouds + indices) on a
batch of synthetic depth frames.  Prints one JSON line; a tool, not the benchmark."""
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
    ap.add_argument("--scenes", type=int, default=512)
    ap.add_argument("--rows", type=int, default=388)
    ap.add_argument("--cols", type=int, default=516)
    ap.add_argument("--scale", type=float, default=2.0)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    B, H, W = a.scenes, int(a.rows / a.scale), int(a.cols / a.scale)
    h = A.Handle(N=20, K=16, max_batch=B, max_points=H * W, max_edge_points=H * W // 2)
    h.set_camera(fx=a.cols / 2, fy=a.cols / 2, cx=a.cols / 2, cy=a.rows / 2, resize_scale=a.scale)
    base = np.stack([A.synth.forest_depth(s, a.rows, a.cols) for s in range(16)])
    depth = torch.from_numpy(base).cuda().repeat((B + 15) // 16, 1, 1)[:B].contiguous()
    Twb = np.eye(4)
    Twb[2, 3] = 1.5
    T = torch.from_numpy(np.tile((Twb @ A.defaults.T_B_C).reshape(1, 16), (B, 1))).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        h.depth_set_batch_dev(depth, T, None, stream=st)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    l0 = h.launch_count()
    ev[0].record()
    for _ in range(a.steps):
        h.depth_set_batch_dev(depth, T, None, stream=st)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.steps
    n_obst = sum(h.cloud_count(s) for s in range(0, B, max(1, B // 8))) / len(range(0, B, max(1, B // 8)))
    in_bytes = B * a.rows * a.cols * 4
    out_bytes = B * n_obst * 16
    print(json.dumps({"scenes": B, "image": [a.rows, a.cols], "resized": [H, W], "ms_per_batch": ms,
                      "frames_per_s": B / ms * 1e3, "avg_obstacle_points": n_obst,
                      "launches_per_batch": (h.launch_count() - l0) / a.steps,
                      "algorithmic_GBps": (in_bytes / a.scale ** 2 * min(4, a.scale ** 2) / 4 + out_bytes) / ms / 1e6}))


if __name__ == "__main__":
    main()
