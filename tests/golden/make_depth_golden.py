#!/usr/bin/env python
"""Mint tests/golden/depth_golden.npz: depth image -> Obstacle cloud + Edge cloud as the
reference computes them (FrameKDMap::ProcessDepth / BuildEdgeCloud,
roswrapper/ros/src/avoid_mpc/src/FrameKDMap.cpp:76-130,176-214).

This script is a line-by-line transcription of those two functions that calls OpenCV
(opencv-python, IPP switched off = the code path of a distro OpenCV) exactly where the
reference calls cv::resize / cv::erode / cv::Canny, and runs the per-pixel loops as scalar
loops with the C++ types.  The numpy restatement in oracle/depth_oracle.py (which has no
OpenCV in it) and the CUDA path are both compared with these vectors bit for bit.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import avoid_mpc_b200 as A  # noqa: E402

f32 = np.float32
TBC = np.array([[0, 0, 1, 0.05], [-1, 0, 0, 0.0], [0, -1, 0, 0.01], [0, 0, 0, 1.0]])


def mat4(a, b):
    c = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            s = a[i, 0] * b[0, j]
            for l in range(1, 4):
                s = s + a[i, l] * b[l, j]
            c[i, j] = s
    return c


def reference_process_depth(depth, P, Twb, Twc_prev):
    fx, fy, cx, cy = (P[k] / P["resize_scale"] for k in ("fx", "fy", "cx", "cy"))  # ctor :21-24
    dmin, dmax, p2m = P["depth_min"], P["depth_max"], P["pixel2meter"]
    rows, cols = depth.shape
    # GetInvDepthImg :76-89
    inv = np.zeros((rows, cols), f32)
    for i in range(rows):
        for j in range(cols):
            d = f32(float(f32(depth[i, j])) * p2m)
            if float(d) < dmin or float(d) > dmax:
                inv[i, j] = 0.0
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    inv[i, j] = f32(np.float64(1.0) / np.float64(d))
    W, H = int(cols / P["resize_scale"]), int(rows / P["resize_scale"])
    small = cv2.resize(inv, (W, H), fx=float(cv2.INTER_MAX))  # :109, INTER_MAX sits in the fx slot
    T = mat4(Twb, TBC)

    def uv2cam_world(M, u, v, depth):
        pc = [(u - cx) * depth / fx, (v - cy) * depth / fy, depth, 1.0]
        return [f32(((M[i, 0] * pc[0] + M[i, 1] * pc[1]) + M[i, 2] * pc[2]) + M[i, 3] * pc[3]) for i in range(3)]

    cloud = []
    for row in range(H):
        for col in range(W):
            invd = float(small[row, col])
            if invd < 1e-2:
                continue
            d = 1 / invd
            if d > dmin and d < dmax:
                cloud.append(uv2cam_world(T, col, row, d) + [f32(0)])
    cloud = np.array(cloud, f32).reshape(-1, 4)
    if len(cloud) == 0:
        return cloud, np.zeros((0, 4), f32)
    # BuildEdgeCloud :176-214
    infl = np.zeros((H, W), np.uint8)
    for row in range(H):
        for col in range(W):
            invd = small[row, col]
            if float(invd) > 1e-2:
                infl[row, col] = int(float(f32(1) / invd) / (dmax - dmin) * float(f32(200.0))) & 0xFF
            else:
                infl[row, col] = 255
    er = cv2.erode(infl, np.ones((3, 3), np.uint8))
    canny = cv2.Canny(er, 0.1, 0.3)
    Te = mat4(Twc_prev, TBC)  # mCurFrame.Twc (previous frame) * mParamTbc, :208-209
    edge = []
    for row in range(H):
        for col in range(W):
            if canny[row, col] > 0:
                d = float(f32(er[row, col]))
                d = d * (dmax - dmin) / 200.0
                if d > dmax or d < dmin:
                    continue
                edge.append(uv2cam_world(Te, col, row, d) + [f32(0)])
    return cloud, np.array(edge, f32).reshape(-1, 4)


def pose(rng):
    yaw, pitch = rng.uniform(-np.pi, np.pi), rng.uniform(-0.2, 0.2)
    cz, sz, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    T = np.eye(4)
    T[:3, :3] = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    T[:3, 3] = rng.uniform([-5, -5, 0.5], [5, 5, 3])
    return T


def main():
    cv2.ipp.setUseIPP(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(2025)
    base = dict(fx=320.0, fy=320.0, cx=320.0, cy=240.0, resize_scale=10.0, pixel2meter=1.0, depth_min=0.1,
                depth_max=100.0)
    q = lambda img: (np.round(img * 1024) / 1024).astype(f32)  # keeps the fixture compressible
    cases = {}
    cases["sim_s10"] = (q(A.synth.forest_depth(1, 480, 640)), dict(base))
    cases["sim_s5"] = (q(A.synth.forest_depth(2, 240, 320)), dict(base, fx=160.0, fy=160.0, cx=160.0, cy=120.0,
                                                                    resize_scale=5.0))
    cases["u16_s2"] = (A.synth.forest_depth(3, 96, 128, u16_scale=0.001),
                       dict(base, fx=64.0, fy=64.0, cx=64.0, cy=48.0, resize_scale=2.0, pixel2meter=0.001))
    odd = A.synth.forest_depth(4, 97, 131).copy()
    odd[rng.random(odd.shape) < 0.02] = np.nan
    odd[rng.random(odd.shape) < 0.02] = np.inf
    odd[rng.random(odd.shape) < 0.02] = -1.0
    cases["nan_s1"] = (odd, dict(base, fx=65.5, fy=65.5, cx=65.5, cy=48.5, resize_scale=1.0))
    frac = A.synth.forest_depth(5, 100, 130).copy()
    frac[rng.random(frac.shape) < 0.01] = np.nan
    cases["frac_s2p5"] = (frac, dict(base, fx=65.0, fy=65.0, cx=65.0, cy=50.0,
                                                                   resize_scale=2.5))
    wall = A.synth.forest_depth(6, 120, 150, sky=False).copy()
    wall[rng.random(wall.shape) < 0.01] = np.nan
    cases["wall_s3"] = (wall, dict(base, fx=75.0, fy=75.0, cx=75.0, cy=60.0,
                                                                            resize_scale=3.0, depth_max=30.0))
    cases["empty"] = (np.full((40, 60), 1e4, f32), dict(base, fx=30.0, fy=30.0, cx=30.0, cy=20.0, resize_scale=2.0))
    out = {"names": np.array(sorted(cases))}
    for name, (depth, P) in cases.items():
        Twb, Twc_prev = pose(rng), mat4(pose(rng), TBC)
        cloud, edge = reference_process_depth(depth, P, Twb, Twc_prev)
        print(f"{name}: depth {depth.shape} {depth.dtype} -> cloud {len(cloud)} edge {len(edge)}")
        out[name + "_depth"] = depth
        out[name + "_params"] = np.array([P[k] for k in ("fx", "fy", "cx", "cy", "resize_scale", "pixel2meter",
                                                         "depth_min", "depth_max")])
        out[name + "_Twb"] = Twb
        out[name + "_Twc_prev"] = Twc_prev
        out[name + "_cloud"] = cloud
        out[name + "_edge"] = edge
    np.savez_compressed(os.path.join(HERE, "depth_golden.npz"), **out)


if __name__ == "__main__":
    main()
