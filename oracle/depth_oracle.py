"""TEST INFRASTRUCTURE ONLY: CPU restatement (numpy) of the reference's depth image ->
Obstacle cloud + Edge cloud step, FrameKDMap::ProcessDepth / BuildEdgeCloud
(roswrapper/ros/src/avoid_mpc/src/FrameKDMap.cpp:76-130,176-214).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.

The reference delegates three steps to OpenCV (a system dependency of the ROS package,
CMakeLists.txt find_package(OpenCV); not vendored): cv::resize (bilinear), cv::erode (3x3)
and cv::Canny(0.1, 0.3).  They are restated here from OpenCV 4.x's own (non-IPP) code path
-- imgproc/src/resize.cpp (resizeGeneric_ + HResizeLinear/VResizeLinear for CV_32F),
morph.dispatch.cpp (erode, default border = ignored) and canny.cpp (Sobel 3x3
BORDER_REPLICATE, L1 magnitude, TG22 fixed-point direction test) -- and PINNED against
opencv-python 4.13.0 with IPP switched off (tests/golden/make_depth_golden.py calls cv2
where the reference calls cv::; tests/test_oracle_depth.py compares bit for bit).  With IPP
on, cv2's bilinear kernel evaluates a + (b - a) f instead of a (1 - f) + b f and differs
from the native path by 1 ulp on ~13 % of the pixels; distro OpenCV (what a ROS install
links) is built without IPP, so the native path is the one restated.

All floating-point steps follow the C++ expression types of the reference (float vs
double, the order of the operations), without FMA contraction.
"""
from __future__ import annotations

import numpy as np

f32, f64 = np.float32, np.float64


class Camera:
    """perception block of config/mpc_parameters.yaml:57-70; fx..cy are the FULL-resolution
    values, divided by resize_scale as the constructor does (FrameKDMap.cpp:21-24)."""

    def __init__(self, fx=320.0, fy=320.0, cx=320.0, cy=240.0, resize_scale=10.0, pixel2meter=1.0,
                 depth_min=0.1, depth_max=100.0):
        self.resize_scale = float(resize_scale)
        self.fx, self.fy = fx / self.resize_scale, fy / self.resize_scale
        self.cx, self.cy = cx / self.resize_scale, cy / self.resize_scale
        self.pixel2meter, self.depth_min, self.depth_max = float(pixel2meter), float(depth_min), float(depth_max)

    def out_size(self, rows, cols):
        # mParamWidth = cols / mParamDepthScale (int <- double), :106-107
        return int(rows / self.resize_scale), int(cols / self.resize_scale)


TBC = np.array([[0, 0, 1, 0.05], [-1, 0, 0, 0.0], [0, -1, 0, 0.01], [0, 0, 0, 1.0]])  # yaml T_b_c


def matmul4(a, b):
    """4x4 product with the plain left-to-right sum of each entry."""
    c = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            s = a[i, 0] * b[0, j]
            for l in range(1, 4):
                s = s + a[i, l] * b[l, j]
            c[i, j] = s
    return c


def inv_depth(depth, cam):
    """GetInvDepthImg<T>, :76-89: float depth = float(pix) * pixel2meter (double product,
    stored as float); out of [min, max] -> 0, else float(1. / depth)."""
    d = (depth.astype(f32).astype(f64) * cam.pixel2meter).astype(f32).astype(f64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (1.0 / d).astype(f32)
    inv[(d < cam.depth_min) | (d > cam.depth_max)] = 0
    return inv


def _lin_coeffs(n_dst, n_src):
    scale = 1.0 / (n_dst / n_src)  # resize(): inv_scale = dsize/ssize; scale = 1./inv_scale
    ofs = np.zeros(n_dst, np.int64)
    w1 = np.zeros(n_dst, f32)
    for d in range(n_dst):
        f = f32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        ofs[d] = s
        w1[d] = f32(f - f32(s))
    return ofs, w1


def resize_linear(src, H, W):
    """cv::resize(src, dst, Size(W, H), INTER_MAX): the 4th argument is fx and is ignored when
    dsize is given, so this is INTER_LINEAR (:109).  Horizontal pass
    S[sx] (1 - fx) + S[sx + 1] fx with fx = 0 outside [0, w - 1]; vertical pass with the two
    row indices clamped and the weights kept (a NaN pixel therefore spreads to every output
    pixel whose 2x2 support touches it, even with weight 0).  Equal sizes are a plain copy."""
    h, w = src.shape
    if (H, W) == (h, w):
        return src.copy()
    xo, xf = _lin_coeffs(W, w)
    yo, yf = _lin_coeffs(H, h)
    rows = np.zeros((h, W), f32)
    for d in range(W):
        s, f = int(xo[d]), xf[d]
        if s < 0:
            s, f = 0, f32(0)
        if s >= w - 1:
            rows[:, d] = src[:, w - 1]
        else:
            rows[:, d] = (src[:, s] * f32(f32(1) - f)).astype(f32) + (src[:, s + 1] * f).astype(f32)
    out = np.zeros((H, W), f32)
    for d in range(H):
        s, f = int(yo[d]), yf[d]
        r0, r1 = min(max(s, 0), h - 1), min(max(s + 1, 0), h - 1)
        out[d] = (rows[r0] * f32(f32(1) - f)).astype(f32) + (rows[r1] * f).astype(f32)
    return out


def _unproject(cols, rows, depth, cam, T):
    """UV2Camera (:131-138) then T * pointCam, stored as float (pcl::PointXYZ)."""
    x = (cols - cam.cx) * depth / cam.fx
    y = (rows - cam.cy) * depth / cam.fy
    out = np.zeros((len(depth), 4), f32)
    for i in range(3):
        out[:, i] = (((T[i, 0] * x + T[i, 1] * y) + T[i, 2] * depth) + T[i, 3]).astype(f32)
    return out  # 16-byte records, w = 0


def obstacle_cloud(inv_small, cam, T):
    """Loop of ProcessDepth, :110-124: row-major over the resized image."""
    H, W = inv_small.shape
    inv = inv_small.astype(f64).ravel()
    rows, cols = np.divmod(np.arange(H * W), W)
    with np.errstate(divide="ignore", invalid="ignore"):
        depth = 1.0 / inv
        keep = ~(inv < 1e-2) & (depth > cam.depth_min) & (depth < cam.depth_max)
    return _unproject(cols[keep].astype(f64), rows[keep].astype(f64), depth[keep], cam, T)


def inflated_u8(inv_small, cam):
    """BuildEdgeCloud step 1, :181-194: uchar(1 / invDepth / (max - min) * 200.0f), 255 where
    invDepth <= 1e-2 (1 / invDepth is a float division; the rest is double)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        d = (f32(1) / inv_small).astype(f32).astype(f64) / (cam.depth_max - cam.depth_min) * 200.0
    ok = inv_small.astype(f64) > 1e-2
    out = np.full(inv_small.shape, 255, np.uint8)
    out[ok] = np.floor(d[ok]).astype(np.int64).astype(np.uint8)
    return out


def erode3(img):
    """cv::erode with a 3x3 kernel of ones; pixels outside the image are ignored."""
    h, w = img.shape
    p = np.pad(img, 1, constant_values=255)
    out = np.full_like(img, 255)
    for dy in range(3):
        for dx in range(3):
            out = np.minimum(out, p[dy:dy + h, dx:dx + w])
    return out


TG22 = 13573  # int(0.4142135623730950488016887242097 * 2^15 + 0.5)


def canny_0(img):
    """cv::Canny(img, 0.1, 0.3), aperture 3, L1 gradient (:196): both thresholds floor to 0,
    so every pixel with a non-zero gradient magnitude that survives the non-maximum
    suppression is a strong edge and the hysteresis pass adds nothing."""
    h, w = img.shape
    p = np.pad(img.astype(np.int32), 1, mode="edge")

    def s(dy, dx):
        return p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    gx = (s(-1, 1) + 2 * s(0, 1) + s(1, 1)) - (s(-1, -1) + 2 * s(0, -1) + s(1, -1))
    gy = (s(1, -1) + 2 * s(1, 0) + s(1, 1)) - (s(-1, -1) + 2 * s(-1, 0) + s(-1, 1))
    mag = np.abs(gx) + np.abs(gy)
    mp = np.pad(mag, 1)  # magnitude is 0 outside the image

    def m(dy, dx):
        return mp[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    x, y = np.abs(gx), np.abs(gy) << 15
    tg22 = x * TG22
    tg67 = tg22 + (x << 16)
    horiz = y < tg22
    vert = ~horiz & (y > tg67)
    diag = ~(horiz | vert)
    same = (gx ^ gy) >= 0
    e_h = (mag > m(0, -1)) & (mag >= m(0, 1))
    e_v = (mag > m(-1, 0)) & (mag >= m(1, 0))
    e_d = np.where(same, (mag > m(-1, -1)) & (mag > m(1, 1)), (mag > m(-1, 1)) & (mag > m(1, -1)))
    return (mag > 0) & ((horiz & e_h) | (vert & e_v) | (diag & e_d))


def edge_cloud(inv_small, cam, T_edge):
    """BuildEdgeCloud, :176-214.  T_edge is the transform the reference applies to the camera
    point: mCurFrame.Twc * mParamTbc, i.e. the PREVIOUS frame's Twb * Tbc times Tbc again."""
    er = erode3(inflated_u8(inv_small, cam))
    edges = canny_0(er)
    rows, cols = np.nonzero(edges)  # row-major
    depth = er[rows, cols].astype(f32).astype(f64) * (cam.depth_max - cam.depth_min) / 200.0
    keep = ~((depth > cam.depth_max) | (depth < cam.depth_min))
    return _unproject(cols[keep].astype(f64), rows[keep].astype(f64), depth[keep], cam, T_edge)


def process_depth(depth, cam, T_obstacle, T_edge):
    """depth (rows x cols, float32 metres or uint16) -> (Obstacle cloud, Edge cloud), both
    (n, 4) float32 with w = 0; the Edge cloud is empty when the Obstacle cloud is (:125-127)."""
    H, W = cam.out_size(*depth.shape)
    inv_small = resize_linear(inv_depth(depth, cam), H, W)
    cloud = obstacle_cloud(inv_small, cam, T_obstacle)
    if len(cloud) == 0:
        return cloud, np.zeros((0, 4), f32)
    return cloud, edge_cloud(inv_small, cam, T_edge)
