"""Seeded synthetic inputs of SURVEY.md §8d: depth-camera "forest" clouds in the
reference's pcl::PointXYZ layout (float32 x,y,z,pad; 16-byte stride), quadrotor
states, reference paths and warm starts.

The cloud emulates FrameKDMap::ProcessDepth (FrameKDMap.cpp:110-125): a pinhole
planar-depth image back-projected through K, T_b_c and the body pose.  The edge
cloud marks pixels whose 4-neighbour depth jump exceeds 0.5 m (a stand-in for the
reference's Canny edge cloud, FrameKDMap.cpp:176-214).

numpy implementation for single scenes (tests); `forest_clouds_torch` generates a
whole batch on a torch device for the benchmark.  This is data generation, not
part of the measured path.
"""
from __future__ import annotations

import numpy as np

from . import defaults as D

SEED0 = 20250310
N_CYL = 40


def image_shape(npts: int) -> tuple[int, int]:
    """W x H with W*H == npts and W:H ~ 5:4 (50k -> 250x200, 10k -> 125x80)."""
    known = {50000: (250, 200), 10000: (125, 80), 20000: (160, 125), 100000: (400, 250),
             200000: (500, 400), 500000: (800, 625), 1000000: (1250, 800)}
    if npts in known:
        return known[npts]
    h = int(np.sqrt(npts / 1.25))
    while h > 1 and npts % h:
        h -= 1
    return npts // h, h


def scene_params(scene_id: int):
    rng = np.random.Generator(np.random.PCG64(SEED0 + int(scene_id)))
    cx = rng.uniform(2.0, 30.0, N_CYL)
    cy = rng.uniform(-10.0, 10.0, N_CYL)
    cr = rng.uniform(0.1, 0.5, N_CYL)
    return rng, cx, cy, cr


def _depth_image(W, H, cx, cy, cr, origin, back_wall=40.0):
    """Planar depth (metres along the optical axis == world +x) of the forest scene."""
    fx = fy = 0.5 * W
    u = (np.arange(W) + 0.0 - W / 2.0) / fx
    v = (np.arange(H) + 0.0 - H / 2.0) / fy
    xn, yn = np.meshgrid(u, v)  # (H, W): cam x right, cam y down
    ox, oy, oz = origin
    # world ray: p(t) = o + t * (1, -xn, -yn)   (T_b_c: body x = cam z, body y = -cam x, body z = -cam y)
    t = np.full((H, W), (back_wall - ox) if back_wall is not None else np.inf)  # back wall x = 40, or sky
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(yn > 1e-9, oz / yn, np.inf)  # ground z = 0
    t = np.minimum(t, tg)
    a = 1.0 + xn * xn
    for j in range(len(cx)):
        bx, by = ox - cx[j], oy - cy[j]
        b = 2.0 * (bx - xn * by)
        c = bx * bx + by * by - cr[j] * cr[j]
        disc = b * b - 4 * a * c
        with np.errstate(invalid="ignore"):
            tc = (-b - np.sqrt(disc)) / (2 * a)
        tc = np.where((disc > 0) & (tc > 0), tc, np.inf)
        t = np.minimum(t, tc)
    if back_wall is None:
        return np.maximum(t, 0.1), xn, yn
    return np.clip(t, 0.1, 100.0), xn, yn


def forest_cloud(scene_id: int, npts: int = 50000, body_pos=(0.0, 0.0, D.HEIGHT)):
    """Returns (cloud16 [npts,4] float32, edge16 [m,4] float32)."""
    rng, cx, cy, cr = scene_params(scene_id)
    W, H = image_shape(npts)
    origin = (body_pos[0] + D.T_B_C[0, 3], body_pos[1] + D.T_B_C[1, 3], body_pos[2] + D.T_B_C[2, 3])
    depth, xn, yn = _depth_image(W, H, cx, cy, cr, origin)
    jit = rng.uniform(-1e-3, 1e-3, (H, W, 3))
    pts = np.empty((H, W, 4), dtype=np.float32)
    pts[..., 0] = origin[0] + depth + jit[..., 0]
    pts[..., 1] = origin[1] - xn * depth + jit[..., 1]
    pts[..., 2] = origin[2] - yn * depth + jit[..., 2]
    pts[..., 3] = 1.0
    jump = np.zeros((H, W), dtype=bool)
    dx = np.abs(np.diff(depth, axis=1)) > 0.5
    dy = np.abs(np.diff(depth, axis=0)) > 0.5
    jump[:, 1:] |= dx
    jump[:, :-1] |= dx
    jump[1:, :] |= dy
    jump[:-1, :] |= dy
    cloud = pts.reshape(-1, 4)
    edge = np.ascontiguousarray(pts[jump])
    return np.ascontiguousarray(cloud), edge


def forest_depth(scene_id: int, rows: int = 480, cols: int = 640, noise: float = 0.02, sky: bool = True,
                 u16_scale: float = 0.0, body_pos=(0.0, 0.0, D.HEIGHT)) -> np.ndarray:
    """Full-resolution depth image of the forest scene as the simulator publishes it: 32FC1
    metres with additive N(0, noise) (airsim_ros_wrapper.cpp:1267-1280),
    or CV_16UC1 in units of `u16_scale` metres when u16_scale > 0 (0 = no return).  With `sky`
    the rays that hit nothing carry 1e4 m (beyond depth_max) instead of a back wall."""
    rng, cx, cy, cr = scene_params(scene_id)
    origin = (body_pos[0] + D.T_B_C[0, 3], body_pos[1] + D.T_B_C[1, 3], body_pos[2] + D.T_B_C[2, 3])
    depth, _, _ = _depth_image(cols, rows, cx, cy, cr, origin, back_wall=None if sky else 40.0)
    hit = np.isfinite(depth)
    depth = np.where(hit, depth + noise * rng.standard_normal((rows, cols)), 1e4)
    if u16_scale > 0:
        return np.where(hit, np.clip(np.rint(depth / u16_scale), 0, 65535), 0).astype(np.uint16)
    return depth.astype(np.float32)


def random_cloud(seed: int, npts: int, lo=(-5, -5, 0), hi=(25, 5, 4)):
    """Uniform random cloud (tie-free with overwhelming probability)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.ones((npts, 4), dtype=np.float32)
    out[:, :3] = rng.uniform(lo, hi, (npts, 3)).astype(np.float32)
    return out


def states(scene_id: int, N: int = 20, dt: float = D.BENCH_DT, speed: float = D.SPEED):
    """Returns (x0[10], ref[N,10], target[10]) per SURVEY §8d."""
    rng = np.random.Generator(np.random.PCG64(SEED0 + 7919 * 1000003 + int(scene_id)))
    p0 = np.array([0.0, 0.0, D.HEIGHT]) + rng.uniform(-0.5, 0.5, 3)
    v0 = np.array([rng.uniform(0, 10), rng.uniform(-1, 1), rng.uniform(-0.5, 0.5)])
    a0 = rng.uniform(-2, 2, 3)
    x0 = np.concatenate([p0, [0.0], v0, a0])
    ref = np.zeros((N, 10))
    for i in range(N):  # GetInitPath("forward"), AvoidanceStateMachine.cpp:29-34,53
        ref[i, 0:3] = p0 + np.array([(i + 1) * speed * dt, 0.0, 0.0])
        ref[i, 4] = speed
    return x0, ref, make_target(ref, x0[0], speed, N * dt)


def make_target(ref: np.ndarray, pos_x: float, speed: float, T: float) -> np.ndarray:
    """GetRefStates target rule, AvoidanceStateMachine.cpp:250-255."""
    tgt = ref[-1].copy()
    dX = speed * T - max(0.0, tgt[0] - pos_x)
    tgt[0] += max(0.0, dX)
    tgt[1] = 0.0
    return tgt


def pack_prefix(x0, ref, obst, target) -> np.ndarray:
    """GetRefStates (AvoidanceStateMachine.cpp:236-257): [x0 | ref N*10 | obst N*K*3 | target]."""
    return np.concatenate([np.ravel(x0), np.ravel(ref), np.ravel(obst), np.ravel(target)]).astype(np.float64)


def full_params(prefix, gains=D.GAINS, tau=D.TAU, weights=D.WEIGHTS, radius=D.DRONE_RADIUS) -> np.ndarray:
    """ObstacleAvoidanceMPC::Solve tail packing (HighLvlMpc.cpp:97-108)."""
    return np.concatenate([prefix, gains, tau, weights, [radius]]).astype(np.float64)


def warm_start(kind: str, x0, ref, N: int) -> np.ndarray:
    """'cold' = the reference's all-zero mNlpW0 (HighLvlMpc.cpp:25-27,35,41-42);
    'ref' = reference-path roll-out with hover thrust."""
    w = np.zeros(10 + 14 * N)
    if kind == "cold":
        return w
    if kind != "ref":
        raise ValueError(kind)
    w[0:10] = x0
    for k in range(N):
        w[14 * k + 10: 14 * k + 14] = [0.0, 0.0, 9.81, 0.0]
        w[14 * (k + 1): 14 * (k + 1) + 10] = ref[k]
    return w


# ------------------------------------------------------------ torch batch ----
def forest_clouds_torch(scene_ids, npts: int, device, chunk: int = 16):
    """Batched generator: returns a (S, npts, 4) float32 torch tensor on `device`.
    Scene geometry comes from the same PCG64 streams as `forest_cloud`; the per-pixel
    jitter comes from a torch generator (so clouds are not bit-identical to the numpy
    path, which no test relies on)."""
    import torch

    S = len(scene_ids)
    W, H = image_shape(npts)
    out = torch.empty((S, npts, 4), dtype=torch.float32, device=device)
    fx = 0.5 * W
    u = (torch.arange(W, device=device, dtype=torch.float64) - W / 2.0) / fx
    v = (torch.arange(H, device=device, dtype=torch.float64) - H / 2.0) / fx
    yn, xn = torch.meshgrid(v, u, indexing="ij")
    xn = xn.reshape(1, -1, 1)
    yn = yn.reshape(1, -1, 1)
    ox, oy, oz = D.T_B_C[0, 3], D.T_B_C[1, 3], D.HEIGHT + D.T_B_C[2, 3]
    gen = torch.Generator(device=device)
    for s0 in range(0, S, chunk):
        ids = scene_ids[s0:s0 + chunk]
        prm = [scene_params(i)[1:] for i in ids]
        cx = torch.tensor(np.stack([p[0] for p in prm]), device=device).unsqueeze(1)
        cy = torch.tensor(np.stack([p[1] for p in prm]), device=device).unsqueeze(1)
        cr = torch.tensor(np.stack([p[2] for p in prm]), device=device).unsqueeze(1)
        bx, by = ox - cx, oy - cy
        a = 1.0 + xn * xn
        b = 2.0 * (bx - xn * by)
        c = bx * bx + by * by - cr * cr
        disc = b * b - 4 * a * c
        tc = (-b - torch.sqrt(disc.clamp_min(0))) / (2 * a)
        tc = torch.where((disc > 0) & (tc > 0), tc, torch.full_like(tc, float("inf")))
        t = tc.min(dim=2).values
        t = torch.minimum(t, torch.full_like(t, 40.0 - ox))
        tg = torch.where(yn[..., 0] > 1e-9, oz / yn[..., 0], torch.full_like(yn[..., 0], float("inf")))
        t = torch.minimum(t, tg).clamp(0.1, 100.0)
        gen.manual_seed(SEED0 + int(ids[0]))
        jit = (torch.rand((len(ids), npts, 3), generator=gen, device=device, dtype=torch.float64) - 0.5) * 2e-3
        o = out[s0:s0 + len(ids)]
        o[..., 0] = (ox + t + jit[..., 0]).float()
        o[..., 1] = (oy - xn[..., 0] * t + jit[..., 1]).float()
        o[..., 2] = (oz - yn[..., 0] * t + jit[..., 2]).float()
        o[..., 3] = 1.0
    return out


def states_batch(scene_ids, N: int = 20, dt: float = D.BENCH_DT, speed: float = D.SPEED):
    xs, refs, tgts = [], [], []
    for i in scene_ids:
        x0, ref, tgt = states(i, N, dt, speed)
        xs.append(x0)
        refs.append(ref)
        tgts.append(tgt)
    return np.stack(xs), np.stack(refs), np.stack(tgts)
