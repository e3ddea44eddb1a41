"""GPU parity: depth image -> Obstacle + Edge cloud through the C-ABI (ampc_depth_set_batch)
against the OpenCV-minted golden vectors and the numpy oracle -- bit-exact clouds, and the k-NN
answers on top of them bit-exact as well."""
import os

import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import depth_oracle as DO
from oracle import oracle as O

pytestmark = pytest.mark.gpu
S = A.synth
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "depth_golden.npz"))
NAMES = [str(n) for n in G["names"]]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", NAMES)
def test_depth_matches_opencv_golden(name):
    fx, fy, cx, cy, scale, p2m, dmin, dmax = G[name + "_params"]
    depth = G[name + "_depth"]
    H, W = int(depth.shape[0] / scale), int(depth.shape[1] / scale)
    h = A.Handle(N=4, K=4, max_batch=2, max_scenes=3, max_points=H * W, max_edge_points=H * W)
    h.set_camera(fx, fy, cx, cy, scale, p2m, dmin, dmax)
    T_obst = DO.matmul4(G[name + "_Twb"], DO.TBC)
    T_edge = DO.matmul4(G[name + "_Twc_prev"], DO.TBC)
    h.depth_set_batch(depth, T_obst, T_edge, first_scene=1)
    cloud, edge = h.cloud_get(1), h.cloud_get(1, A.capi.CLOUD_EDGE)
    assert cloud.shape == G[name + "_cloud"].shape and edge.shape == G[name + "_edge"].shape
    assert np.array_equal(_bits(cloud), _bits(G[name + "_cloud"]))
    assert np.array_equal(_bits(edge), _bits(G[name + "_edge"]))
    assert h.cloud_count(0) == 0 and h.cloud_count(2) == 0  # neighbours untouched
    h.close()


def _poses(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        yaw = rng.uniform(-0.3, 0.3)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        T[:3, 3] = rng.uniform([-1, -1, 1.0], [1, 1, 2.0])
        out.append(T)
    return out


@pytest.mark.parametrize("dtype", ["f32", "u16"])
def test_depth_batch_then_knn_matches_oracle(dtype):
    B, rows, cols, scale, Q, k = 6, 240, 320, 2.0, 20, 16
    cam = dict(fx=160.0, fy=160.0, cx=160.0, cy=120.0, resize_scale=scale,
               pixel2meter=0.001 if dtype == "u16" else 1.0, depth_min=0.1, depth_max=100.0)
    ocam = DO.Camera(**cam)
    depth = np.stack([S.forest_depth(40 + s, rows, cols, u16_scale=0.001 if dtype == "u16" else 0.0)
                      for s in range(B)])
    Twb = _poses(B, 1)
    T_obst = np.stack([DO.matmul4(T, DO.TBC) for T in Twb])
    T_edge = np.stack([DO.matmul4(DO.matmul4(T, DO.TBC), DO.TBC) for T in _poses(B, 2)])
    H, W = ocam.out_size(rows, cols)
    h = A.Handle(N=Q, K=k, max_batch=B, max_points=H * W, max_edge_points=H * W // 2)
    h.set_camera(**cam)
    h.depth_set_batch(depth, T_obst, T_edge)
    queries = np.stack([S.states(40 + s, Q)[1][:, :3] + Twb[s][:3, 3] - [0, 0, 1.5] for s in range(B)])
    idx, d2, _, cnt = h.knn(queries, k)
    eidx, ed2, _, ecnt = h.knn(queries[:, :3], 4, kind=A.capi.CLOUD_EDGE)
    for s in range(B):
        oc, oe = DO.process_depth(depth[s], ocam, T_obst[s], T_edge[s])
        assert np.array_equal(_bits(h.cloud_get(s)), _bits(oc))
        assert np.array_equal(_bits(h.cloud_get(s, A.capi.CLOUD_EDGE)), _bits(oe))
        ri, rd, rc = O.knn_bruteforce(oc, queries[s], k)
        assert (idx[s] == ri).all() and (d2[s] == rd).all() and (cnt[s] == rc).all()
        ri, rd, rc = O.knn_bruteforce(oe, queries[s, :3], 4)
        assert (eidx[s] == ri).all() and (ed2[s] == rd).all() and (ecnt[s] == rc).all()
    h.close()


def test_depth_full_image_keeps_patch_layout_and_round_runs():
    """No invalid pixel (back wall, no sky): the Obstacle cloud is the whole resized image, its
    index uses 8x8 image patches; the k-NN answers must not depend on that."""
    B, rows, cols, N, K = 4, 200, 250, 20, 16
    cam = dict(fx=125.0, fy=125.0, cx=125.0, cy=100.0, resize_scale=1.0)
    ocam = DO.Camera(**cam)
    depth = np.stack([S.forest_depth(60 + s, rows, cols, sky=False) for s in range(B)])
    Twb = np.eye(4)
    Twb[2, 3] = 1.5  # the pose the synthetic scene was rendered from
    T = np.stack([DO.matmul4(Twb, DO.TBC)] * B)
    h = A.Handle(N=N, K=K, max_batch=B, max_points=rows * cols, max_edge_points=rows * cols // 2)
    h.set_camera(**cam)
    h.depth_set_batch(depth, T)
    assert all(h.cloud_count(s) == rows * cols for s in range(B))
    x0s, refs, _ = zip(*[S.states(60 + s, N) for s in range(B)])
    queries = np.stack([r[:, :3] for r in refs])
    idx, d2, _, cnt = h.knn(queries, K)
    for s in range(B):
        oc, _ = DO.process_depth(depth[s], ocam, T[s], T[s])
        ri, rd, rc = O.knn_bruteforce(oc, queries[s], K)
        assert (idx[s] == ri).all() and (d2[s] == rd).all() and (cnt[s] == rc).all()
    # and a control round on top of the depth-built clouds converges
    W0 = np.stack([S.warm_start("ref", x0s[s], refs[s], N) for s in range(B)])
    w, info, _ = h.round(np.stack(x0s), np.stack(refs), W0, speed=A.defaults.SPEED)
    assert np.isfinite(w).all() and (info["status"] <= 1).all()
    h.close()


def test_depth_full_size_batch_properties():
    """BASELINE-size clouds (~50k points) for a batch of scenes: counts equal the oracle's for a
    sample, every record is finite and lies inside the depth range of its camera ray."""
    B, rows, cols = 32, 388, 516
    cam = dict(fx=258.0, fy=258.0, cx=258.0, cy=194.0, resize_scale=2.0)
    ocam = DO.Camera(**cam)
    depth = np.stack([S.forest_depth(80 + s, rows, cols) for s in range(B)])
    Twb = np.eye(4)
    Twb[2, 3] = 1.5
    T = np.stack([DO.matmul4(Twb, DO.TBC)] * B)
    H, W = ocam.out_size(rows, cols)
    h = A.Handle(N=20, K=16, max_batch=B, max_points=H * W, max_edge_points=H * W // 4)
    h.set_camera(**cam)
    h.depth_set_batch(depth, T)
    for s in (0, 13, 31):
        oc, oe = DO.process_depth(depth[s], ocam, T[s], T[s])
        assert np.array_equal(_bits(h.cloud_get(s)), _bits(oc))
        assert np.array_equal(_bits(h.cloud_get(s, A.capi.CLOUD_EDGE)), _bits(oe))
    for s in range(B):
        c = h.cloud_get(s)
        assert 10000 < len(c) <= H * W and np.isfinite(c).all()
        assert (c[:, 0] > 0.1).all() and (c[:, 0] < 100.1).all()  # body x = optical axis
    h.close()


def test_depth_errors():
    h = A.Handle(N=4, K=4, max_batch=1, max_points=64 * 48, max_edge_points=16)
    h.set_camera(resize_scale=10.0)
    T = DO.matmul4(np.eye(4), DO.TBC)
    with pytest.raises(A.AmpcError):  # Edge cloud outgrows its 16-point slot
        h.depth_set_batch(S.forest_depth(1, 480, 640), T)
    assert h.cloud_count(0, A.capi.CLOUD_EDGE) == 16
    with pytest.raises(A.AmpcError):  # resized image larger than max_points
        h.depth_set_batch(S.forest_depth(1, 960, 1280), T)
    with pytest.raises(TypeError):
        h.depth_set_batch(np.zeros((48, 64), np.float64), T)
    with pytest.raises(A.AmpcError):
        h.set_camera(resize_scale=0.5)
    # an all-sky image empties both clouds
    h.depth_set_batch(np.full((480, 640), 1e4, np.float32), T)
    assert h.cloud_count(0) == 0 and h.cloud_count(0, A.capi.CLOUD_EDGE) == 0
    idx, d2, _, cnt = h.knn(np.zeros((1, 4, 3)), 4)
    assert (cnt == 0).all()
    h.close()
