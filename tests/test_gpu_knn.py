"""GPU parity: exact k-NN through the C-ABI vs the CPU oracle (bit-exact indices and dist2)."""
import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import oracle as O

pytestmark = pytest.mark.gpu
S = A.synth


def _check(h, clouds, queries, k, scene_of=None):
    idx, d2, pts, cnt = h.knn(queries, k, scene_of=scene_of)
    B, Q = queries.shape[:2]
    for b in range(B):
        c = clouds[b if scene_of is None else scene_of[b]]
        cf = O.filter_nan(c)
        ri, rd, rc = O.knn_bruteforce(cf, queries[b], k)
        assert (cnt[b] == rc).all()
        assert (idx[b] == ri).all(), f"index mismatch instance {b}"
        assert (d2[b] == rd).all(), "dist2 must be bit-exact"
        for q in range(Q):
            m = rc[q]
            assert (pts[b, q, :m] == cf[ri[q, :m], :3].astype(np.float64)).all()
            assert (pts[b, q, m:] == 1e4).all()


@pytest.mark.parametrize("npts,k", [(10000, 8), (50000, 16), (3072, 3), (10000, 1), (2000, 32)])
def test_knn_forest_matches_oracle(npts, k):
    B, Q = 6, 20
    h = A.Handle(N=Q, K=min(k, 32), max_batch=B, max_points=npts)
    clouds = [S.forest_cloud(100 + s, npts)[0] for s in range(B)]
    for s, c in enumerate(clouds):
        h.cloud_set(s, c)
    queries = np.stack([S.states(100 + s, Q)[1][:, :3] for s in range(B)])
    _check(h, clouds, queries, k)
    h.close()


def test_knn_matches_reference_tree_when_tie_free(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    B, Q, k, npts = 4, 20, 16, 20000
    h = A.Handle(N=Q, K=k, max_batch=B, max_points=npts)
    for s in range(B):
        c = S.random_cloud(7 + s, npts)
        h.cloud_set(s, c)
        q = np.random.default_rng(s).uniform([0, -3, 0.5], [20, 3, 3], (Q, 3))
        assert O.is_tie_free(c, q, k)
        idx, d2, _, cnt = h.knn(q[None], k, scene_of=np.array([s], dtype=np.int32))
        ri, rd, rc = O.RefTree(c).search(q, k)
        assert (idx[0] == ri).all() and (d2[0] == rd).all() and (cnt[0] == rc).all()
    h.close()


def test_knn_edge_cases():
    Q, k = 5, 4
    h = A.Handle(N=Q, K=k, max_batch=8, max_points=4096)
    rng = np.random.default_rng(0)
    q = rng.uniform(-1, 1, (1, Q, 3))
    # empty cloud -> no results (kd_tree_two.h:112)
    h.cloud_set(0, np.zeros((0, 4), dtype=np.float32))
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([0], dtype=np.int32))
    assert (cnt == 0).all() and (idx == -1).all() and np.isinf(d2).all() and (pts == 1e4).all()
    # fewer points than k -> n results; exactly k points -> 0 results (kd_tree_two.h:117-124)
    for s, n in ((1, 3), (2, 4), (3, 5)):
        c = S.random_cloud(s, n, lo=(-1, -1, -1), hi=(1, 1, 1))
        h.cloud_set(s, c)
        idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([s], dtype=np.int32))
        ri, rd, rc = O.knn_bruteforce(c, q[0], k)
        assert (cnt[0] == rc).all() and (idx[0] == ri).all() and (d2[0] == rd).all()
        assert (cnt[0] == (n if n < k else (k if n > k else 0))).all()
    # NaN x dropped and the rest re-indexed (kd_tree_two.h:99-101); ragged size; stride 12
    c = S.random_cloud(9, 1237, lo=(-1, -1, -1), hi=(1, 1, 1))
    c[::7, 0] = np.nan
    c[5, 1] = np.nan  # NaN y stays in the cloud but can never be a neighbour
    h.cloud_set(4, c)
    assert h.cloud_count(4) == int((~np.isnan(c[:, 0])).sum())
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([4], dtype=np.int32))
    ri, rd, rc = O.knn_bruteforce(O.filter_nan(c), q[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    c3 = np.ascontiguousarray(S.random_cloud(11, 999)[:, :3])
    h.cloud_set(5, c3)
    qq = rng.uniform([0, -3, 0], [20, 3, 3], (1, Q, 3))
    idx, d2, pts, cnt = h.knn(qq, k, scene_of=np.array([5], dtype=np.int32))
    ri, rd, rc = O.knn_bruteforce(c3, qq[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    h.close()


def test_knn_duplicate_points_canonical_order():
    """Exact ties are returned in (dist2, index) order."""
    c = np.ones((64, 4), dtype=np.float32)
    c[:, :3] = np.array([1.0, 2.0, 3.0], dtype=np.float32)
    c[40:, 0] = 5.0
    h = A.Handle(N=1, K=8, max_batch=1, max_points=64)
    h.cloud_set(0, c)
    idx, d2, _, cnt = h.knn(np.zeros((1, 1, 3)), 8)
    assert (idx[0, 0] == np.arange(8)).all() and cnt[0, 0] == 8
    h.close()


def test_knn_single_instance_uses_segments_and_large_cloud():
    """B = 1 spreads one cloud over many CTAs (multi-segment merge path); 1M points."""
    npts, Q, k = 1000000, 20, 16
    h = A.Handle(N=Q, K=k, max_batch=1, max_points=npts)
    c = S.random_cloud(3, npts, lo=(0, -20, 0), hi=(40, 20, 10))
    h.cloud_set(0, c)
    q = np.random.default_rng(1).uniform([0, -5, 0.5], [30, 5, 3], (1, Q, 3))
    idx, d2, _, cnt = h.knn(q, k)
    ri, rd, rc = O.knn_bruteforce(c, q[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    # size-independent property: results are sorted and are true distances of the indices
    assert (np.diff(d2[0], axis=1) >= 0).all()
    h.close()


@pytest.mark.parametrize("row_w", [0, 125, 64, 9, 1000])
def test_layout_hint_changes_nothing_but_speed(row_w):
    """8x8 image-patch tiles (organised-cloud hint) vs linear tiles: identical results, also with
    NaN records (which shift the grid) and a width that does not divide the cloud."""
    npts, Q, k = 10000, 20, 16
    h = A.Handle(N=Q, K=k, max_batch=3, max_points=npts)
    h.cloud_set_layout(row_w)
    clouds = [S.forest_cloud(40, npts)[0], S.forest_cloud(41, 9973)[0], S.forest_cloud(42, npts)[0].copy()]
    clouds[2][5::97, 0] = np.nan
    for s, c in enumerate(clouds):
        h.cloud_set(s, c)
    queries = np.stack([S.states(40 + s, Q)[1][:, :3] for s in range(3)])
    _check(h, clouds, queries, k)
    h.close()


def _shuffled(c, seed):
    """The same points in arbitrary storage order (what KDTreeTwo::InitializeNew may be handed)."""
    return np.ascontiguousarray(c[np.random.default_rng(seed).permutation(len(c))])


@pytest.mark.parametrize("npts,k", [(50000, 16), (10000, 8), (777, 3), (64, 1), (65, 32)])
def test_sort_layout_on_shuffled_clouds_matches_oracle(npts, k):
    """AMPC_LAYOUT_SORT: tiles over the Morton-bucketed copy; indices are still those of the
    caller's record order, indices and dist2 bit-exact against brute force."""
    B, Q = 5, 20
    h = A.Handle(N=Q, K=min(k, 32), max_batch=B, max_points=npts)
    h.cloud_set_layout(A.capi.LAYOUT_SORT)
    clouds = [_shuffled(S.forest_cloud(300 + s, npts)[0], s) for s in range(B)]
    clouds[1] = clouds[1][: max(1, npts - 37)]          # ragged batch
    h.cloud_set_batch(_pad_batch(clouds, npts), counts=np.array([len(c) for c in clouds], dtype=np.int32))
    queries = np.stack([S.states(300 + s, Q)[1][:, :3] for s in range(B)])
    _check(h, clouds, queries, k)
    # the caller's cloud is untouched by the bucketing
    assert (h.cloud_get(2)[:, :3] == clouds[2][:, :3]).all()
    h.close()


def _pad_batch(clouds, npts):
    out = np.zeros((len(clouds), npts, 4), dtype=np.float32)
    for s, c in enumerate(clouds):
        out[s, : len(c)] = c
    return out


def test_sort_layout_edge_cases():
    """Empty cloud, n < k, n == k, NaN x records (dropped before the bucketing, indices refer to
    the filtered cloud), NaN / inf in y or z, all points identical (degenerate box), and a switch
    of the same slot back to the unorganised layout."""
    Q, k = 6, 4
    h = A.Handle(N=Q, K=k, max_batch=8, max_points=5000)
    h.cloud_set_layout(A.capi.LAYOUT_SORT)
    rng = np.random.default_rng(5)
    q = rng.uniform(-1, 1, (1, Q, 3))
    h.cloud_set(0, np.zeros((0, 4), dtype=np.float32))
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([0], dtype=np.int32))
    assert (cnt == 0).all() and (idx == -1).all() and np.isinf(d2).all() and (pts == 1e4).all()
    for s, n in ((1, 3), (2, 4), (3, 5)):
        c = S.random_cloud(20 + s, n, lo=(-1, -1, -1), hi=(1, 1, 1))
        h.cloud_set(s, c)
        idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([s], dtype=np.int32))
        ri, rd, rc = O.knn_bruteforce(c, q[0], k)
        assert (cnt[0] == rc).all() and (idx[0] == ri).all() and (d2[0] == rd).all()
    c = S.random_cloud(31, 4321, lo=(-1, -1, -1), hi=(1, 1, 1))
    c[::5, 0] = np.nan
    c[7, 1] = np.nan
    c[9, 2] = np.inf
    c[11, 1] = -np.inf
    h.cloud_set(4, c)
    cf = O.filter_nan(c)
    assert h.cloud_count(4) == len(cf)
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([4], dtype=np.int32))
    ri, rd, rc = O.knn_bruteforce(cf, q[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    same = np.ones((300, 4), dtype=np.float32)
    h.cloud_set(5, same)
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([5], dtype=np.int32))
    assert (idx[0] == np.arange(k)).all()       # exact ties in (dist2, index) order
    # same slot, now unorganised: the layout is per scene and per build
    h.cloud_set_layout(0)
    c2 = S.random_cloud(33, 2000, lo=(-1, -1, -1), hi=(1, 1, 1))
    h.cloud_set(4, c2)
    idx, d2, pts, cnt = h.knn(q, k, scene_of=np.array([4], dtype=np.int32))
    ri, rd, rc = O.knn_bruteforce(c2, q[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    # scene 5 keeps its bucketed index while scene 4 is unorganised: mixed batch
    qq = np.concatenate([q, q])
    idx, d2, pts, cnt = h.knn(qq, k, scene_of=np.array([4, 5], dtype=np.int32))
    assert (idx[0] == ri).all() and (idx[1] == np.arange(k)).all()
    h.close()


def test_sort_layout_large_cloud_and_segments():
    """1M shuffled points, B = 1 (multi-segment path over the bucketed copy), uniform random
    cloud = the worst case for the flat index."""
    npts, Q, k = 1000000, 20, 16
    h = A.Handle(N=Q, K=k, max_batch=1, max_points=npts)
    h.cloud_set_layout(A.capi.LAYOUT_SORT)
    c = S.random_cloud(4, npts, lo=(0, -20, 0), hi=(40, 20, 10))
    h.cloud_set(0, c)
    q = np.random.default_rng(2).uniform([0, -5, 0.5], [30, 5, 3], (1, Q, 3))
    idx, d2, _, cnt = h.knn(q, k)
    ri, rd, rc = O.knn_bruteforce(c, q[0], k)
    assert (idx[0] == ri).all() and (d2[0] == rd).all()
    h.close()


def test_sort_layout_round_equals_unorganised_round():
    """The control round on shuffled clouds: same neighbours, hence the same prefixes and
    bit-identical trajectories, whatever the layout."""
    N, K, B, npts = 20, 16, 8, 20000
    clouds = [_shuffled(S.forest_cloud(500 + s, npts)[0], 50 + s) for s in range(B)]
    x0, ref, _ = S.states_batch(list(range(500, 500 + B)), N, A.defaults.BENCH_DT)
    W0 = np.stack([S.warm_start("ref", x0[b], ref[b], N) for b in range(B)])
    res = []
    for lay in (0, A.capi.LAYOUT_SORT):
        h = A.Handle(N=N, K=K, max_batch=B, max_points=npts)
        h.cloud_set_layout(lay)
        h.cloud_set_batch(np.stack(clouds))
        W, info, replan = h.round(x0, ref, W0)
        res.append((W, info["cost"].copy(), replan.copy()))
        h.close()
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all()


def test_one_row_organised_cloud_with_a_wide_row_hint():
    """A row hint wider than the cloud is tall (here: ONE image row of 20000 records in slots sized
    for 65536 points) would need ceil(w/8) = 2500 tiles of 8 points against 2112 tile slots; such a
    scene must fall back to the linear tiling instead of writing boxes past its slots."""
    Q, k, npts = 20, 16, 20000
    h = A.Handle(N=Q, K=k, max_batch=3, max_points=65536)
    h.cloud_set_layout(20000)
    clouds = [S.forest_cloud(40 + s, npts)[0] for s in range(3)]
    for s, c in enumerate(clouds):
        h.cloud_set(s, c)
    queries = np.stack([S.states(40 + s, Q)[1][:, :3] for s in range(3)])
    _check(h, clouds, queries, k)      # scene 1's boxes would have been overwritten by scene 0's
    # a hint that fits (250 columns x 80 rows) on the same handle still gives the same answers
    h.cloud_set_layout(250)
    for s, c in enumerate(clouds):
        h.cloud_set(s, c)
    _check(h, clouds, queries, k)
    h.close()


def test_cloud_set_batch_reads_only_each_scenes_own_points():
    """ampc_cloud_set_batch on a buffer that ends right after the LAST scene's points (shorter than
    the largest scene): counts[s] records of scene s are all the call may read."""
    import ctypes
    Q, k = 8, 4
    counts = np.array([3000, 4096, 1000], dtype=np.int32)
    stride = 4096 * 16
    h = A.Handle(N=Q, K=k, max_batch=3, max_points=4096)
    clouds = [S.random_cloud(60 + s, int(n)) for s, n in enumerate(counts)]
    # page-aligned allocation whose last valid byte is the last point of the last scene
    nbytes = 2 * stride + int(counts[2]) * 16
    buf = np.zeros(nbytes, dtype=np.uint8)
    for s, c in enumerate(clouds):
        raw = np.ascontiguousarray(c, dtype=np.float32).view(np.uint8).ravel()
        buf[s * stride:s * stride + raw.size] = raw
    rc = h.L.ampc_cloud_set_batch(h.h, A.capi.CLOUD_OBSTACLE, 0, 3, buf.ctypes.data_as(ctypes.c_void_p),
                                  counts.ctypes.data, stride, 16)
    assert rc == 0
    q = np.stack([np.random.default_rng(s).uniform(-1, 1, (Q, 3)) for s in range(3)])
    _check(h, clouds, q, k)
    h.close()
