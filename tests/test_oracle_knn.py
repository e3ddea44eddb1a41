"""CPU: the k-NN oracle (C restatement) against the golden vectors minted from the
reference's own KDTreeTwo/nanoflann, and against that reference directly when present."""
import os

import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import oracle as O

S = A.synth
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "knn_golden.npz"))
CASES = sorted({k[:-6] for k in GOLD.files if k.endswith("_cloud")})


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(name):
    c, q = GOLD[name + "_cloud"], GOLD[name + "_q"]
    gi, gd, gc = GOLD[name + "_idx"], GOLD[name + "_d2"], GOLD[name + "_cnt"]
    k = gi.shape[1]
    for fn in (lambda: O.PortTree(c).search(q, k), lambda: O.knn_bruteforce(O.filter_nan(c), q, k)):
        idx, d2, cnt = fn()
        assert (cnt == gc).all()
        assert (idx == gi).all()
        assert (d2 == gd).all()  # bit-exact squared distances


def test_restatement_matches_reference_live():
    if not O.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(0)
    for seed, n, k in [(1, 1, 1), (2, 10, 3), (3, 11, 3), (4, 257, 8), (5, 4099, 16), (6, 20000, 32)]:
        c = S.random_cloud(seed, n)
        q = rng.uniform([0, -3, 0.5], [20, 3, 3], (7, 3))
        r = O.RefTree(c).search(q, k)
        p = O.PortTree(c).search(q, k)
        for a, b in zip(r, p):
            assert (a == b).all()
        if O.is_tie_free(c, q, k):
            bf = O.knn_bruteforce(c, q, k)
            for a, b in zip(r, bf):
                assert (a == b).all()


def test_count_rule_and_nan_filter():
    q = np.zeros((1, 3))
    for n, k, want in [(0, 4, 0), (3, 4, 3), (4, 4, 0), (5, 4, 4)]:  # kd_tree_two.h:112,117-124
        c = S.random_cloud(n + 1, n)
        assert O.PortTree(c).search(q, k)[2][0] == want
        assert O.knn_bruteforce(c, q, k)[2][0] == want
    c = S.random_cloud(3, 50)
    c[::3, 0] = np.nan
    f = O.filter_nan(c)
    assert f.shape[0] == int((~np.isnan(c[:, 0])).sum())
    assert (f[:, :3] == c[~np.isnan(c[:, 0]), :3]).all()  # order preserved (kd_tree_two.h:99-101)


def test_tie_detection():
    c = np.ones((8, 4), dtype=np.float32)
    c[:, 0] = [1, -1, 2, -2, 3, -3, 4, -4]
    assert not O.is_tie_free(c, np.array([[0.0, 1.0, 1.0]]), 3)
    assert O.is_tie_free(c, np.array([[0.1, 1.0, 1.0]]), 3)


def test_restatement_matches_sklearn_kdtree():
    """The reference's own Python harness queries obstacles with sklearn.neighbors.KDTree
    (tools/mpc_obstacle_casadi.py:10,457,482): same neighbours, same order, on a tie-free cloud
    (sklearn's distances are float64 from float64 copies of the float32 points: equal to 1e-12)."""
    from sklearn.neighbors import KDTree
    c = S.forest_cloud(77, 10000)[0]
    q = S.states(77, 20)[1][:, :3]
    k = 16
    assert O.is_tie_free(c, q, k)
    idx, d2, cnt = O.knn_bruteforce(c, q, k)
    dist, ind = KDTree(c[:, :3].astype(np.float64), leaf_size=10).query(q, k=k)
    assert (cnt == k).all() and (ind == idx).all()
    assert np.abs(dist ** 2 - d2).max() <= 1e-12 * d2.max()
