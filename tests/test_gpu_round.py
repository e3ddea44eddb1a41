"""GPU: one control-tick round (k-NN at the waypoints + prefix packing + solve) and the
best-of-G reduction through the C-ABI, against the oracle composed on the CPU."""
import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import oracle as O
from helpers import make_instances, oracle_solve_batch

pytestmark = pytest.mark.gpu
D, S = A.defaults, A.synth


def test_round_equals_knn_pack_solve():
    N, K, B, npts = 20, 16, 24, 10000
    inst = make_instances(range(500, 500 + B), N, K, npts)
    h = A.Handle(N=N, K=K, max_batch=B, max_points=npts)
    h.cloud_set_batch(np.stack(inst["clouds"]))
    W0 = np.stack([S.warm_start("ref", inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    W, info, replan = h.round(inst["x0"], inst["ref"], W0, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE)
    # solving the CPU-packed prefixes gives bit-identical results: the device-side
    # GetRefStates packing + k-NN obstacle block equal the CPU composition
    W2, info2 = h.solve(inst["prefix"], W0)
    assert np.abs(W - W2).max() == 0.0 and (info["iters"] == info2["iters"]).all()
    oW, ost, oit, ocost = oracle_solve_batch(N, K, 0.05, inst["params"], W0)
    both = (info["status"] == 0) & (ost == 0)
    assert both.mean() > 0.9
    assert (np.abs(W - oW).max(axis=1)[both] < 1e-4).mean() >= 0.98
    # needReplan (AvoidanceStateMachine.cpp:228-231)
    for b in range(B):
        _, d2, cnt = O.knn_bruteforce(inst["clouds"][b], inst["ref"][b][:, :3], K)
        want = int(any(cnt[q] == 0 or np.sqrt(d2[q, 0]) <= D.SAFETY_DISTANCE for q in range(N)))
        assert replan[b] == want
    h.close()


def test_target_rule_and_shared_scenes():
    """pos_x feeds the target rule (:250-255); several instances may share one scene's cloud."""
    N, K, npts = 20, 8, 10000
    inst = make_instances([600, 601], N, K, npts)
    h = A.Handle(N=N, K=K, max_batch=4, max_scenes=2, max_points=npts)
    h.cloud_set_batch(np.stack(inst["clouds"]))
    scene_of = np.array([0, 1, 1, 0], dtype=np.int32)
    x0 = inst["x0"][scene_of]
    ref = inst["ref"][scene_of]
    pos_x = x0[:, 0] + np.array([0.0, 0.0, 3.0, -2.0])
    W0 = np.zeros((4, 10 + 14 * N))
    W, info, _ = h.round(x0, ref, W0, scene_of=scene_of, pos_x=pos_x)
    lb, ub = D.u_bounds()
    for b in range(4):
        tgt = S.make_target(ref[b], pos_x[b], D.SPEED, N * 0.05)
        p = S.full_params(S.pack_prefix(x0[b], ref[b], inst["obst"][scene_of[b]], tgt))
        ow, oi = O.solve(N, K, 0.05, p, W0[b], lb, ub)
        if oi.status == 0 and info["status"][b] == 0:
            assert np.abs(W[b] - ow).max() < 1e-4
    h.close()


def test_best_of_matches_host_semantics():
    rng = np.random.default_rng(4)
    n_scenes, G = 37, 32
    info = np.zeros(n_scenes * G, dtype=A.capi.INFO_DTYPE)
    info["cost"] = rng.uniform(1, 1000, n_scenes * G)
    info["status"] = rng.choice([0, 0, 0, 1, 2, 3], n_scenes * G)
    info["status"][5 * G:6 * G] = 3  # a scene with no usable guess
    h = A.Handle(N=20, K=16, max_batch=n_scenes * G, max_scenes=1, max_points=64)
    arg, best = h.best_of(info, n_scenes, G)
    warg, wbest = A.shard.best_of_scenes(info["cost"], info["status"], G)
    assert (arg == warg).all() and (best == wbest).all() and arg[5] == -1
    h.close()
