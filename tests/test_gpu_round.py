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


def test_full_size_batch_properties():
    """BASELINE config C1 at full size (B=1024, N=20, K=16, 50k-point clouds): size-independent
    properties of every instance + exact parity with the oracle on a sample of instances."""
    import torch
    N, K, B, npts = 20, 16, 1024, 50000
    dev = torch.device("cuda", 0)
    ids = list(range(B))
    clouds = S.forest_clouds_torch(ids, npts, dev)
    x0, ref, _ = S.states_batch(ids, N, 0.05)
    W0 = np.stack([S.warm_start("ref", x0[b], ref[b], N) for b in range(B)])
    h = A.Handle(N=N, K=K, max_batch=B, max_points=npts)
    h.cloud_set_layout(S.image_shape(npts)[0])
    h.cloud_set_batch_dev(clouds, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    W, info, replan = h.round(x0, ref, W0, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE)
    idx, d2, pts, cnt = h.knn(ref[:, :, :3], K)
    assert (cnt == K).all()
    assert (np.diff(d2, axis=2) >= 0).all()                      # sorted
    # every reported neighbour distance is the true distance of the reported index
    sample = np.random.default_rng(0).choice(B, 12, replace=False)
    lb, ub = D.u_bounds()
    Phi, Gam, gam = h.dynamics()
    X = np.stack([W[:, 14 * k:14 * k + 10] for k in range(N + 1)], axis=1)
    U = np.stack([W[:, 14 * k + 10:14 * k + 14] for k in range(N)], axis=1)
    assert np.abs(X[:, 0] - x0).max() == 0.0                     # X_0 = x0
    defect = np.einsum("ij,bkj->bki", Phi, X[:, :-1]) + np.einsum("ij,bkj->bki", Gam, U) + gam - X[:, 1:]
    assert np.abs(defect).max() < 1e-9                            # dynamics
    assert (U >= lb - 1e-12).all() and (U <= ub + 1e-12).all()    # control box
    conv = info["status"] == 0
    assert conv.mean() > 0.98
    assert (info["kkt_dual"][conv] <= 1e-8).all() and (info["kkt_compl"][conv] <= 1e-8).all()
    for b in sample:
        c = clouds[b].cpu().numpy()
        ri, rd, rc = O.knn_bruteforce(c, ref[b][:, :3], K)
        assert (idx[b] == ri).all() and (d2[b] == rd).all()      # bit-exact k-NN at full size
        tgt = S.make_target(ref[b], x0[b][0], D.SPEED, N * 0.05)
        p = S.full_params(S.pack_prefix(x0[b], ref[b], pts[b], tgt))
        assert abs(O.f(N, K, W[b], p) - info["cost"][b]) <= 1e-9 * max(1.0, abs(info["cost"][b]))
        ow, oi = O.solve(N, K, 0.05, p, W0[b], lb, ub)
        if oi.status == 0 and info["status"][b] == 0:
            assert np.abs(W[b] - ow).max() < 1e-4
    h.close()


def test_edge_guesses_best_of():
    """BASELINE config C2 in small: scenes x G Edge-tree initial guesses, best-cost reduction.
    Guess g replaces waypoint 0 of the reference path by the g-th nearest Edge point
    (PlanWapionts, AvoidanceStateMachine.cpp:259-281, generalised; g = 0 is the reference)."""
    N, K, n_scenes, G, npts = 20, 16, 6, 8, 10000
    h = A.Handle(N=N, K=K, max_batch=n_scenes * G, max_scenes=n_scenes, max_points=npts, max_edge_points=npts)
    clouds, edges, x0s, refs = [], [], [], []
    for s in range(n_scenes):
        c, e = S.forest_cloud(700 + s, npts)
        h.cloud_set(s, c)
        h.cloud_set(s, e, kind=A.capi.CLOUD_EDGE)
        x0, ref, _ = S.states(700 + s, N)
        clouds.append(c), edges.append(e), x0s.append(x0), refs.append(ref)
    # G nearest Edge points of waypoint 0 (Edge cloud, k = G), through the C-ABI
    q0 = np.stack([r[0, :3] for r in refs])[:, None, :]
    eidx, ed2, epts, ecnt = h.knn(q0, G, kind=A.capi.CLOUD_EDGE)
    for s in range(n_scenes):
        ri, rd, rc = O.knn_bruteforce(edges[s], q0[s], G)
        assert (eidx[s] == ri).all() and (ed2[s] == rd).all()
    scene_of = np.repeat(np.arange(n_scenes, dtype=np.int32), G)
    X0 = np.stack([x0s[s] for s in scene_of])
    REF = np.stack([refs[s].copy() for s in scene_of])
    for s in range(n_scenes):
        for g in range(G):
            REF[s * G + g, 0, :3] = epts[s, 0, g]
    W0 = np.zeros((n_scenes * G, 10 + 14 * N))
    W, info, _ = h.round(X0, REF, W0, scene_of=scene_of)
    arg, best = h.best_of(info, n_scenes, G)
    warg, wbest = A.shard.best_of_scenes(info["cost"], info["status"], G)
    assert (arg == warg).all() and (best == wbest).all()
    # the one-call form with shared queries (N-1+G Obstacle queries per scene instead of N*G)
    # must give the very same instances, bit for bit
    W2, info2, arg2, best2 = h.guess_round(np.stack(x0s), np.stack(refs), W0, G)
    assert (W2 == W).all() and (info2["cost"] == info["cost"]).all() and (info2["iters"] == info["iters"]).all()
    assert (arg2 == arg).all() and (best2 == best).all()
    # the winning instance of each scene against the oracle
    lb, ub = D.u_bounds()
    for s in range(n_scenes):
        b = s * G + arg[s]
        idx, d2, cnt = O.knn_bruteforce(clouds[s], REF[b][:, :3], K)
        ob = clouds[s][idx][:, :, :3].astype(np.float64)
        tgt = S.make_target(REF[b], X0[b][0], D.SPEED, N * 0.05)
        p = S.full_params(S.pack_prefix(X0[b], REF[b], ob, tgt))
        ow, oi = O.solve(N, K, 0.05, p, W0[b], lb, ub)
        if oi.status == 0 and info["status"][b] == 0:
            assert np.abs(W[b] - ow).max() < 1e-4
            assert abs(oi.cost - best[s]) <= 1e-6 * max(1.0, abs(best[s]))
    h.close()
