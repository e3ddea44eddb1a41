"""GPU: the device-side control tick (ampc_tick_batch) against the same loop driven from the
host round by round (AvoidanceStateMachine.cpp:328-344 semantics)."""
import numpy as np
import pytest

import avoid_mpc_b200 as A

pytestmark = pytest.mark.gpu
D, S = A.defaults, A.synth


@pytest.fixture(params=["warp", "quad"], autouse=True)
def solve_kernel(request, monkeypatch):
    """Both solve kernels (the active-instance mask of the tick reaches both)."""
    monkeypatch.setenv("AMPC_SOLVE_KERNEL", request.param)
    return request.param


def _host_tick(h, x0, ref, w0, scene_of, max_rounds, safety, have_edge):
    B, N = x0.shape[0], h.N
    ref, w = ref.copy(), w0.copy()
    active = np.ones(B, dtype=bool)
    rounds = np.zeros(B, dtype=np.int32)
    safe = np.ones(B, dtype=np.int32)
    info = np.zeros(B, dtype=A.capi.INFO_DTYPE)
    for it in range(max_rounds):
        # PlanWapionts (:259-281)
        q0 = ref[:, 0:1, :3].copy()
        _, d1, _, c1 = h.knn(q0, 1, scene_of=scene_of)
        if have_edge:
            _, _, ep, ec = h.knn(q0, 1, scene_of=scene_of, kind=A.capi.CLOUD_EDGE)
        for b in np.nonzero(active)[0]:
            near = np.sqrt(d1[b, 0, 0]) if c1[b, 0] > 0 else np.finfo(np.float64).max
            s = 1
            if near <= safety:
                if have_edge and ec[b, 0] > 0:
                    ref[b, 0, :3] = ep[b, 0, 0]
                else:
                    s = 0
            safe[b] = s
        # ProcessWaypoints + gate + Solve (:204-257,333,337)
        wr, ir, rp = h.round(x0, ref, w, scene_of=scene_of, speed=D.SPEED, safety_distance=safety)
        for b in np.nonzero(active)[0]:
            if (not rp[b]) and it > 0 and safe[b]:
                active[b] = False
                continue
            rounds[b] += 1
            w[b] = wr[b]
            info[b] = ir[b]
            for i in range(N):
                ref[b, i] = wr[b][14 * i:14 * i + 10]
    return w, ref, info, rounds, safe


@pytest.mark.parametrize("have_edge", [True, False])
def test_tick_matches_host_driven_rounds(have_edge):
    N, K, B, npts = 20, 16, 24, 10000
    h = A.Handle(N=N, K=K, max_batch=B, max_points=npts, max_edge_points=npts if have_edge else 0)
    h.cloud_set_layout(S.image_shape(npts)[0])
    x0s, refs = [], []
    for s in range(B):
        c, e = S.forest_cloud(800 + s, npts)
        h.cloud_set(s, c)
        if have_edge:
            h.cloud_set(s, e, kind=A.capi.CLOUD_EDGE)
        x0, ref, _ = S.states(800 + s, N)
        x0s.append(x0), refs.append(ref)
    x0s, refs = np.stack(x0s), np.stack(refs)
    # push a few first waypoints right onto an obstacle point so that PlanWapionts acts
    for b in range(0, B, 5):
        c, _ = S.forest_cloud(800 + b, npts)
        refs[b, 0, :3] = c[np.argmin(np.linalg.norm(c[:, :3] - refs[b, 0, :3], axis=1)), :3] + 0.05
    W0 = np.stack([S.warm_start("ref", x0s[b], refs[b], N) for b in range(B)])
    so = np.arange(B, dtype=np.int32)
    w, ref, info, rounds, safe = h.tick(x0s, refs, W0, scene_of=so, speed=D.SPEED,
                                        safety_distance=D.SAFETY_DISTANCE, max_rounds=3)
    hw, href, hinfo, hrounds, hsafe = _host_tick(h, x0s, refs, W0, so, 3, D.SAFETY_DISTANCE, have_edge)
    assert (rounds == hrounds).all() and (safe == hsafe).all()
    assert (rounds >= 1).all() and (rounds <= 3).all()
    assert len(set(rounds.tolist())) > 1          # both early exits and full ticks occur
    assert np.abs(w - hw).max() == 0.0 and np.abs(ref - href).max() == 0.0
    assert (info["iters"] == hinfo["iters"]).all()
    h.close()


def _oracle_tick(clouds, edges, x0, ref, w0, max_rounds, safety, N, K):
    """The TASK loop of AvoidanceStateMachine::Step (:328-344) composed from the ORACLE only:
    reference-order k-NN (oracle brute force == the compiled nanoflann on tie-free clouds) for
    PlanWapionts (:259-281) and ProcessWaypoints (:204-235), GetRefStates packing (:236-257)
    and the oracle interior-point solve per round."""
    from oracle import oracle as O
    lb, ub = D.u_bounds()
    B = x0.shape[0]
    ref, w = ref.copy(), w0.copy()
    rounds = np.zeros(B, dtype=np.int32)
    safe = np.ones(B, dtype=np.int32)
    conv = np.ones(B, dtype=bool)
    for b in range(B):
        cf = O.filter_nan(clouds[b])
        ef = O.filter_nan(edges[b]) if edges is not None else None
        for it in range(max_rounds):
            # PlanWapionts: GetNearestDistance(ref[0]) then the nearest Edge point
            _, d1, c1 = O.knn_bruteforce(clouds[b], ref[b, 0:1, :3], 1)
            near = np.sqrt(d1[0, 0]) if c1[0] > 0 else np.finfo(np.float64).max
            s = 1
            if near <= safety:
                ec = 0
                if ef is not None and len(ef) > 0:
                    ei, _, ec_ = O.knn_bruteforce(edges[b], ref[b, 0:1, :3], 1)
                    ec = ec_[0]
                if ec > 0:
                    ref[b, 0, :3] = ef[ei[0, 0], :3].astype(np.float64)
                else:
                    s = 0
            safe[b] = s
            # ProcessWaypoints: K-NN per waypoint, padding, needReplan
            idx, d2, cnt = O.knn_bruteforce(clouds[b], ref[b, :, :3], K)
            ob = np.full((N, K, 3), 1e4)
            replan = False
            for q in range(N):
                ob[q, :cnt[q]] = cf[idx[q, :cnt[q]], :3].astype(np.float64)
                if cnt[q] == 0 or np.sqrt(d2[q, 0]) <= safety:
                    replan = True
            if (not replan) and it > 0 and s:
                break
            rounds[b] += 1
            tgt = S.make_target(ref[b], x0[b][0], D.SPEED, N * D.BENCH_DT)
            p = S.full_params(S.pack_prefix(x0[b], ref[b], ob, tgt))
            ow, oinfo = O.solve(N, K, D.BENCH_DT, p, w[b], lb, ub)
            conv[b] &= oinfo.status == 0
            w[b] = ow
            for i in range(N):
                ref[b, i] = ow[14 * i:14 * i + 10]
    return w, ref, rounds, safe, conv


def test_tick_matches_oracle_composed_tick():
    """Device tick vs the tick composed from the oracle's k-NN and the oracle's solve, round by
    round (no GPU kernel on the checking side): same number of rounds, same isSafety, and the
    final trajectories within the north-star tolerance wherever both sides converged."""
    N, K, B, npts = 20, 16, 12, 10000
    h = A.Handle(N=N, K=K, max_batch=B, max_points=npts, max_edge_points=npts)
    h.cloud_set_layout(S.image_shape(npts)[0])
    clouds, edges, x0s, refs = [], [], [], []
    for s in range(B):
        c, e = S.forest_cloud(900 + s, npts)
        h.cloud_set(s, c)
        h.cloud_set(s, e, kind=A.capi.CLOUD_EDGE)
        x0, ref, _ = S.states(900 + s, N)
        clouds.append(c), edges.append(e), x0s.append(x0), refs.append(ref)
    x0s, refs = np.stack(x0s), np.stack(refs)
    for b in range(0, B, 4):  # PlanWapionts acts on these
        c = clouds[b]
        refs[b, 0, :3] = c[np.argmin(np.linalg.norm(c[:, :3] - refs[b, 0, :3], axis=1)), :3] + 0.05
    W0 = np.stack([S.warm_start("ref", x0s[b], refs[b], N) for b in range(B)])
    so = np.arange(B, dtype=np.int32)
    w, ref, info, rounds, safe = h.tick(x0s, refs, W0, scene_of=so, speed=D.SPEED,
                                        safety_distance=D.SAFETY_DISTANCE, max_rounds=3)
    ow, oref, orounds, osafe, oconv = _oracle_tick(clouds, edges, x0s, refs, W0, 3, D.SAFETY_DISTANCE, N, K)
    assert (safe == osafe).all()
    assert (rounds == orounds).all()
    ok = oconv & (info["status"] == 0)
    assert ok.sum() >= B - 2
    assert np.abs(w - ow)[ok].max() < 1e-4          # north-star tolerance
    assert np.abs(ref - oref)[ok].max() < 1e-4
    h.close()
