"""Shared builders for the parity tests (seeded inputs of SURVEY.md §8d)."""
import numpy as np

import avoid_mpc_b200 as A
from oracle import oracle as O

D, S = A.defaults, A.synth


def make_instances(scene_ids, N, K, npts, cloud_fn=None):
    """Returns dict with clouds (list of [n,4] f32), x0 [B,10], ref [B,N,10], tgt [B,10],
    obstacle blocks from the ORACLE k-NN [B,N,K,3], prefixes [B,n_prefix], full params."""
    clouds, x0s, refs, tgts, obsts, prefixes, params = [], [], [], [], [], [], []
    for sid in scene_ids:
        c = cloud_fn(sid) if cloud_fn else S.forest_cloud(sid, npts)[0]
        x0, ref, tgt = S.states(sid, N, 1.0 / N if N != 20 else D.BENCH_DT)
        idx, d2, cnt = O.knn_bruteforce(c, ref[:, :3], K)
        ob = np.full((N, K, 3), 1e4)
        cf = O.filter_nan(c)
        for q in range(N):
            ob[q, :cnt[q]] = cf[idx[q, :cnt[q]], :3].astype(np.float64)
        pre = S.pack_prefix(x0, ref, ob, tgt)
        clouds.append(c), x0s.append(x0), refs.append(ref), tgts.append(tgt), obsts.append(ob)
        prefixes.append(pre), params.append(S.full_params(pre))
    return dict(clouds=clouds, x0=np.stack(x0s), ref=np.stack(refs), tgt=np.stack(tgts),
                obst=np.stack(obsts), prefix=np.stack(prefixes), params=np.stack(params))


def oracle_solve_batch(N, K, dt, params, W0, **opts):
    lb, ub = D.u_bounds()
    W, infos = O.solve_batch(N, K, dt, params, W0, lb, ub, O.default_opts(**opts))
    st = np.array([i.status for i in infos])
    it = np.array([i.iters for i in infos])
    cost = np.array([i.cost for i in infos])
    return W, st, it, cost


def golden2_groups():
    """tests/golden/solve_golden2.npz regrouped by shape: {(N, K): dict(ids, params, W0, w, w_cert,
    cert, stage, cost, sid)} with the parameter vectors rebuilt from the stored seeds exactly as
    tests/golden/make_solve_golden2.py built them (synthetic scene -> oracle k-NN -> packing)."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_golden2.npz"))
    groups = {}
    for i, (sid, N, K, npts, ref_warm, stage) in enumerate(G["meta"].tolist()):
        groups.setdefault((N, K, npts), []).append(i)
    out = {}
    for (N, K, npts), ids in groups.items():
        sids = [int(G["meta"][i, 0]) for i in ids]
        inst = make_instances(sids, N, K, npts)
        W0 = np.stack([S.warm_start("ref" if G["meta"][i, 4] else "cold", inst["x0"][j], inst["ref"][j], N)
                       for j, i in enumerate(ids)])
        nw = 10 + 14 * N
        out[(N, K)] = dict(ids=ids, sid=sids, params=inst["params"], prefix=inst["prefix"], W0=W0,
                           w=G["w"][ids, :nw], w_cert=G["w_cert"][ids, :nw], cert=G["cert"][ids],
                           stage=G["meta"][ids, 5], cost=G["cost"][ids])
    return out


def golden2_classify(W, g, tol=1e-4):
    """Per instance of a golden2 group: 'same' (within tol of the independent optimum),
    'certified' (another basin: within tol of a point an independent epigraph solve, started at
    this algorithm's optimum, confirmed as a KKT point), else 'open'."""
    out = []
    for j in range(len(g["ids"])):
        if g["stage"][j] > 0 and np.abs(W[j] - g["w"][j]).max() < tol:
            out.append("same")
        elif g["cert"][j] == 1 and np.abs(W[j] - g["w_cert"][j]).max() < tol:
            out.append("certified")
        else:
            out.append("open")
    return out


# solve_golden2.npz: the instances on which this repository's algorithm (oracle and CUDA alike) does
# NOT land within 1e-4 of an independently computed or independently certified optimum, by cause.
GOLDEN2_KINK = {1192, 1275}          # minimiser on the |v.n| kink: one control differs by 1.4e-4 / 1.5e-4 (cost by 1e-5)
GOLDEN2_OTHER_KKT = {1081, 1105, 1191, 1242, 1329}  # converged (KKT <= 1e-8 of the smoothed problem); the epigraph solve
#   started AT this point leaves it for another KKT point 5e-3 .. 0.95 away -- of HIGHER cost for four of them
#   (+6.2, +0.06, +0.005, +0.56) and 0.07 % lower for 1329; the independent cold solve ends in a third basin
GOLDEN2_POOR_LOCAL = {1353}  # ends at a stationary point of cost 1577.8 (oracle: stalls there at KKT 4e-8 until its
#   100-iteration cap; CUDA: converges there in 41 iterations, same point to 1e-13); the epigraph solve started AT it
#   escapes to the independent optimum, cost 17.2 -- the one instance of 512 where this algorithm is clearly worse


def golden2_check(N, K, g, W, status, cost, f_eval):
    """The assertions both golden2 tests make; returns the class counts."""
    cl = golden2_classify(W, g)
    counts = {"same": 0, "certified": 0, "kink": 0, "other_kkt": 0, "poor_local": 0}
    for j, c in enumerate(cl):
        sid = g["sid"][j]
        if c != "open":
            counts[c] += 1
            # (an OTHER_KKT instance may also end on its certificate point: the warp kernel does on 1105)
            assert sid not in GOLDEN2_KINK | GOLDEN2_POOR_LOCAL and (c == "certified" or sid not in GOLDEN2_OTHER_KKT), \
                ("stale allow-list entry", sid)
            if c == "same":
                assert abs(cost[j] - g["cost"][j]) <= 1e-6 * abs(g["cost"][j]), (sid, cost[j], g["cost"][j], status[j])
        elif sid in GOLDEN2_KINK:
            counts["kink"] += 1
            assert status[j] == 0 and np.abs(W[j] - g["w"][j]).max() < 2e-4, sid
            assert abs(cost[j] - g["cost"][j]) <= 1e-7 * abs(g["cost"][j]), sid
        elif sid in GOLDEN2_OTHER_KKT:
            counts["other_kkt"] += 1
            # (status 1 = the KKT residual stalls just above 1e-8 at the same point: 1191 on the warp kernel, 2e-8)
            assert status[j] in (0, 1) and g["cert"][j] == 1, sid
            assert cost[j] <= 1.001 * f_eval(N, K, g["w_cert"][j], g["params"][j]), sid
        else:
            assert sid in GOLDEN2_POOR_LOCAL, ("unexplained mismatch against solve_golden2.npz", sid, N, K)
            counts["poor_local"] += 1
            assert status[j] in (0, 1) and abs(cost[j] - 1577.792789) < 1e-5, sid
    return counts
