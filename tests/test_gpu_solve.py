"""GPU parity: the batched interior-point solve through the C-ABI vs the CPU oracle."""
import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import oracle as O
from helpers import make_instances, oracle_solve_batch

pytestmark = pytest.mark.gpu
D, S = A.defaults, A.synth

TRAJ_TOL = 1e-4   # north-star tolerance on state/control trajectories (l_inf)
TIGHT_TOL = 1e-6  # what the two implementations actually agree to at convergence


@pytest.fixture(params=["warp", "quad"], autouse=True)
def solve_kernel(request, monkeypatch):
    """Every test of this file runs on both solve kernels (the library picks by batch size:
    one warp per instance below 8192 instances, four lanes per instance above)."""
    monkeypatch.setenv("AMPC_SOLVE_KERNEL", request.param)
    return request.param


def test_dynamics_match_oracle():
    h = A.Handle(N=20, K=16, max_batch=1, max_points=16)
    Phi, Gam, gam = h.dynamics()
    oP, oG, og = O.dyn_matrices(D.TAU, 0.05)
    assert np.abs(Phi - oP).max() < 1e-15 and np.abs(Gam - oG).max() < 1e-15 and np.abs(gam - og).max() < 1e-15
    h.close()


@pytest.mark.parametrize("N,K,npts,kind", [(20, 16, 10000, "cold"), (20, 16, 10000, "ref"),
                                           (20, 8, 10000, "ref"), (30, 3, 3072, "cold")])
def test_solve_matches_oracle(N, K, npts, kind):
    B = 48
    dt = 1.0 / N if N != 20 else 0.05
    inst = make_instances(range(200, 200 + B), N, K, npts)
    W0 = np.stack([S.warm_start(kind, inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    h = A.Handle(N=N, K=K, dt=dt, max_batch=B, max_points=16)
    W, info = h.solve(inst["prefix"], W0)
    oW, ost, oit, ocost = oracle_solve_batch(N, K, dt, inst["params"], W0)
    # Same algorithm on both sides: EVERY instance must end with the same status at the same point.
    # Allow-list of known divergences (scene id -> reason): empty -- tools/enumerate_mismatch.py
    # found none on these 4 x 48 instances with either kernel (worst l_inf 1e-13); an entry here
    # needs the two iterate histories that explain it.  Iteration counts may differ by rounding
    # (sums are associated differently on the two sides, and an Armijo or inertia test can fall
    # the other way): the same count is required on 95 % of the instances, within 2 on all.
    ALLOW = {}
    bad = [200 + b for b in range(B) if (200 + b) not in ALLOW and
           (info["status"][b] != ost[b] or np.abs(W[b] - oW[b]).max() >= TIGHT_TOL)]
    assert not bad, f"GPU and oracle disagree on scenes {bad}"
    dit = np.abs(info["iters"].astype(int) - oit)
    assert (dit == 0).mean() >= 0.95 and dit[ost == 0].max() <= 2, dit
    both = (info["status"] == 0) & (ost == 0)
    assert both.mean() >= 0.9, f"converged on both sides: {both.mean():.2f}"
    err = np.abs(W - oW).max(axis=1)
    assert err[both].max() < TIGHT_TOL
    rel = np.abs(info["cost"][both] - ocost[both]) / np.maximum(1.0, np.abs(ocost[both]))
    assert rel.max() < 1e-9
    # KKT residuals reported by the kernel
    assert (info["kkt_dual"][info["status"] == 0] <= 1e-8).all()
    assert (info["kkt_compl"][info["status"] == 0] <= 1e-8).all()
    h.close()


def test_solution_is_feasible_and_outputs_slice_like_reference():
    N, K, B = 20, 16, 16
    inst = make_instances(range(300, 300 + B), N, K, 10000)
    W0 = np.stack([S.warm_start("ref", inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    h = A.Handle(N=N, K=K, max_batch=B, max_points=16)
    W, info = h.solve(inst["prefix"], W0)
    lb, ub = D.u_bounds()
    for b in range(B):
        g = O.g(N, K, W[b], inst["params"][b], 0.05)
        assert np.abs(g).max() < 1e-9            # X_0 = x0 and the dynamics hold
        U = np.stack([W[b][14 * k + 10:14 * k + 14] for k in range(N)])
        assert (U >= lb - 1e-12).all() and (U <= ub + 1e-12).all()
        assert abs(O.f(N, K, W[b], inst["params"][b]) - info["cost"][b]) <= 1e-9 * max(1, abs(info["cost"][b]))
    h.close()


def test_status_reports_iteration_cap():
    N, K, B = 20, 16, 8
    inst = make_instances(range(400, 400 + B), N, K, 10000)
    W0 = np.zeros((B, 10 + 14 * N))
    h = A.Handle(N=N, K=K, max_batch=B, max_points=16)
    h.set_solver_opts(max_iter=2)
    W, info = h.solve(inst["prefix"], W0)
    assert (info["status"] == A.capi.SOLVE_MAX_ITER).all() and (info["iters"] == 2).all()
    oW, ost, oit, _ = oracle_solve_batch(N, K, 0.05, inst["params"], W0, max_iter=2)
    assert np.abs(W - oW).max() < 1e-8  # same two iterates
    h.close()


def test_solve_matches_full_space_interior_point_golden():
    """The CUDA solve against optima computed by an independent interior-point solver (scipy
    trust-constr) on the reference's full-space formulation (tests/golden/make_solve_golden.py):
    same point wherever that solver converged, l_inf < 1e-4 on states and controls."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "solve_golden.npz"))
    groups = {}
    for i in range(int(G["n"])):
        sid, N, K, npts, ref_warm = (int(v) for v in G[f"i{i}_meta"])
        groups.setdefault((N, K), []).append(i)
    checked = 0
    for (N, K), ids in groups.items():
        dt = 0.05 if N == 20 else 1.0 / N
        n_prefix = 20 + 10 * N + 3 * K * N
        P = np.stack([G[f"i{i}_p"][:n_prefix] for i in ids])
        W0 = np.stack([G[f"i{i}_w0"] for i in ids])
        h = A.Handle(N=N, K=K, dt=dt, max_batch=len(ids), max_points=16)
        W, info = h.solve(P, W0)
        h.close()
        assert (info["status"] == A.capi.SOLVE_CONVERGED).all()
        for j, i in enumerate(ids):
            if not bool(G[f"i{i}_conv"]):
                continue
            checked += 1
            assert np.abs(W[j] - G[f"i{i}_w"]).max() < TRAJ_TOL, (N, K, i)
            assert abs(info["cost"][j] - float(G[f"i{i}_cost"])) <= 1e-7 * abs(float(G[f"i{i}_cost"]))
    assert checked >= 49


def test_solve_matches_512_independent_optima():
    """The CUDA solve against tests/golden/solve_golden2.npz (512 optima from independent solvers,
    see tests/test_oracle_nlp.py::test_converged_optimum_matches_512_independent_optima): the same
    bar, the same eight allow-listed instances by cause, and the same class counts as the oracle."""
    from helpers import golden2_check, golden2_groups
    total = {}
    for (N, K), g in golden2_groups().items():
        dt = 0.05 if N == 20 else 1.0 / N
        n = len(g["ids"])
        h = A.Handle(N=N, K=K, dt=dt, max_batch=n, max_points=16)
        h.set_solver_opts(max_iter=100)  # the oracle's default cap, so the iteration-cap class is the same
        W, info = h.solve(g["prefix"], g["W0"])
        h.close()
        for k, v in golden2_check(N, K, g, W, info["status"], info["cost"], O.f).items():
            total[k] = total.get(k, 0) + v
    assert sum(total.values()) == 512
    assert (total["same"], total["certified"] + total["other_kkt"], total["kink"], total["poor_local"]) == (467, 42, 2, 1), total


def test_quad_kernel_refill_and_packing_do_not_change_results(monkeypatch):
    """The quad kernel's results must not depend on how instances share warps: 1, 2, 4 or 8
    instances per warp, with and without refill from the queue (one resident warp per SM and
    2 instances per warp -> 296 slots for 400 instances), one or two warps per CTA, bit for
    bit; and they must agree with the warp-per-instance kernel to rounding."""
    N, K, B = 20, 16, 400
    inst = make_instances([500 + (b % 40) for b in range(B)], N, K, 10000)
    rng = np.random.default_rng(7)
    W0 = np.stack([S.warm_start("ref", inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    W0[:, :] += 0.0
    for b in range(B):  # distinct problems from 40 scenes: perturb the warm-start controls
        for k in range(N):
            W0[b, 14 * k + 10:14 * k + 14] += rng.normal(0, 0.3, 4) * (b // 40)
    results = {}
    for name, env in [("warp", {"AMPC_SOLVE_KERNEL": "warp"}),
                      ("q8", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "8"}),
                      ("q4", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "4"}),
                      ("q1", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "1"}),
                      ("q2_refill", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "2",
                                     "AMPC_QUAD_WARPS_PER_SM": "1"}),
                      # two warps per CTA meeting once per pass (the default for full machines), with refill
                      ("q1_cta2", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "1",
                                   "AMPC_QUAD_WARPS_PER_SM": "2", "AMPC_QUAD_CTA_WARPS": "2"}),
                      ("q1_cta1", {"AMPC_SOLVE_KERNEL": "quad", "AMPC_QUADS_PER_WARP": "1",
                                   "AMPC_QUAD_WARPS_PER_SM": "2", "AMPC_QUAD_CTA_WARPS": "1"})]:
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        h = A.Handle(N=N, K=K, max_batch=B, max_points=16)
        results[name] = h.solve(inst["prefix"], W0)
        h.close()
        monkeypatch.delenv("AMPC_QUADS_PER_WARP", raising=False)
        monkeypatch.delenv("AMPC_QUAD_WARPS_PER_SM", raising=False)
        monkeypatch.delenv("AMPC_QUAD_CTA_WARPS", raising=False)
    W8, i8 = results["q8"]
    for name in ("q4", "q1", "q2_refill", "q1_cta2", "q1_cta1"):
        W, info = results[name]
        assert (W == W8).all(), name
        for f in ("cost", "iters", "status", "n_reg", "n_backtrack"):
            assert (info[f] == i8[f]).all(), (name, f)
    Ww, iw = results["warp"]
    both = (iw["status"] == 0) & (i8["status"] == 0)
    assert both.mean() > 0.9
    assert (np.abs(Ww - W8).max(axis=1)[both] < TRAJ_TOL).mean() >= 0.98
    assert np.median(np.abs(Ww - W8).max(axis=1)[both]) < TIGHT_TOL
