"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/liboracle.so)
and, when present, the reference's own KDTreeTwo/nanoflann (oracle/_ref/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(avoid-mpc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libampc_ref_kdtree.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.exists(_LIB) or (
        os.path.getmtime(_LIB) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("knn_oracle.c", "nlp_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/roswrapper/ros/src/avoid_mpc/include/kd_tree_two.h") and (
        force or not os.path.exists(_REF) or not os.path.exists(os.path.join(_HERE, "_ref", "libampc_cpu_arm.so"))
    ):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class Opts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int32), ("mu_init", C.c_double),
                ("bound_push", C.c_double), ("bound_frac", C.c_double),
                ("eps_min", C.c_double), ("eps_scale", C.c_double), ("tau_min", C.c_double),
                ("kappa_eps", C.c_double), ("kappa_mu", C.c_double), ("theta_mu", C.c_double),
                ("proj_step", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("cost", C.c_double), ("iters", C.c_int32), ("status", C.c_int32),
                ("kkt_dual", C.c_double), ("kkt_primal", C.c_double), ("kkt_compl", C.c_double),
                ("mu", C.c_double), ("reg_last", C.c_double), ("n_reg", C.c_int32),
                ("n_backtrack", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.knn_oracle_filter_nan.restype = C.c_int64
        L.knn_oracle_filter_nan.argtypes = [_vp, C.c_int64, C.c_int64, _vp]
        L.knn_oracle_bruteforce.restype = C.c_int
        L.knn_oracle_bruteforce.argtypes = [_vp, C.c_int64, C.c_int64, _dp, C.c_int, _ip, _dp]
        L.knn_oracle_is_tie_free.restype = C.c_int
        L.knn_oracle_is_tie_free.argtypes = [_vp, C.c_int64, C.c_int64, _dp, C.c_int]
        L.knn_oracle_tree_build.restype = _vp
        L.knn_oracle_tree_build.argtypes = [_vp, C.c_int64]
        L.knn_oracle_tree_free.argtypes = [_vp]
        L.knn_oracle_tree_search.restype = C.c_int
        L.knn_oracle_tree_search.argtypes = [_vp, _dp, C.c_int, _ip, _dp]
        L.knn_oracle_tree_search_batch.argtypes = [_vp, _dp, C.c_int, C.c_int, _ip, _dp, _ip]
        for name in ("nlp_oracle_nw", "nlp_oracle_ng"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = [C.c_int]
        L.nlp_oracle_np.restype = C.c_int
        L.nlp_oracle_np.argtypes = [C.c_int, C.c_int]
        L.nlp_oracle_F.argtypes = [_dp, _dp, _dp, C.c_double, _dp]
        L.nlp_oracle_dyn_matrices.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.nlp_oracle_g.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_double, _dp]
        L.nlp_oracle_f.restype = C.c_double
        L.nlp_oracle_f.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.nlp_oracle_grad_f.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.nlp_oracle_hess_f.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.nlp_oracle_default_opts.argtypes = [C.POINTER(Opts)]
        L.nlp_oracle_solve.restype = C.c_int
        L.nlp_oracle_solve.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp,
                                       C.POINTER(Opts), C.POINTER(Info), _dp]
        L.nlp_oracle_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp,
                                             C.POINTER(Opts), C.POINTER(Info)]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def as_xyz16(xyz) -> np.ndarray:
    """(n,3) or (n,4) float32 -> contiguous (n,4) float32 records (pcl::PointXYZ layout)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    if xyz.ndim != 2:
        raise ValueError("cloud must be 2-D")
    if xyz.shape[1] == 4:
        return np.ascontiguousarray(xyz)
    out = np.ones((xyz.shape[0], 4), dtype=np.float32)
    out[:, :3] = xyz
    return out


# ----------------------------------------------------------------- k-NN ----
def filter_nan(xyz16: np.ndarray) -> np.ndarray:
    xyz16 = as_xyz16(xyz16)
    out = np.empty_like(xyz16)
    m = lib().knn_oracle_filter_nan(xyz16.ctypes.data, xyz16.shape[0], 16, out.ctypes.data)
    return out[:m].copy()


def knn_bruteforce(xyz16: np.ndarray, queries: np.ndarray, k: int):
    """Canonical exact k-NN. Returns (idx[Q,k] int32 (-1 pad), dist2[Q,k] (inf pad), count[Q])."""
    xyz16 = as_xyz16(xyz16)
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 3)
    Q = q.shape[0]
    idx = np.full((Q, k), -1, dtype=np.int32)
    d2 = np.full((Q, k), np.inf, dtype=np.float64)
    cnt = np.zeros(Q, dtype=np.int32)
    L = lib()
    for i in range(Q):
        ii = np.empty(k, dtype=np.int32)
        dd = np.empty(k, dtype=np.float64)
        c = L.knn_oracle_bruteforce(xyz16.ctypes.data, xyz16.shape[0], 16, _d(q[i]), k, _i(ii), _d(dd))
        idx[i, :c] = ii[:c]
        d2[i, :c] = dd[:c]
        cnt[i] = c
    return idx, d2, cnt


def is_tie_free(xyz16: np.ndarray, queries: np.ndarray, k: int) -> bool:
    xyz16 = as_xyz16(xyz16)
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 3)
    L = lib()
    return all(L.knn_oracle_is_tie_free(xyz16.ctypes.data, xyz16.shape[0], 16, _d(q[i]), k) for i in range(q.shape[0]))


class PortTree:
    """kd-tree restatement (oracle/knn_oracle.c)."""

    def __init__(self, xyz16: np.ndarray):
        self.xyz = filter_nan(xyz16)
        self.h = lib().knn_oracle_tree_build(self.xyz.ctypes.data, self.xyz.shape[0])

    def search(self, queries, k):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 3)
        Q = q.shape[0]
        idx = np.full((Q, k), -1, dtype=np.int32)
        d2 = np.full((Q, k), np.inf, dtype=np.float64)
        cnt = np.zeros(Q, dtype=np.int32)
        lib().knn_oracle_tree_search_batch(self.h, _d(q), Q, k, _i(idx), _d(d2), _i(cnt))
        for i in range(Q):
            idx[i, cnt[i]:] = -1
            d2[i, cnt[i]:] = np.inf
        return idx, d2, cnt

    def __del__(self):
        if getattr(self, "h", None):
            lib().knn_oracle_tree_free(self.h)
            self.h = None


_ref = None


def ref_available() -> bool:
    build()
    return os.path.exists(_REF)


def ref_lib():
    global _ref
    if _ref is None:
        R = C.CDLL(_REF)
        R.ref_tree_create.restype = _vp
        R.ref_tree_create.argtypes = [_vp, C.c_int64]
        R.ref_tree_destroy.argtypes = [_vp]
        R.ref_tree_size.restype = C.c_int64
        R.ref_tree_size.argtypes = [_vp]
        R.ref_tree_search.restype = C.c_int
        R.ref_tree_search.argtypes = [_vp, C.c_double, C.c_double, C.c_double, C.c_int, _ip, _dp, C.POINTER(C.c_float)]
        R.ref_tree_search_batch.argtypes = [_vp, _dp, C.c_int, C.c_int, _ip, _dp, _ip]
        _ref = R
    return _ref


class RefTree:
    """The reference's own KDTreeTwo<double> (compiled from /root/reference)."""

    def __init__(self, xyz16: np.ndarray):
        xyz16 = as_xyz16(xyz16)
        self.h = ref_lib().ref_tree_create(xyz16.ctypes.data, xyz16.shape[0])

    def size(self):
        return ref_lib().ref_tree_size(self.h)

    def search(self, queries, k):
        q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 3)
        Q = q.shape[0]
        idx = np.full((Q, k), -1, dtype=np.int32)
        d2 = np.full((Q, k), np.inf, dtype=np.float64)
        cnt = np.zeros(Q, dtype=np.int32)
        ref_lib().ref_tree_search_batch(self.h, _d(q), Q, k, _i(idx), _d(d2), _i(cnt))
        for i in range(Q):
            idx[i, cnt[i]:] = -1
            d2[i, cnt[i]:] = np.inf
        return idx, d2, cnt

    def __del__(self):
        if getattr(self, "h", None):
            ref_lib().ref_tree_destroy(self.h)
            self.h = None


# ------------------------------------------------- native multi-threaded CPU arm ----
_ARM = os.path.join(_HERE, "_ref", "libampc_cpu_arm.so")
_arm = None


def cpu_arm_available() -> bool:
    build()
    return os.path.exists(_ARM)


def cpu_arm_run(n, threads, clouds16, x0, ref, tgt, tail34, W0, N, K, dt, lbu, ubu, opts: "Opts | None" = None):
    """n control rounds (reference kd-tree build + N x K-NN + packing + oracle NLP solve) on `threads`
    std::threads (oracle/cpu_arm.cpp).  clouds16: (n_distinct, npts, 4) f32; x0/ref/tgt/W0 per distinct
    scene.  Returns (wall seconds, status[n], iters[n], stage seconds [build, knn, solve] summed)."""
    global _arm
    if _arm is None:
        A = C.CDLL(_ARM)
        A.cpu_arm_run.restype = C.c_double
        A.cpu_arm_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, _vp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int,
                                  C.c_double, _dp, _dp, C.POINTER(Opts), _ip, _ip, _dp]
        _arm = A
    clouds16 = np.ascontiguousarray(clouds16, dtype=np.float32)
    nd, npts = clouds16.shape[0], clouds16.shape[1]
    x0, ref, tgt, W0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (x0, ref, tgt, W0))
    tail34 = np.ascontiguousarray(tail34, dtype=np.float64)
    assert tail34.size == 34 and clouds16.shape[2] == 4
    lbu = np.ascontiguousarray(lbu, dtype=np.float64)
    ubu = np.ascontiguousarray(ubu, dtype=np.float64)
    opts = opts or default_opts()
    st = np.zeros(n, dtype=np.int32)
    it = np.zeros(n, dtype=np.int32)
    stage = np.zeros(3)
    wall = _arm.cpu_arm_run(n, threads, nd, npts, clouds16.ctypes.data, _d(x0), _d(ref), _d(tgt), _d(tail34), _d(W0),
                            N, K, dt, _d(lbu), _d(ubu), C.byref(opts), _i(st), _i(it), _d(stage))
    return wall, st, it, stage


# ------------------------------------------------------------------ NLP ----
def nw(N):
    return lib().nlp_oracle_nw(N)


def ng(N):
    return lib().nlp_oracle_ng(N)


def np_(N, K):
    return lib().nlp_oracle_np(N, K)


def F(x, u, tau, dt):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    out = np.empty(10)
    lib().nlp_oracle_F(_d(x), _d(u), _d(tau), dt, _d(out))
    return out


def dyn_matrices(tau, dt):
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    Phi, Gam, gam = np.empty((10, 10)), np.empty((10, 4)), np.empty(10)
    lib().nlp_oracle_dyn_matrices(_d(tau), dt, _d(Phi), _d(Gam), _d(gam))
    return Phi, Gam, gam


def f(N, K, w, p):
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    assert w.size == nw(N) and p.size == np_(N, K)
    return lib().nlp_oracle_f(N, K, _d(w), _d(p))


def grad_f(N, K, w, p):
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    g = np.empty(nw(N))
    lib().nlp_oracle_grad_f(N, K, _d(w), _d(p), _d(g))
    return g


def g(N, K, w, p, dt):
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty(ng(N))
    lib().nlp_oracle_g(N, K, _d(w), _d(p), dt, _d(out))
    return out


def hess_f(N, K, w, p):
    """Returns (Hx[N,10,10] for X_1..X_N, Hu[4] diagonal)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    Hx = np.empty((N, 10, 10))
    Hu = np.empty(4)
    lib().nlp_oracle_hess_f(N, K, _d(w), _d(p), _d(Hx), _d(Hu))
    return Hx, Hu


def default_opts(**kw) -> Opts:
    o = Opts()
    lib().nlp_oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def solve(N, K, dt, p, w0, lbu, ubu, opts: Opts | None = None, want_lam=False):
    p = np.ascontiguousarray(p, dtype=np.float64)
    w = np.array(w0, dtype=np.float64).copy()
    lbu = np.ascontiguousarray(lbu, dtype=np.float64)
    ubu = np.ascontiguousarray(ubu, dtype=np.float64)
    assert w.size == nw(N) and p.size == np_(N, K)
    opts = opts or default_opts()
    info = Info()
    lam = np.empty(ng(N)) if want_lam else None
    lib().nlp_oracle_solve(N, K, dt, _d(p), _d(w), _d(lbu), _d(ubu), C.byref(opts), C.byref(info),
                           _d(lam) if want_lam else None)
    if want_lam:
        return w, info, lam
    return w, info


def solve_batch(N, K, dt, P, W0, lbu, ubu, opts: Opts | None = None):
    P = np.ascontiguousarray(P, dtype=np.float64)
    W = np.array(W0, dtype=np.float64).copy()
    B = P.shape[0]
    lbu = np.ascontiguousarray(lbu, dtype=np.float64)
    ubu = np.ascontiguousarray(ubu, dtype=np.float64)
    opts = opts or default_opts()
    infos = (Info * B)()
    lib().nlp_oracle_solve_batch(B, N, K, dt, _d(P), _d(W), _d(lbu), _d(ubu), C.byref(opts), infos)
    return W, infos
