"""ctypes binding of libampc.so (include/ampc.h) — the same C-ABI the C++ shim
classes in avoid-mpc_b200/host/ call.  There is no fallback: if the library is
missing, or there is no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMPC_LIB: developer knob to A/B another build of the same library (e.g. lib/variants/*.so)
LIB_PATH = os.environ.get("AMPC_LIB") or os.path.join(_HERE, "lib", "libampc.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
SOLVE_CONVERGED, SOLVE_MAX_ITER, SOLVE_STALLED, SOLVE_NUMERIC = 0, 1, 2, 3
CLOUD_OBSTACLE, CLOUD_EDGE = 0, 1
LAYOUT_UNORGANISED, LAYOUT_SORT = 0, -1  # ampc_cloud_set_layout; a value >= 8 is an image row pitch

# every symbol include/ampc.h declares (tests check the library exports all of them)
SYMBOLS = [
    "ampc_api_version", "ampc_create", "ampc_destroy", "ampc_last_error",
    "ampc_set_weights", "ampc_set_tau", "ampc_set_gains", "ampc_set_radius", "ampc_set_accel_limits",
    "ampc_default_solver_opts", "ampc_set_solver_opts", "ampc_get_dynamics",
    "ampc_cloud_set", "ampc_cloud_set_batch", "ampc_cloud_set_batch_dev", "ampc_cloud_index_dev", "ampc_cloud_set_layout", "ampc_cloud_count",
    "ampc_knn_batch", "ampc_knn_batch_dev", "ampc_solve_batch", "ampc_solve_batch_dev",
    "ampc_round_batch", "ampc_round_batch_dev", "ampc_tick_batch", "ampc_tick_batch_dev", "ampc_last_prefix_dev",
    "ampc_best_of", "ampc_best_of_dev", "ampc_launch_count", "ampc_stream", "ampc_synchronize",
    "ampc_profile_enable", "ampc_profile_get",
    "ampc_cloud_get", "ampc_set_camera", "ampc_depth_set_batch", "ampc_depth_set_batch_dev",
    "ampc_guess_round_batch", "ampc_guess_round_batch_dev", "ampc_measure_fp64_peak",
]


class Config(C.Structure):
    _fields_ = [("N", C.c_int32), ("K", C.c_int32), ("dt", C.c_double), ("max_batch", C.c_int32),
                ("max_scenes", C.c_int32), ("max_points", C.c_int32), ("max_edge_points", C.c_int32),
                ("device", C.c_int32)]


class Camera(C.Structure):
    """perception block of config/mpc_parameters.yaml:58-66 (full-resolution intrinsics)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("resize_scale", C.c_double), ("pixel2meter", C.c_double), ("depth_min", C.c_double),
                ("depth_max", C.c_double)]


DEPTH_F32, DEPTH_U16 = 0, 1


class SolverOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int32), ("mu_init", C.c_double),
                ("bound_push", C.c_double), ("bound_frac", C.c_double), ("eps_min", C.c_double),
                ("eps_scale", C.c_double), ("kappa_eps", C.c_double)]


INFO_DTYPE = np.dtype([("cost", "f8"), ("kkt_dual", "f8"), ("kkt_compl", "f8"), ("mu", "f8"),
                       ("iters", "i4"), ("status", "i4"), ("n_reg", "i4"), ("n_backtrack", "i4")])
assert INFO_DTYPE.itemsize == 48


class AmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libampc error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile libampc.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building libampc.so failed")
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "host")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout, r.stderr)
    if r.returncode:
        raise RuntimeError("building the C++ host shim failed")
    return LIB_PATH


_lib = None
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AmpcError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH)
        L.ampc_last_error.restype = C.c_char_p
        L.ampc_last_error.argtypes = [_vp]
        L.ampc_create.argtypes = [C.POINTER(Config), C.POINTER(_vp)]
        L.ampc_destroy.argtypes = [_vp]
        L.ampc_destroy.restype = None
        for n in ("ampc_set_weights", "ampc_set_tau", "ampc_set_gains"):
            getattr(L, n).argtypes = [_vp, _vp]
        L.ampc_set_radius.argtypes = [_vp, C.c_double]
        L.ampc_set_accel_limits.argtypes = [_vp] + [C.c_double] * 4
        L.ampc_default_solver_opts.argtypes = [C.POINTER(SolverOpts)]
        L.ampc_default_solver_opts.restype = None
        L.ampc_set_solver_opts.argtypes = [_vp, C.POINTER(SolverOpts)]
        L.ampc_get_dynamics.argtypes = [_vp, _vp, _vp, _vp]
        L.ampc_cloud_set.argtypes = [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32]
        L.ampc_cloud_set_batch.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int64, C.c_int32]
        L.ampc_cloud_set_batch_dev.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.c_int64, _vp]
        L.ampc_cloud_count.argtypes = [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.ampc_cloud_get.argtypes = [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.POINTER(C.c_int32)]
        L.ampc_set_camera.argtypes = [_vp, C.POINTER(Camera)]
        L.ampc_depth_set_batch.argtypes = [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int64, C.c_int64, _vp, _vp]
        L.ampc_depth_set_batch_dev.argtypes = [_vp, C.c_int32, C.c_int32, _vp, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int64, C.c_int64, _vp, _vp, _vp]
        L.ampc_knn_batch.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]
        L.ampc_knn_batch_dev.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp]
        L.ampc_solve_batch.argtypes = [_vp, C.c_int32, _vp, _vp, _vp]
        L.ampc_solve_batch_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp]
        L.ampc_round_batch.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp]
        L.ampc_round_batch_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp]
        L.ampc_last_prefix_dev.argtypes = [_vp, C.POINTER(_vp)]
        L.ampc_tick_batch.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int32,
                                      _vp, _vp, _vp, _vp]
        L.ampc_tick_batch_dev.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, _vp, C.c_double, C.c_double, C.c_int32,
                                          _vp, _vp, _vp, _vp, _vp]
        L.ampc_best_of.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp]
        L.ampc_best_of_dev.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]
        L.ampc_guess_round_batch.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp]
        L.ampc_guess_round_batch_dev.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp,
                                                 _vp, _vp]
        L.ampc_measure_fp64_peak.argtypes = [_vp, C.POINTER(C.c_double)]
        L.ampc_launch_count.restype = C.c_int64
        L.ampc_launch_count.argtypes = [_vp]
        L.ampc_stream.restype = _vp
        L.ampc_stream.argtypes = [_vp]
        L.ampc_synchronize.argtypes = [_vp]
        L.ampc_profile_enable.argtypes = [_vp, C.c_int]
        L.ampc_profile_get.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.ampc_cloud_index_dev.argtypes = [_vp, C.c_int32, C.c_int32, C.c_int32, _vp]
        L.ampc_cloud_set_layout.argtypes = [_vp, C.c_int32, C.c_int32]
        _lib = L
    return _lib


def _ptr(a):
    """numpy array / torch tensor / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if isinstance(a, int):
        return a
    return a.data_ptr()  # torch tensor


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def default_solver_opts(**kw) -> SolverOpts:
    o = SolverOpts()
    lib().ampc_default_solver_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Handle:
    """One solver/map handle on one CUDA device (wraps ampc_handle*)."""

    def __init__(self, N=20, K=16, dt=0.05, max_batch=1024, max_scenes=None, max_points=50000,
                 max_edge_points=0, device=0, defaults=True):
        self.L = lib()
        self.N, self.K, self.dt = N, K, dt
        self.n_w = 10 + 14 * N
        self.n_prefix = 20 + 10 * N + 3 * K * N
        cfg = Config(N, K, dt, max_batch, max_batch if max_scenes is None else max_scenes, max_points,
                     max_edge_points, device)
        self.cfg = cfg
        h = _vp()
        rc = self.L.ampc_create(C.byref(cfg), C.byref(h))
        if rc:
            raise AmpcError(rc, self.L.ampc_last_error(None).decode())
        self.h = h
        if defaults:
            self._shipped_parameters()

    @classmethod
    def borrowed(cls, ptr, N, K, dt, defaults=True):
        """View of an ampc_handle* owned by someone else (e.g. a device of ampc_multi): never destroyed here."""
        self = cls.__new__(cls)
        self.L = lib()
        self.N, self.K, self.dt = N, K, dt
        self.n_w = 10 + 14 * N
        self.n_prefix = 20 + 10 * N + 3 * K * N
        self.h = _vp(ptr)
        self._borrowed = True
        if defaults:
            self._shipped_parameters()
        return self

    def _shipped_parameters(self):  # the shipped mpc_parameters.yaml values
        from . import defaults as D
        self.set_weights(D.WEIGHTS)
        self.set_tau(D.TAU)
        self.set_gains(D.GAINS)
        self.set_radius(D.DRONE_RADIUS)
        self.set_accel_limits(D.A_MIN_Z, D.A_MAX_Z, D.A_MAX_XY, D.A_MAX_YAW_DOT)

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.L.ampc_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc:
            raise AmpcError(rc, self.L.ampc_last_error(self.h).decode())

    # parameters
    def set_weights(self, w):
        w = _f64(w, (25,))
        self._ck(self.L.ampc_set_weights(self.h, w.ctypes.data))

    def set_tau(self, t):
        t = _f64(t, (4,))
        self._ck(self.L.ampc_set_tau(self.h, t.ctypes.data))

    def set_gains(self, g):
        g = _f64(g, (4,))
        self._ck(self.L.ampc_set_gains(self.h, g.ctypes.data))

    def set_radius(self, r):
        self._ck(self.L.ampc_set_radius(self.h, float(r)))

    def set_accel_limits(self, a_min_z, a_max_z, a_max_xy, a_max_yaw_dot):
        self._ck(self.L.ampc_set_accel_limits(self.h, a_min_z, a_max_z, a_max_xy, a_max_yaw_dot))

    def set_solver_opts(self, **kw):
        o = default_solver_opts(**kw)
        self._ck(self.L.ampc_set_solver_opts(self.h, C.byref(o)))

    def dynamics(self):
        Phi, Gam, gam = np.empty((10, 10)), np.empty((10, 4)), np.empty(10)
        self._ck(self.L.ampc_get_dynamics(self.h, Phi.ctypes.data, Gam.ctypes.data, gam.ctypes.data))
        return Phi, Gam, gam

    # clouds
    def cloud_set(self, scene, xyz, kind=CLOUD_OBSTACLE):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        n = xyz.shape[0]
        stride = xyz.strides[0] if n > 0 else 16
        self._ck(self.L.ampc_cloud_set(self.h, scene, kind, xyz.ctypes.data if n else None, n, stride))

    def cloud_set_batch(self, xyz, counts=None, first_scene=0, kind=CLOUD_OBSTACLE):
        """xyz: (S, P, 3|4) float32 host array (numpy, or a pinned torch tensor)."""
        S, Pn, Cn = xyz.shape
        cnt = np.full(S, Pn, dtype=np.int32) if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
        self._ck(self.L.ampc_cloud_set_batch(self.h, kind, first_scene, S, _ptr(xyz), cnt.ctypes.data,
                                             Pn * Cn * 4, Cn * 4))

    def cloud_set_batch_dev(self, xyz_dev, counts=None, first_scene=0, kind=CLOUD_OBSTACLE, stream=None):
        """xyz_dev: (S, P, 4) float32 CUDA tensor."""
        S, Pn, Cn = xyz_dev.shape
        assert Cn == 4
        cnt = np.full(S, Pn, dtype=np.int32) if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
        self._ck(self.L.ampc_cloud_set_batch_dev(self.h, kind, first_scene, S, _ptr(xyz_dev), cnt.ctypes.data,
                                                 Pn * 16, stream))

    def cloud_set_layout(self, row_width, kind=CLOUD_OBSTACLE):
        """row_width >= 8: organised cloud (image row pitch); LAYOUT_UNORGANISED; LAYOUT_SORT: arbitrary
        storage order, index built over a Morton-bucketed copy."""
        self._ck(self.L.ampc_cloud_set_layout(self.h, kind, row_width))

    def cloud_count(self, scene, kind=CLOUD_OBSTACLE):
        n = C.c_int32()
        self._ck(self.L.ampc_cloud_count(self.h, scene, kind, C.byref(n)))
        return n.value

    def cloud_get(self, scene, kind=CLOUD_OBSTACLE):
        """The cloud held for (scene, kind) as (n, 4) float32 records, in index order."""
        n = self.cloud_count(scene, kind)
        out = np.empty((n, 4), dtype=np.float32)
        m = C.c_int32()
        self._ck(self.L.ampc_cloud_get(self.h, scene, kind, _ptr(out) if n else None, n, C.byref(m)))
        return out

    # depth image -> Obstacle + Edge cloud (FrameKDMap::ProcessDepth)
    def set_camera(self, fx=320.0, fy=320.0, cx=320.0, cy=240.0, resize_scale=10.0, pixel2meter=1.0,
                   depth_min=0.1, depth_max=100.0):
        cam = Camera(fx, fy, cx, cy, resize_scale, pixel2meter, depth_min, depth_max)
        self._ck(self.L.ampc_set_camera(self.h, C.byref(cam)))

    def depth_set_batch(self, depth, T_obstacle, T_edge=None, first_scene=0):
        """depth: (S, rows, cols) or (rows, cols), float32 or uint16 (host)."""
        d = np.ascontiguousarray(depth)
        if d.ndim == 2:
            d = d[None]
        if d.dtype not in (np.float32, np.uint16):
            raise TypeError("depth images are float32 (CV_32FC1) or uint16 (CV_16UC1)")
        S, rows, cols = d.shape
        To = _f64(T_obstacle).reshape(S, 16)
        Te = None if T_edge is None else _f64(T_edge).reshape(S, 16)
        self._ck(self.L.ampc_depth_set_batch(self.h, first_scene, S, d.ctypes.data,
                                             DEPTH_U16 if d.dtype == np.uint16 else DEPTH_F32, rows, cols,
                                             cols * d.itemsize, rows * cols * d.itemsize, To.ctypes.data, _ptr(Te)))

    def depth_set_batch_dev(self, depth_dev, T_obstacle_dev, T_edge_dev=None, first_scene=0, stream=None):
        """depth_dev: contiguous (S, rows, cols) torch tensor (float32 or int16/uint16 storage) on the
        handle's device; transforms (S, 16) float64 on the device."""
        S, rows, cols = depth_dev.shape
        esz = depth_dev.element_size()
        self._ck(self.L.ampc_depth_set_batch_dev(self.h, first_scene, S, _ptr(depth_dev),
                                                 DEPTH_U16 if esz == 2 else DEPTH_F32, rows, cols, cols * esz,
                                                 rows * cols * esz, _ptr(T_obstacle_dev), _ptr(T_edge_dev), stream))

    # k-NN (host buffers)
    def knn(self, queries, k, scene_of=None, kind=CLOUD_OBSTACLE, want_pts=True):
        q = _f64(queries)
        B, Q = q.shape[0], q.shape[1]
        so = None if scene_of is None else np.ascontiguousarray(scene_of, dtype=np.int32)
        idx = np.empty((B, Q, k), dtype=np.int32)
        d2 = np.empty((B, Q, k), dtype=np.float64)
        pts = np.empty((B, Q, k, 3), dtype=np.float64) if want_pts else None
        cnt = np.empty((B, Q), dtype=np.int32)
        self._ck(self.L.ampc_knn_batch(self.h, kind, B, _ptr(so), q.ctypes.data, Q, k, idx.ctypes.data,
                                       d2.ctypes.data, _ptr(pts), cnt.ctypes.data))
        return idx, d2, pts, cnt

    def knn_dev(self, queries_dev, k, idx_dev, d2_dev, pts_dev, cnt_dev, scene_of_dev=None,
                kind=CLOUD_OBSTACLE, stream=None):
        B, Q = queries_dev.shape[0], queries_dev.shape[1]
        self._ck(self.L.ampc_knn_batch_dev(self.h, kind, B, _ptr(scene_of_dev), _ptr(queries_dev), Q, k,
                                           _ptr(idx_dev), _ptr(d2_dev), _ptr(pts_dev), _ptr(cnt_dev), stream))

    # solve
    def solve(self, prefix, w0):
        p = _f64(prefix).reshape(-1, self.n_prefix)
        B = p.shape[0]
        w = np.array(w0, dtype=np.float64).reshape(B, self.n_w).copy()
        info = np.zeros(B, dtype=INFO_DTYPE)
        self._ck(self.L.ampc_solve_batch(self.h, B, p.ctypes.data, w.ctypes.data, info.ctypes.data))
        return w, info

    def solve_dev(self, B, prefix_dev, w_dev, info_dev=None, stream=None):
        self._ck(self.L.ampc_solve_batch_dev(self.h, B, _ptr(prefix_dev), _ptr(w_dev), _ptr(info_dev), stream))

    # one round: k-NN + pack + solve
    def round(self, x0, ref, w0, scene_of=None, pos_x=None, speed=10.0, safety_distance=0.2):
        x0 = _f64(x0).reshape(-1, 10)
        B = x0.shape[0]
        ref = _f64(ref).reshape(B, self.N, 10)
        w = np.array(w0, dtype=np.float64).reshape(B, self.n_w).copy()
        so = None if scene_of is None else np.ascontiguousarray(scene_of, dtype=np.int32)
        px = None if pos_x is None else _f64(pos_x, (B,))
        info = np.zeros(B, dtype=INFO_DTYPE)
        replan = np.zeros(B, dtype=np.int32)
        self._ck(self.L.ampc_round_batch(self.h, B, _ptr(so), x0.ctypes.data, ref.ctypes.data, _ptr(px),
                                         speed, safety_distance, w.ctypes.data, info.ctypes.data,
                                         replan.ctypes.data))
        return w, info, replan

    def round_host_ptrs(self, B, x0, ref, w, info, replan, scene_of=None, pos_x=None, speed=10.0,
                        safety_distance=0.2):
        """Same call on caller-owned (e.g. pinned) host buffers; no allocation."""
        self._ck(self.L.ampc_round_batch(self.h, B, _ptr(scene_of), _ptr(x0), _ptr(ref), _ptr(pos_x),
                                         speed, safety_distance, _ptr(w), _ptr(info), _ptr(replan)))

    def round_dev(self, B, x0_dev, ref_dev, w_dev, info_dev=None, replan_dev=None, scene_of_dev=None,
                  pos_x_dev=None, speed=10.0, safety_distance=0.2, stream=None):
        self._ck(self.L.ampc_round_batch_dev(self.h, B, _ptr(scene_of_dev), _ptr(x0_dev), _ptr(ref_dev),
                                             _ptr(pos_x_dev), speed, safety_distance, _ptr(w_dev),
                                             _ptr(info_dev), _ptr(replan_dev), stream))

    def tick(self, x0, ref, w0, scene_of=None, pos_x=None, speed=10.0, safety_distance=0.2, max_rounds=3):
        """One control tick (<= max_rounds rounds on the device). Returns (w, ref, info, rounds, is_safety)."""
        x0 = _f64(x0).reshape(-1, 10)
        B = x0.shape[0]
        ref = np.array(ref, dtype=np.float64).reshape(B, self.N, 10).copy()
        w = np.array(w0, dtype=np.float64).reshape(B, self.n_w).copy()
        so = None if scene_of is None else np.ascontiguousarray(scene_of, dtype=np.int32)
        px = None if pos_x is None else _f64(pos_x, (B,))
        info = np.zeros(B, dtype=INFO_DTYPE)
        rounds = np.zeros(B, dtype=np.int32)
        safe = np.zeros(B, dtype=np.int32)
        self._ck(self.L.ampc_tick_batch(self.h, B, _ptr(so), x0.ctypes.data, ref.ctypes.data, _ptr(px), speed,
                                        safety_distance, max_rounds, w.ctypes.data, info.ctypes.data,
                                        rounds.ctypes.data, safe.ctypes.data))
        return w, ref, info, rounds, safe

    def tick_dev(self, B, x0_dev, ref_dev, w_dev, info_dev=None, rounds_dev=None, safe_dev=None, scene_of_dev=None,
                 pos_x_dev=None, speed=10.0, safety_distance=0.2, max_rounds=3, stream=None):
        self._ck(self.L.ampc_tick_batch_dev(self.h, B, _ptr(scene_of_dev), _ptr(x0_dev), _ptr(ref_dev),
                                            _ptr(pos_x_dev), speed, safety_distance, max_rounds, _ptr(w_dev),
                                            _ptr(info_dev), _ptr(rounds_dev), _ptr(safe_dev), stream))

    def last_prefix_ptr(self):
        p = _vp()
        self._ck(self.L.ampc_last_prefix_dev(self.h, C.byref(p)))
        return p.value

    def best_of(self, info, n_scenes, G):
        info = np.ascontiguousarray(info, dtype=INFO_DTYPE)
        arg = np.empty(n_scenes, dtype=np.int32)
        best = np.empty(n_scenes, dtype=np.float64)
        self._ck(self.L.ampc_best_of(self.h, n_scenes, G, info.ctypes.data, arg.ctypes.data, best.ctypes.data))
        return arg, best

    def guess_round(self, x0, ref, w0, G, pos_x=None, speed=10.0):
        """Edge-tree guesses + best-of-G (BASELINE config C2): x0 (S,10), ref (S,N,10), w0 (S*G, n_w).
        Returns (w, info, argmin, best_cost)."""
        x0 = _f64(x0).reshape(-1, 10)
        Sn = x0.shape[0]
        ref = _f64(ref, (Sn, self.N, 10))
        w = np.array(w0, dtype=np.float64).reshape(Sn * G, self.n_w).copy()
        px = None if pos_x is None else _f64(pos_x, (Sn,))
        info = np.zeros(Sn * G, dtype=INFO_DTYPE)
        arg = np.empty(Sn, dtype=np.int32)
        best = np.empty(Sn, dtype=np.float64)
        self._ck(self.L.ampc_guess_round_batch(self.h, Sn, G, x0.ctypes.data, ref.ctypes.data, _ptr(px), speed,
                                               w.ctypes.data, info.ctypes.data, arg.ctypes.data, best.ctypes.data))
        return w, info, arg, best

    def guess_round_dev(self, n_scenes, G, x0_dev, ref_dev, w_dev, info_dev=None, argmin_dev=None, best_dev=None,
                        pos_x_dev=None, speed=10.0, stream=None):
        self._ck(self.L.ampc_guess_round_batch_dev(self.h, n_scenes, G, _ptr(x0_dev), _ptr(ref_dev), _ptr(pos_x_dev),
                                                   speed, _ptr(w_dev), _ptr(info_dev), _ptr(argmin_dev),
                                                   _ptr(best_dev), stream))

    def measure_fp64_peak(self):
        """FP64 FMA throughput of the device in TFLOP/s (hand-written dependent-FMA-chain kernel)."""
        v = C.c_double()
        self._ck(self.L.ampc_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def best_of_dev(self, n_scenes, G, info_dev, argmin_dev, best_dev, stream=None):
        self._ck(self.L.ampc_best_of_dev(self.h, n_scenes, G, _ptr(info_dev), _ptr(argmin_dev), _ptr(best_dev), stream))

    def profile_enable(self, on=True):
        self._ck(self.L.ampc_profile_enable(self.h, 1 if on else 0))

    def profile_get(self):
        """{'index'|'knn'|'solve': (ms total, launches)}"""
        ms, n = (C.c_double * 3)(), (C.c_int64 * 3)()
        self._ck(self.L.ampc_profile_get(self.h, ms, n))
        return {name: (ms[i], n[i]) for i, name in enumerate(("index", "knn", "solve"))}

    def cloud_index_dev(self, first_scene, n_scenes, kind=CLOUD_OBSTACLE, stream=None):
        self._ck(self.L.ampc_cloud_index_dev(self.h, kind, first_scene, n_scenes, stream))

    def launch_count(self):
        return self.L.ampc_launch_count(self.h)

    def stream(self):
        return self.L.ampc_stream(self.h)

    def synchronize(self):
        self._ck(self.L.ampc_synchronize(self.h))
