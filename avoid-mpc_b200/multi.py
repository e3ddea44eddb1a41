"""ctypes binding of include/ampc_multi.h (libampc_multi.so): scene-sharded rounds over several
devices of ONE process, costs all-gathered with NCCL.  Used by the tests; a C++ host links the
library directly."""
import ctypes as C
import os

import numpy as np

from . import capi

LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libampc_multi.so")
SYMBOLS = ["ampc_multi_create", "ampc_multi_destroy", "ampc_multi_last_error", "ampc_multi_device_count",
           "ampc_multi_handle", "ampc_multi_shard", "ampc_multi_cloud_set_batch", "ampc_multi_cloud_set_layout",
           "ampc_multi_round_batch", "ampc_multi_costs_dev", "ampc_multi_block"]
_lib = None
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise capi.AmpcError(-1, f"{LIB_PATH} is missing: run __graft_entry__.build()")
        capi.lib()  # libampc.so first (same directory, $ORIGIN rpath)
        L = C.CDLL(LIB_PATH)
        L.ampc_multi_create.argtypes = [C.POINTER(capi.Config), _vp, C.c_int32, C.POINTER(_vp)]
        L.ampc_multi_destroy.argtypes = [_vp]
        L.ampc_multi_destroy.restype = None
        L.ampc_multi_last_error.argtypes = [_vp]
        L.ampc_multi_last_error.restype = C.c_char_p
        L.ampc_multi_device_count.argtypes = [_vp]
        L.ampc_multi_handle.argtypes = [_vp, C.c_int32]
        L.ampc_multi_handle.restype = _vp
        L.ampc_multi_shard.argtypes = [_vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.ampc_multi_shard.restype = None
        L.ampc_multi_cloud_set_batch.argtypes = [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_int64, C.c_int32]
        L.ampc_multi_cloud_set_layout.argtypes = [_vp, C.c_int32, C.c_int32]
        L.ampc_multi_round_batch.argtypes = [_vp, C.c_int32, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp]
        L.ampc_multi_costs_dev.argtypes = [_vp, C.c_int32]
        L.ampc_multi_costs_dev.restype = _vp
        L.ampc_multi_block.argtypes = [_vp, C.c_int32]
        _lib = L
    return _lib


class MultiHandle:
    def __init__(self, devices, N=20, K=16, dt=0.05, max_batch=1024, max_points=50000, max_edge_points=0):
        self.L = lib()
        self.N, self.K = N, K
        cfg = capi.Config(N, K, dt, max_batch, max_batch, max_points, max_edge_points, 0)
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        self.m = _vp()
        rc = self.L.ampc_multi_create(C.byref(cfg), dv.ctypes.data, len(dv), C.byref(self.m))
        if rc:
            raise capi.AmpcError(rc, (self.L.ampc_multi_last_error(None) or b"").decode())
        self.n = len(dv)
        # per-device views (parameters are per handle: the shipped mpc_parameters.yaml values, as capi.Handle)
        self.handles = [capi.Handle.borrowed(self.L.ampc_multi_handle(self.m, i), N, K, dt) for i in range(self.n)]

    def _ck(self, rc):
        if rc:
            raise capi.AmpcError(rc, (self.L.ampc_multi_last_error(self.m) or b"").decode())

    def shard(self, batch, i):
        f, c = C.c_int32(), C.c_int32()
        self.L.ampc_multi_shard(self.m, batch, i, C.byref(f), C.byref(c))
        return f.value, c.value

    def cloud_set_layout(self, row_width, kind=capi.CLOUD_OBSTACLE):
        self._ck(self.L.ampc_multi_cloud_set_layout(self.m, kind, row_width))

    def cloud_set_batch(self, clouds, kind=capi.CLOUD_OBSTACLE):
        """clouds: (S, n, 4) float32 (16-byte records)."""
        c = np.ascontiguousarray(clouds, dtype=np.float32)
        S, n = c.shape[0], c.shape[1]
        counts = np.full(S, n, dtype=np.int32)
        self._ck(self.L.ampc_multi_cloud_set_batch(self.m, kind, S, c.ctypes.data, counts.ctypes.data, n * 16, 16))

    def round(self, x0, ref, w0, speed, safety_distance):
        B = x0.shape[0]
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        w = np.array(w0, dtype=np.float64, copy=True)
        info = np.zeros(B, dtype=capi.INFO_DTYPE)
        replan = np.zeros(B, dtype=np.int32)
        costs = np.zeros(B, dtype=np.float64)
        self._ck(self.L.ampc_multi_round_batch(self.m, B, x0.ctypes.data, ref.ctypes.data, None, speed, safety_distance,
                                               w.ctypes.data, info.ctypes.data, replan.ctypes.data, costs.ctypes.data))
        return w, info, replan, costs

    def close(self):
        for hv in getattr(self, "handles", []):
            hv.close()
        if self.m:
            self.L.ampc_multi_destroy(self.m)
            self.m = _vp()
