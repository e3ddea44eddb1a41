"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol
include/ampc.h declares, and refuses loudly to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import avoid_mpc_b200 as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    A.capi.build()
    return ctypes.CDLL(A.capi.LIB_PATH)


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "ampc.h")).read()
    declared = set(re.findall(r"\b(ampc_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(A.capi.SYMBOLS), declared ^ set(A.capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"libampc.so does not export {name}"
    lib.ampc_api_version.restype = ctypes.c_int
    assert lib.ampc_api_version() == 1


def test_multi_device_library_exports_every_declared_symbol(lib):
    """include/ampc_multi.h -> libampc_multi.so (host C++ + NCCL over libampc's C-ABI)."""
    hdr = open(os.path.join(ROOT, "include", "ampc_multi.h")).read()
    declared = set(re.findall(r"\b(ampc_multi_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(A.multi.SYMBOLS), declared ^ set(A.multi.SYMBOLS)
    L = A.multi.lib()
    for name in declared:
        assert hasattr(L, name), f"libampc_multi.so does not export {name}"
    import torch
    if not torch.cuda.is_available():  # no device: creation fails loudly with the CUDA error code
        with pytest.raises(A.AmpcError) as ei:
            A.multi.MultiHandle([0], max_batch=4, max_points=64)
        assert ei.value.code == A.capi.ERR_CUDA


def test_struct_layouts_match_header():
    assert ctypes.sizeof(A.capi.Config) == 40
    assert ctypes.sizeof(A.capi.SolverOpts) == 64
    assert A.capi.INFO_DTYPE.itemsize == 48
    o = A.capi.default_solver_opts()
    assert (o.tol, o.max_iter, o.mu_init, o.bound_push, o.eps_min) == (1e-8, 100, 0.1, 1e-2, 1e-5)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(A.AmpcError) as ei:
        A.Handle(N=20, K=16, max_batch=4, max_points=64)
    assert ei.value.code == A.capi.ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The product package may not reference oracle/ anywhere."""
    pkg = os.path.join(ROOT, "avoid-mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), (dp, f)
