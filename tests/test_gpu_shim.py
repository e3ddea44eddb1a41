"""GPU: the C++ drop-in classes (ObstacleAvoidanceMPC, KDTreeTwo, FrameKDMap, tick loop)
through their own test program, which calls libampc's C-ABI exactly as a ROS node would."""
import os
import subprocess

import pytest

import avoid_mpc_b200 as A

pytestmark = pytest.mark.gpu


def test_cpp_shim_end_to_end():
    exe = os.path.join(os.path.dirname(A.capi.LIB_PATH), "test_shim")
    if not os.path.exists(exe):
        A.capi.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "SHIM_OK" in r.stdout
