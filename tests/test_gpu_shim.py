"""GPU: the C++ drop-in classes (ObstacleAvoidanceMPC, KDTreeTwo, FrameKDMap, tick loop)
through their own test program, which calls libampc's C-ABI exactly as a ROS node would."""
import os
import subprocess

import numpy as np
import pytest

import avoid_mpc_b200 as A
from oracle import depth_oracle as DO

pytestmark = pytest.mark.gpu


def _depth_case(path):
    """Two depth frames + the clouds the oracle expects FrameKDMap::AddVertex to hold after the
    second one (Obstacle: Twb1 * Tbc; Edge: the previous frame's Twc * Tbc, FrameKDMap.cpp:208-209)."""
    rows, cols = 240, 320
    par = dict(fx=160.0, fy=160.0, cx=160.0, cy=120.0, resize_scale=5.0, pixel2meter=1.0, depth_min=0.1,
               depth_max=100.0)
    cam = DO.Camera(**par)
    Twb0, Twb1 = np.eye(4), np.eye(4)
    Twb0[:3, 3], Twb1[:3, 3] = (0, 0, 1.5), (0.3, -0.1, 1.6)
    yaw = 0.1
    Twb1[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    d0, d1 = A.synth.forest_depth(7, rows, cols), A.synth.forest_depth(8, rows, cols)
    Twc0 = DO.matmul4(Twb0, DO.TBC)
    cloud, edge = DO.process_depth(d1, cam, DO.matmul4(Twb1, DO.TBC), DO.matmul4(Twc0, DO.TBC))
    with open(path, "wb") as f:
        f.write(np.array([rows, cols, len(cloud), len(edge)], np.int32).tobytes())
        f.write(np.array(list(par.values()), np.float64).tobytes())
        f.write(Twb0.tobytes() + Twb1.tobytes() + d0.tobytes() + d1.tobytes() + cloud.tobytes() + edge.tobytes())


def test_cpp_shim_end_to_end(tmp_path):
    exe = os.path.join(os.path.dirname(A.capi.LIB_PATH), "test_shim")
    if not os.path.exists(exe):
        A.capi.build()
    case = str(tmp_path / "depth_case.bin")
    _depth_case(case)
    r = subprocess.run([exe, case], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "SHIM_OK" in r.stdout
