"""The numpy restatement of ProcessDepth / BuildEdgeCloud (oracle/depth_oracle.py) against the
golden vectors minted with OpenCV in the reference's call pattern
(tests/golden/make_depth_golden.py): bit-exact clouds, every case."""
import os

import numpy as np
import pytest

from oracle import depth_oracle as DO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "depth_golden.npz"))
NAMES = [str(n) for n in G["names"]]


def golden_case(name):
    fx, fy, cx, cy, scale, p2m, dmin, dmax = G[name + "_params"]
    cam = DO.Camera(fx, fy, cx, cy, scale, p2m, dmin, dmax)
    T_obst = DO.matmul4(G[name + "_Twb"], DO.TBC)
    T_edge = DO.matmul4(G[name + "_Twc_prev"], DO.TBC)
    return G[name + "_depth"], cam, T_obst, T_edge, G[name + "_cloud"], G[name + "_edge"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_opencv_golden(name):
    depth, cam, T_obst, T_edge, cloud, edge = golden_case(name)
    c, e = DO.process_depth(depth, cam, T_obst, T_edge)
    assert c.shape == cloud.shape and e.shape == edge.shape
    assert np.array_equal(c.view(np.uint32), cloud.view(np.uint32))
    assert np.array_equal(e.view(np.uint32), edge.view(np.uint32))


def test_edge_cases():
    cam = DO.Camera(32, 32, 32, 24, 1.0)
    T = np.eye(4)
    # nothing in range -> both clouds empty (FrameKDMap.cpp:125-127)
    c, e = DO.process_depth(np.full((48, 64), 500.0, np.float32), cam, T, T)
    assert len(c) == 0 and len(e) == 0
    # a flat wall: every pixel valid, no depth edge except the image border response of Canny
    c, e = DO.process_depth(np.full((48, 64), 5.0, np.float32), cam, T, T)
    assert len(c) == 48 * 64 and len(e) == 0
    assert np.allclose(c[:, 2], 5.0)
    # the averaging quirk: a 2x2 block with one invalid pixel halves... the inverse depth mean
    d = np.full((2, 2), 4.0, np.float32)
    d[0, 0] = 1e4
    cam2 = DO.Camera(2, 2, 1, 1, 2.0)
    c, _ = DO.process_depth(d, cam2, T, T)
    assert len(c) == 1 and np.isclose(c[0, 2], 1.0 / (0.75 * 0.25))
