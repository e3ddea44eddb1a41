"""GPU: the in-process multi-device driver (include/ampc_multi.h) gives, instance for instance,
what one handle gives -- on one device, and on two when the box has them."""
import numpy as np
import pytest
import torch

import avoid_mpc_b200 as A
from helpers import make_instances

pytestmark = pytest.mark.gpu
D, S = A.defaults, A.synth


@pytest.mark.parametrize("n_dev", [1, 2])
def test_multi_round_matches_single_handle(n_dev):
    if torch.cuda.device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    N, K, B, npts = 20, 16, 13, 4096  # 13 instances: uneven blocks (7 + 6)
    inst = make_instances(range(700, 700 + B), N, K, npts)
    clouds = np.stack(inst["clouds"])
    W0 = np.stack([S.warm_start("ref", inst["x0"][b], inst["ref"][b], N) for b in range(B)])
    h = A.Handle(N=N, K=K, max_batch=B, max_points=npts)
    h.cloud_set_batch(clouds)
    W1, info1, replan1 = h.round(inst["x0"], inst["ref"], W0, speed=D.SPEED, safety_distance=D.SAFETY_DISTANCE)
    h.close()
    m = A.multi.MultiHandle(list(range(n_dev)), N=N, K=K, max_batch=B, max_points=npts)
    assert [m.shard(B, i) for i in range(n_dev)] == ([(0, 13)] if n_dev == 1 else [(0, 7), (7, 6)])
    m.cloud_set_batch(clouds)
    W2, info2, replan2, costs = m.round(inst["x0"], inst["ref"], W0, D.SPEED, D.SAFETY_DISTANCE)
    m.close()
    assert (W1 == W2).all() and (replan1 == replan2).all()
    assert (info1["status"] == info2["status"]).all() and (info1["iters"] == info2["iters"]).all()
    assert (costs == info1["cost"]).all()  # what came back through the NCCL all-gather
