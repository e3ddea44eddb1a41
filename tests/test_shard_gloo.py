"""CPU, world_size 2 over gloo: scene sharding + the all-gather of per-instance costs +
best-of-G reduce to what a single process computes."""
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import avoid_mpc_b200 as A


def test_scene_ranges_partition():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 1024, 65536):
            r = [A.shard.scene_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
            for s in range(0, n, max(1, n // 13)):
                o = A.shard.owner_of(s, world, n)
                assert r[o][0] <= s < r[o][1]


def _worker(rank, world, port, n_scenes, G, q):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = A.shard.scene_range(rank, world, n_scenes)
    rng = np.random.default_rng(7)
    costs_all = rng.uniform(1, 100, n_scenes * G)
    mine = torch.tensor(costs_all[lo * G:hi * G])
    out = A.shard.gather_costs(mine, world)
    q.put((rank, out.numpy().copy()))
    dist.destroy_process_group()


def test_gather_costs_world2():
    n_scenes, G, world = 16, 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + np.random.randint(0, 2000)
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, G, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.random.default_rng(7).uniform(1, 100, n_scenes * G)
    for r in range(world):
        assert (res[r] == want).all()
    status = np.zeros(n_scenes * G, dtype=np.int32)
    status[5] = 3
    arg, best = A.shard.best_of_scenes(res[0], status, G)
    c = want.reshape(n_scenes, G).copy()
    c[1, 1] = np.inf
    assert (arg == c.argmin(1)).all() and (best == c.min(1)).all()
