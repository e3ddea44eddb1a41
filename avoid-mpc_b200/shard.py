"""Multi-GPU plumbing: the batch of MPC instances (scenes x initial guesses) shards
across ranks with no exchange during k-NN or solve.  The only collective
is an all-gather of the per-instance costs for the best-cost reduction.  torch.distributed
(NCCL on GPUs, gloo in the CPU tests) is plumbing; nothing here is on the measured path
except `gather_costs`."""
from __future__ import annotations


def scene_range(rank: int, world: int, n_scenes: int) -> tuple[int, int]:
    """Contiguous scene block of `rank`: scene s lives on rank floor(s*world/n_scenes)
    (all G guesses of a scene stay on one GPU, so best-of-G needs no communication)."""
    if not (0 <= rank < world) or n_scenes < 0:
        raise ValueError("bad rank/world/n_scenes")
    lo = (rank * n_scenes + world - 1) // world
    hi = ((rank + 1) * n_scenes + world - 1) // world
    return lo, hi


def owner_of(scene: int, world: int, n_scenes: int) -> int:
    return scene * world // n_scenes


def gather_costs(costs, world: int, out=None):
    """All-gather equally sized per-rank cost vectors (1-D float64 tensors)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return costs
    if out is None:
        out = torch.empty(costs.numel() * world, dtype=costs.dtype, device=costs.device)
    dist.all_gather_into_tensor(out, costs)
    return out


def best_of_scenes(costs, status, G: int):
    """Reference semantics of ampc_best_of on host tensors/arrays: per scene, the guess with
    the lowest cost among instances whose status is CONVERGED(0) or MAX_ITER(1); -1 if none."""
    import numpy as np

    c = np.asarray(costs, dtype=np.float64).reshape(-1, G).copy()
    s = np.asarray(status).reshape(-1, G)
    c[~((s == 0) | (s == 1))] = np.inf
    arg = c.argmin(axis=1).astype(np.int32)
    best = c.min(axis=1)
    arg[~np.isfinite(best)] = -1
    return arg, best
