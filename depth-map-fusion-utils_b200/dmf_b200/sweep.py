"""Sharding a candidate-view sweep over the GPUs of one box.

Views are independent units: the grid is replicated on every GPU, rank g casts the contiguous block
[g*V/G, (g+1)*V/G) of the pose list, and the per-view visibility bitsets are combined with one all-gather
(gathered row order == view order, which is what greedy set cover's lowest-index tie-break needs).  A fused
"seen" map is the bitwise OR of all rows.  torch.distributed is only plumbing here: NCCL on GPUs, gloo in the
CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank `rank`: [rank*n/world, (rank+1)*n/world) with integer floor, so blocks tile [0,n)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return (rank * n_views) // world, ((rank + 1) * n_views) // world


def gather_rows(local, n_total: int, group=None):
    """All-gather row blocks produced under shard_range into the full [n_total, words] array (torch tensor in,
    torch tensor out, on the same device).  Ragged blocks are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        assert local.shape[0] == n_total
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(b - a for a, b in sizes)
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local block does not match shard_range"
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(b - a == mx for a, b in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (b - a)] for r, (a, b) in enumerate(sizes)], 0)


def or_rows(bitsets) -> np.ndarray:
    """Bitwise OR over views: the fused "seen" map of a sweep (numpy, host)."""
    b = np.asarray(bitsets)
    return np.bitwise_or.reduce(b, axis=0) if len(b) else np.zeros(b.shape[1:], b.dtype)


def sweep_visibility(engine, volume, poses, mode: int, zdelta: int, sparse: bool = False, reverse: bool = False, group=None) -> np.ndarray:
    """Visibility bitsets [n_views, words] of the whole pose list, computed on this rank's shard and all-gathered.
    Every rank returns the full array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    a, b = shard_range(len(poses), rank, world)
    if reverse:
        local = engine.reverse_views(volume, poses[a:b], fast=True, want=("visibility",))["visibility"]
    else:
        local = engine.forward_views(volume, poses[a:b], mode, zdelta, sparse, want=("visibility",))["visibility"]
    if world == 1:
        return local
    t = torch.from_numpy(local.view(np.int64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    return gather_rows(t, len(poses), group).cpu().numpy().view(np.uint64)
