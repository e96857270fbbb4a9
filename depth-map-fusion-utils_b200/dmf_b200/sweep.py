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


def shard_indices(n_views: int, rank: int, world: int, layout: str = "block") -> np.ndarray:
    """View indices of rank `rank`.  "block": the contiguous block of shard_range.  "strided": rank, rank+world, ... --
    neighbouring views of a sweep cost about the same, so interleaving them evens out the per-rank work (a step is as
    slow as its slowest rank) at no cost: views are independent."""
    if layout == "block":
        a, b = shard_range(n_views, rank, world)
        return np.arange(a, b)
    if layout != "strided":
        raise ValueError(f"unknown layout {layout!r}")
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return np.arange(rank, n_views, world)


def gather_rows(local, n_total: int, group=None, layout: str = "block"):
    """All-gather the rows each rank produced for shard_indices(..., layout) into the full [n_total, words] array in VIEW
    order (torch tensor in, torch tensor out, on the same device).  Ragged shards are padded for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        assert local.shape[0] == n_total
        return local
    if layout == "strided":
        counts = [len(range(r, n_total, world)) for r in range(world)]
        mx = max(counts)
        assert local.shape[0] == counts[rank], "local rows do not match shard_indices"
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        # row i of rank r is view i*world + r: transposing (rank, i) -> (i, rank) restores view order; padding rows fall at the end
        out = out.reshape((world, mx) + tuple(local.shape[1:])).transpose(0, 1).reshape((world * mx,) + tuple(local.shape[1:]))
        return out[:n_total].contiguous()
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


def or_allreduce(t, group=None, or_rows_dev=None):
    """Bitwise-OR all-reduce of a packed bit grid (integer torch tensor, same shape on every rank), in place.
    gloo reduces with BOR directly.  NCCL has no bitwise-OR: the grids are all-gathered (NVSwitch: every rank receives the
    others' rows at full link rate) and OR-ed locally by `or_rows_dev(dst, rows)` -- K5 k_or_reduce via dmf_or_reduce_dev."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return t
    if dist.get_backend(group) != "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.BOR, group=group)
        return t
    rows = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(rows, t.contiguous(), group=group)
    if or_rows_dev is None:
        raise ValueError("or_allreduce over NCCL needs or_rows_dev (the device OR kernel)")
    or_rows_dev(t, rows)
    return t


def fuse_observed(ctx, group=None) -> dict:
    """Carve mode over several GPUs: every rank has marched its share of the views into its own observed-voxel bit grid;
    afterwards every rank's grid is the union (the fused free/occupied map of the whole sweep).  Returns observed_counts()."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from ._lib import check
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        n64 = ctx.lib.dmf_observed_words(ctx.h) // 2                     # the grid is a whole number of 256-bit blocks
        ptr = ctx.observed_dev_ptr()                                     # synchronises the context's stream
        ctx.synchronize()
        dev = torch.device("cuda", ctx.device)
        mine = torch.zeros(n64, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)                                       # `mine` is zeroed before the context's stream touches it
        # both kernels run on the context's own stream (NULL = that stream in the C ABI); the collective runs on torch's.
        # A one-off step at the end of a sweep: plain synchronisation between the three, no event plumbing.
        check(ctx.lib.dmf_or_reduce_dev(ctx.h, C.c_void_p(mine.data_ptr()), C.c_void_p(ptr), 1, n64, None))           # export: 0 | grid
        ctx.synchronize()

        def or_rows_dev(dst, rows):
            torch.cuda.synchronize(dev)                                   # the all-gather has landed
            check(ctx.lib.dmf_or_reduce_dev(ctx.h, C.c_void_p(ptr), C.c_void_p(rows.data_ptr()), world, n64, None))   # merge into the context
            ctx.synchronize()
        or_allreduce(mine, group, or_rows_dev)
    return ctx.observed_counts()


def sweep_visibility(engine, volume, poses, mode: int, zdelta: int, sparse: bool = False, reverse: bool = False, group=None,
                     layout: str = "strided") -> np.ndarray:
    """Visibility bitsets [n_views, words] of the whole pose list in view order, computed on this rank's shard and
    all-gathered.  Every rank returns the full array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    mine = poses[shard_indices(len(poses), rank, world, layout)]
    if reverse:
        local = engine.reverse_views(volume, mine, fast=True, want=("visibility",))["visibility"]
    else:
        local = engine.forward_views(volume, mine, mode, zdelta, sparse, want=("visibility",))["visibility"]
    if world == 1:
        return local
    t = torch.from_numpy(local.view(np.int64))
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    return gather_rows(t, len(poses), group, layout).cpu().numpy().view(np.uint64)


def bind_to_gpu_numa_node(device: int) -> str:
    """Pin the calling process to the CPUs that are local to GPU `device` (sysfs local_cpulist of its PCI function), so
    that pinned host buffers allocated afterwards land on that socket's memory: with one rank per GPU, device-to-host
    copies then do not cross the inter-socket link.  Returns a description; a no-op when the topology cannot be read."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"gpu {device} ({bdf}): local cpus {txt or '?'} = current affinity, nothing to do"
        os.sched_setaffinity(0, cpus)
        return f"gpu {device} ({bdf}): bound to its {len(cpus)} local cpus ({txt})"
    except Exception as e:                                             # noqa: BLE001 - topology is best effort
        return f"gpu {device}: not bound ({type(e).__name__}: {e})"
