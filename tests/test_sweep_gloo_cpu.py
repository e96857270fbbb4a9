"""CPU tier, world_size 2 over gloo: the view-sharding plumbing of a multi-GPU sweep (shard_range + all-gather of
per-view visibility bitsets + OR combine + set cover on the gathered rows).  The per-rank compute is stood in by the
CPU oracle (allowed in tests/ only); on GPUs the same plumbing runs over NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_tiles_exactly(dmf):
    from dmf_b200.sweep import shard_range
    for n in (0, 1, 7, 8, 1024, 10000):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, q, layout="block"):
    sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle_py as O
    from dmf_b200 import scenes
    from dmf_b200.sweep import gather_rows, or_rows, shard_indices
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scenes.scene("S64")
    vol = O.volume_from_scene(sc)
    occ = vol.occupied()
    index = {int(h): i for i, h in enumerate(occ)}
    K = scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.125
    poses = scenes.poses_sphere_lookat(1.024, 200)[:: 200 // n_views][:n_views]
    mine = shard_indices(n_views, rank, world, layout)
    words = (len(occ) + 63) // 64
    local = np.zeros((len(mine), words), np.uint64)
    for j, p in enumerate(poses[mine]):
        for h in O.forward(vol, K, 60, 80, p, O.MODE_POINTS, 8, False, want_pixels=False)["ids"]:
            i = index[int(h)]
            local[j, i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    full = gather_rows(torch.from_numpy(local.view(np.int64)), n_views, layout=layout).numpy().view(np.uint64)
    q.put((rank, full, or_rows(full)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_indices_partition_the_views(dmf):
    from dmf_b200.sweep import shard_indices
    for layout in ("block", "strided"):
        for n in (0, 1, 7, 8, 1000):
            for world in (1, 2, 3, 8):
                parts = [shard_indices(n, r, world, layout) for r in range(world)]
                assert sorted(np.concatenate(parts).tolist()) == list(range(n))
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert shard_indices(10, 1, 4, "strided").tolist() == [1, 5, 9]
    with pytest.raises(ValueError):
        shard_indices(10, 0, 2, "diagonal")


@pytest.mark.parametrize("n_views,layout", [(6, "block"), (7, "block"), (6, "strided"), (7, "strided")])   # even and ragged splits over 2 ranks
def test_two_rank_sweep_equals_single_process(dmf, oracle, n_views, layout):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q, layout)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, full, seen = q.get(timeout=240)
        got[r] = (full, seen)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    # single-process answer
    sc = dmf.scenes.scene("S64")
    vol = oracle.volume_from_scene(sc)
    occ = vol.occupied()
    index = {int(h): i for i, h in enumerate(occ)}
    K = dmf.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.125
    poses = dmf.scenes.poses_sphere_lookat(1.024, 200)[:: 200 // n_views][:n_views]
    want = np.zeros((n_views, (len(occ) + 63) // 64), np.uint64)
    sets = []
    for j, p in enumerate(poses):
        ids = oracle.forward(vol, K, 60, 80, p, oracle.MODE_POINTS, 8, False, want_pixels=False)["ids"]
        sets.append(np.sort(ids))
        for h in ids:
            i = index[int(h)]
            want[j, i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    assert np.array_equal(got[0][0], want)
    assert np.array_equal(got[0][1], np.bitwise_or.reduce(want, axis=0))
    # the gathered rows are in view order, so set cover over them selects the same views as the reference pipeline
    sel = oracle.greedy_set_cover(sets)
    pop = [int(np.unpackbits(r.view(np.uint8)).sum()) for r in want]
    assert pop == [len(s) for s in sets] and len(sel) >= 1


def _carve_worker(rank, world, port, n_views, q):
    sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle_py as O
    from dmf_b200 import scenes
    from dmf_b200.sweep import or_allreduce, shard_indices
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = scenes.scene("S64")
    vol = O.volume_from_scene(sc)
    K = scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.125
    poses = scenes.poses_sphere_lookat(1.024, 200)[:: 200 // n_views][:n_views]
    obs, _ = O.forward_observed(vol, K, 60, 80, poses[0], O.MODE_POINTS, 8, False)
    obs[:] = 0
    for p in poses[shard_indices(n_views, rank, world, "strided")]:
        obs, _ = O.forward_observed(vol, K, 60, 80, p, O.MODE_POINTS, 8, False, observed=obs)
    fused = or_allreduce(torch.from_numpy(obs.view(np.int32).copy())).numpy().view(np.uint32)
    q.put((rank, obs, fused))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_observed_grids_fuse_to_the_single_process_grid(dmf, oracle):
    """carve mode over 2 ranks: each rank's observed-voxel grid covers its share of the views; the OR all-reduce leaves the
    union on both -- equal to the grid one process builds from all views"""
    import torch.multiprocessing as mp
    n_views = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_carve_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        r, own, fused = q.get(timeout=240)
        got[r] = (own, fused)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = dmf.scenes.scene("S64")
    vol = oracle.volume_from_scene(sc)
    K = dmf.scenes.REFERENCE_K.copy(); K[[0, 2, 4, 5]] *= 0.125
    poses = dmf.scenes.poses_sphere_lookat(1.024, 200)[:: 200 // n_views][:n_views]
    want = None
    for p in poses:
        want, _ = oracle.forward_observed(vol, K, 60, 80, p, oracle.MODE_POINTS, 8, False, observed=want)
    assert np.array_equal(got[0][1], want) and np.array_equal(got[1][1], want)
    assert not np.array_equal(got[0][0], want) and np.array_equal(got[0][0] | got[1][0], want)
