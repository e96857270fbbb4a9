"""Checks the multi-GPU sweep of libdmf_b200 (dmf_comm_*, dmf_sweep_*) against the single-GPU entry points.

    python tools/comm_check.py all [n_gpus]          one process drives all GPUs (dmf_comm_init_all)
    torchrun ... tools/comm_check.py rank            one process per GPU (dmf_comm_init_rank; the id travels over gloo)

Every check compares bit for bit: gathered visibility rows == dmf_forward / dmf_reverse rows of one GPU, set cover on the
gathered rows == dmf_greedy_set_cover, fused observed grid == the grid one GPU builds from all views, fused marks ==
the marks of one batched rayTraceAndClassify call.  Prints OK per process.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import dmf_b200 as D  # noqa: E402

SCENE = os.environ.get("DMF_CHECK_SCENE", "S128")
N_VIEWS = int(os.environ.get("DMF_CHECK_VIEWS", "37"))          # ragged over 2, 4 and 8 ranks on purpose


def build_volume(ctx):
    sc = D.scenes.scene(SCENE)
    vol = D.VoxelVolume(ctx)
    vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
    vol._commit(ctx)
    return sc, vol


def single_gpu_reference(ctx, sc, vol, poses):
    """everything the sweeps are compared against, computed by ONE context through the ordinary entry points"""
    eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K, 480, 640), ctx, D.GRID_BYTE)
    ref = {}
    r = eng.forward_views(vol, poses, D.MODE_POINTS, sc.zdelta, False, want=("visibility",))
    ref["fwd_vis"], ref["fwd_found"] = r["visibility"], r["found_any"]
    r = eng.forward_views(vol, poses, D.MODE_GOOD_POINTS, sc.zdelta, False, want=("visibility",))
    ref["good_vis"], ref["good_found"] = r["visibility"], r["found_any"]
    r = eng.reverse_views(vol, poses, fast=True, want=("visibility",))
    ref["rev_vis"], ref["rev_found"] = r["visibility"], r["found_any"]
    ref["cover"] = D.greedySetCover(ref["rev_vis"], ctx)
    ctx.clear_observed()
    eng.forward_views(vol, poses, D.MODE_POINTS, sc.zdelta, False, want=(), carve=True)
    ref["observed"] = ctx.observed_words().copy()
    vol.clear_marks()
    eng.forward_views(vol, poses, D.MODE_CLASSIFY, sc.zdelta, False, view_id0=7, want=())
    ref["marks"] = vol.marks()
    vol.clear_marks(); ctx.clear_observed()
    return ref


def run_checks(comm, ctx_list, sc, poses, ref, tag):
    info = comm.info()
    n = len(poses)
    own = np.arange(info["first_rank"], n, info["world"]) if info["n_local"] == 1 and info["world"] > 1 else np.arange(n)
    for rows_mode in ((D.comm.ROWS_ALL, D.comm.ROWS_OWN) if info["n_local"] == 1 else (D.comm.ROWS_ALL,)):
        sel = np.arange(n) if rows_mode == D.comm.ROWS_ALL else own
        for _ in range(3):                     # three passes: both parities of the double buffer, and re-use
            r = comm.sweep_forward(poses, D.MODE_POINTS, sc.zdelta, False, rows_to_host=rows_mode)
            assert np.array_equal(r["visibility"][sel], ref["fwd_vis"][sel]), f"{tag}: forward rows differ"
            assert np.array_equal(r["found_any"][sel], ref["fwd_found"][sel]), f"{tag}: forward found_any differs"
        r = comm.sweep_forward(poses, D.MODE_GOOD_POINTS, sc.zdelta, False, rows_to_host=rows_mode)
        assert np.array_equal(r["visibility"][sel], ref["good_vis"][sel]) and np.array_equal(r["found_any"][sel], ref["good_found"][sel]), f"{tag}: good-points rows differ"
        r = comm.sweep_reverse(poses, rows_to_host=rows_mode)
        assert np.array_equal(r["visibility"][sel], ref["rev_vis"][sel]) and np.array_equal(r["found_any"][sel], ref["rev_found"][sel]), f"{tag}: reverse rows differ"
    cover = comm.set_cover()               # over the rows every GPU now holds
    assert np.array_equal(cover, ref["cover"]), f"{tag}: set cover {cover} != {ref['cover']}"
    # an empty and a one-view sweep (ranks without work still take part in the exchange)
    assert comm.sweep_reverse(poses[:0])["visibility"].shape[0] == 0
    r = comm.sweep_reverse(poses[:1])
    assert np.array_equal(r["visibility"][:1][own[own < 1]], ref["rev_vis"][:1][own[own < 1]])
    # carve: observed grids fused over the group == one GPU's grid from all views
    for c in ctx_list:
        c.clear_observed()
    comm.sweep_forward(poses, D.MODE_POINTS, sc.zdelta, False, flags=D.FWD_CARVE, want_rows=False)
    comm.fuse_observed()
    for c in ctx_list:
        assert np.array_equal(c.observed_words(), ref["observed"]), f"{tag}: fused observed grid differs"
    # classify: Voxel::view first-wins (min over the GPUs), Voxel::good OR
    n_occ = len(ref["marks"][0])
    for c in ctx_list:
        D._lib.check(c.lib.dmf_clear_marks(c.h))
    comm.sweep_forward(poses, D.MODE_CLASSIFY, sc.zdelta, False, view_id0=7, want_rows=False)
    comm.fuse_marks(7)
    import ctypes as C
    for c in ctx_list:
        view, good = np.zeros(n_occ, np.int32), np.zeros(n_occ, np.uint8)
        D._lib.check(c.lib.dmf_download_marks(c.h, view.ctypes.data_as(C.POINTER(C.c_int32)), good.ctypes.data_as(C.POINTER(C.c_uint8)), n_occ))
        assert np.array_equal(view, ref["marks"][0]), f"{tag}: fused Voxel::view differs ({int((view != ref['marks'][0]).sum())} voxels)"
        assert np.array_equal(good, ref["marks"][1]), f"{tag}: fused Voxel::good differs"


def main_all(n_gpus):
    comm = D.Comm.init_all(n_gpus)
    info = comm.info()
    ctx0 = comm.contexts[0]
    sc, vol = build_volume(ctx0)
    poses = D.scenes.bench_poses(float(sc.bounds[1]), N_VIEWS)
    comm.set_camera(D.scenes.REFERENCE_K, 480, 640)
    ref = single_gpu_reference(ctx0, sc, vol, poses)
    comm.replicate_volume(0)                                     # GPU 0 -> the others, GPU to GPU
    for c in comm.contexts[1:]:                                  # the replica is the same volume
        ids = np.zeros(len(vol.occupied_cells_), np.uint64)
        import ctypes as C
        D._lib.check(c.lib.dmf_volume_get_occupied(c.h, ids.ctypes.data_as(C.POINTER(C.c_uint64))))
        assert np.array_equal(ids, vol.occupied_cells_), "replicated occupied list differs"
    run_checks(comm, comm.contexts, sc, poses, ref, f"init_all[{info['world']}]")
    print(f"OK single process, {info['world']} GPUs, exchange: {info['exchange_name']}", flush=True)
    comm.close()


def main_rank():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")                               # only the out-of-band channel for the 128-byte id
    torch.cuda.set_device(local)
    ctx = D.Context(local)
    uid = [D.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = D.Comm.init_rank(ctx, uid[0], rank, world)
    info = comm.info()
    sc = D.scenes.scene(SCENE)
    poses = D.scenes.bench_poses(float(sc.bounds[1]), N_VIEWS)
    comm.set_camera(D.scenes.REFERENCE_K, 480, 640)
    vol = None
    if rank == 0:
        sc, vol = build_volume(ctx)
    comm.replicate_volume(0)                                     # rank 0's volume -> every rank (ncclBroadcast of the id list)
    if vol is None:                                               # a host-side handle on the replica for the single-GPU reference
        vol = D.VoxelVolume(ctx)
        vol.occupied_cells_ = np.zeros(ctx.lib.dmf_visibility_words(ctx.h) * 0, np.uint64)
        import ctypes as C
        no = C.c_size_t()
        D._lib.check(ctx.lib.dmf_volume_info(ctx.h, None, None, None, C.byref(no), None))
        vol.occupied_cells_ = np.zeros(no.value, np.uint64)
        D._lib.check(ctx.lib.dmf_volume_get_occupied(ctx.h, vol.occupied_cells_.ctypes.data_as(C.POINTER(C.c_uint64))))
        vol._dirty = False; ctx._volume_token = vol; vol.ctx = ctx
    ref = single_gpu_reference(ctx, sc, vol, poses)                # every rank computes the whole reference on its own replica
    run_checks(comm, [ctx], sc, poses, ref, f"rank {rank}/{world}")
    dist.barrier()
    print(f"OK rank {rank}/{world}, exchange: {info['exchange_name']}", flush=True)
    comm.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        main_all(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    else:
        main_rank()
