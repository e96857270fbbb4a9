"""bench.py --config 3 | 4: the two large BASELINE.json configurations as first-class runs (same JSON contract, same Rig).

config 3 -- moving-camera fusion: every step integrates V views per GPU of the helix trajectory P10k (subsampled evenly, dealt
  round-robin over the GPUs) into the 1024^3 grid: rayTraceAndClassify marks (Voxel::view / Voxel::good, RayTracingEngine.hpp:311-375)
  AND the occupied/free update (carve) in one launch.  After the trajectory the marks are fused over the GPUs (first view = min,
  good = OR) and so are the observed grids (OR), as reduce-scatters over peer memory (dmf_comm_fuse_marks / _observed).
config 4 -- high-res stress: V views per GPU of the Fibonacci sweep P4096, 1920x1080, rayTraceAndGetPoints semantics; the visibility
  rows are gathered on every GPU by the march kernels and OR-reduced into the "seen" map.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import time

import numpy as np


def _marks(rig):
    from dmf_b200._lib import check
    view, good = np.zeros(rig.n_occ, np.int32), np.zeros(rig.n_occ, np.uint8)
    check(rig.ctx.lib.dmf_download_marks(rig.ctx.h, view.ctypes.data_as(C.POINTER(C.c_int32)), good.ctypes.data_as(C.POINTER(C.c_uint8)), rig.n_occ))
    return view, good


def _cpu_one_view(B, args, rig, mode):
    """the reference's own headers on ONE view of the step (bounded sample: a 1024^3 pointer grid and ~0.3-2 G samples per view)"""
    M, kind, how = B.cpu_backend()
    t0 = time.perf_counter()
    vol = B.cpu_volume(M, rig.sc)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = M.forward(vol, rig.K, rig.H, rig.W, rig.all_poses[0], mode, rig.sc.zdelta, False, 1)
    sec = time.perf_counter() - t0
    out = {"value": rig.H * rig.W / sec, "unit": "rays/s", "cores": 1, "kind": kind, "sec_per_view": sec, "volume_build_s": build_s,
           "sample": f"view 0 of the step, 1 thread; {how}"}
    return out, r, vol


def run_config34(args, B):
    rig = B.Rig(args)
    torch, dist, D = rig.torch, rig.dist, rig.D
    from dmf_b200._lib import ForwardOut, ForwardParams, SweepOut, check
    ctx, comm, lib = rig.ctx, rig.comm, rig.ctx.lib
    H, W, V, vw, n_occ, world, rank, dev, sc, fmt = rig.H, rig.W, rig.V, rig.vw, rig.n_occ, rig.world, rig.rank, rig.dev, rig.sc, rig.fmt
    n_total = rig.n_total
    fusion = args.config == 3
    mode = D.MODE_CLASSIFY if fusion else D.MODE_POINTS
    flags = D.FWD_CARVE if fusion else 0
    params = ForwardParams(mode, sc.zdelta, 0, 1, fmt, flags | (0 if args.count_in_timed else D.FWD_NO_COUNTERS))   # timed launches: probe counters off
    params_counted = ForwardParams(mode, sc.zdelta, 0, 1, fmt, flags)
    d_poses = torch.from_numpy(rig.my_poses).to(dev)
    poses_dev = B.ptr_array(d_poses.data_ptr())
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_dev(p=None):
        check(lib.dmf_sweep_forward_dev(comm.h, C.byref(p or params), poses_dev, n_total, None, B.ptr_array(torch.cuda.current_stream().cuda_stream)))

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    if fusion:
        ctx.clear_observed(); check(lib.dmf_clear_marks(ctx.h))
    ctx.reset_counters()
    sampler = B.ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    hot_ms = []
    rig.barrier(); torch.cuda.synchronize()
    for a, b in evs:
        if flush is not None:
            flush.fill_(1)
        a.record(); step_dev(); b.record()
        if rank == 0 and len(hot_ms) < 4:
            hot_ms.append(ctx.last_hot_kernel_ms())
    torch.cuda.synchronize(); rig.barrier()
    dev_ms = rig.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    launches_timed = ctx.counters()["launches"]
    ctx.reset_counters()
    step_dev(params_counted); torch.cuda.synchronize()                  # one identical untimed step with the probe counters on
    cnt = ctx.counters()
    inb1, samp1, launches_total, skip1 = rig.sum_over_ranks([cnt["inbounds"], cnt["samples"], launches_timed, cnt["skipped"]])
    inbounds_total, samples_total, skipped_total = inb1 * args.steps, samp1 * args.steps, skip1 * args.steps
    value = args.steps * V * H * W * world / (dev_ms * 1e-3)
    parity, extra = {}, {}

    # ---- the exchange step that ends a run, timed once: fused marks + observed grids (config 3) / OR-reduced visibility (config 4) ----
    if fusion:
        rig.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        comm.fuse_marks(1)
        torch.cuda.synchronize(); rig.barrier(); t1 = time.perf_counter()
        comm.fuse_observed()
        torch.cuda.synchronize(); rig.barrier(); t2 = time.perf_counter()
        view, good = _marks(rig)
        obs = ctx.observed_words()
        oc = ctx.observed_counts()
        extra["fuse"] = {"marks_ms": 1e3 * (t1 - t0), "observed_ms": 1e3 * (t2 - t1), "observed_grid_bytes": int(lib.dmf_observed_words(ctx.h)) * 4,
                         "how": "reduce-scatter + all-gather over peer memory: first-view min and good-bit OR over n_occ words, observed-grid OR over the bit grid" if world > 1 else "single GPU: nothing to fuse",
                         "voxels_marked": int((view != 0).sum()), "voxels_good": int(good.sum()), "observed_voxels": oc["observed"], "free_voxels": oc["free"], "hit_voxels": oc["hit"]}
        digest = hashlib.sha1(view.tobytes() + good.tobytes() + obs.tobytes()).hexdigest()
        if world > 1:
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            parity["fused_marks_and_grid_identical_on_all_ranks"] = len(set(digests)) == 1
        # rank 0 alone, all views of the step through the single-GPU host call: the fused result must equal it
        if rank == 0:
            ctx.clear_observed(); check(lib.dmf_clear_marks(ctx.h))
            rig.eng.forward_views(rig.vol, rig.all_poses, D.MODE_CLASSIFY, sc.zdelta, False, view_id0=1, want=(), carve=True)
            v1, g1 = _marks(rig)
            parity["fused_marks_equal_single_gpu_recompute"] = bool(np.array_equal(v1, view) and np.array_equal(g1, good))
            parity["fused_observed_grid_equal_single_gpu_recompute"] = bool(np.array_equal(ctx.observed_words(), obs))
    else:
        so = SweepOut()
        h_all = np.zeros((n_total, max(vw, 1)), np.uint64); h_fall = np.zeros(n_total, np.int32)
        so.visibility, so.found_any, so.rows_to_host = h_all.ctypes.data, h_fall.ctypes.data, D.comm.ROWS_ALL
        rig.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        check(lib.dmf_sweep_forward(comm.h, C.byref(params), rig.all_poses.ctypes.data_as(C.POINTER(C.c_float)), n_total, C.byref(so)))
        seen = np.bitwise_or.reduce(h_all, axis=0)
        torch.cuda.synchronize(); rig.barrier()
        extra["or_reduced_visibility"] = {"voxels_seen": int(np.unpackbits(seen.view(np.uint8)).sum()), "of_occupied": n_occ, "views": n_total,
                                          "ms_incl_rows_to_host": 1e3 * (time.perf_counter() - t0)}
        t0 = time.perf_counter()
        cover = comm.set_cover()
        extra["set_cover_over_gathered_rows"] = {"selected_views": [int(x) for x in cover], "ms": 1e3 * (time.perf_counter() - t0), "n_sets": n_total}
        digest = hashlib.sha1(h_all.tobytes() + h_fall.tobytes()).hexdigest()
        if world > 1:
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            parity["gathered_rows_identical_on_all_ranks"] = len(set(digests)) == 1
        if rank == 0:
            full = rig.eng.forward_views(rig.vol, rig.all_poses, D.MODE_POINTS, sc.zdelta, False, want=("visibility",))
            parity["gathered_rows_equal_single_gpu_recompute"] = bool(np.array_equal(h_all[:, :vw], full["visibility"]) and np.array_equal(h_fall, full["found_any"]))

    # ---- e2e: host poses in (pinned), per-view found flags (+ own visibility rows for config 4) back every step; config 3 adds the
    # fuse + the download of the marks once per run of `steps` steps, inside the timed region -------------------------------------
    h_poses, p0 = rig.pinned((n_total, 12), np.float32); h_poses[:] = rig.all_poses
    h_vis, p1 = rig.pinned((n_total, max(vw, 1)), np.uint64)
    h_found, p2 = rig.pinned((n_total,), np.int32)
    so = SweepOut()
    so.visibility = None if fusion else h_vis.ctypes.data
    so.found_any, so.rows_to_host = h_found.ctypes.data, D.comm.ROWS_OWN
    fp_all = h_poses.ctypes.data_as(C.POINTER(C.c_float))
    e2e_steps = max(2, min(args.steps, 10))
    check(lib.dmf_sweep_forward(comm.h, C.byref(params), fp_all, n_total, C.byref(so)))
    if fusion:
        ctx.clear_observed(); check(lib.dmf_clear_marks(ctx.h))
    rig.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        check(lib.dmf_sweep_forward(comm.h, C.byref(params), fp_all, n_total, C.byref(so)))
    if fusion:
        comm.fuse_marks(1); comm.fuse_observed()
        _marks(rig); ctx.observed_counts()
    torch.cuda.synchronize(); rig.barrier()
    e2e_s = rig.max_over_ranks(time.perf_counter() - t0)
    e2e_value = e2e_steps * V * H * W * world / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = B.measured_peak()
        per_launch_inb = cnt["inbounds"]
        # 1 B distance byte (+ 1/8 B observed bit in carve mode) per in-bounds sample; per view the visibility row + pose
        alg_bytes = per_launch_inb * (1.125 if fusion else 1.0) + V * (vw * 8 + 48)
        hot = float(np.mean(hot_ms)) if hot_ms else dev_ms / args.steps
        line = {
            "metric": "rays/s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": {"workload": rig.cfg["workload"], "views_per_step_per_gpu": V, "views_per_step": n_total,
                       "view_indices": f"{int(rig.idx[0])}, {int(rig.idx[1])}, ... {int(rig.idx[-1])} of {rig.cfg['pose_set']} (evenly spread; view g on GPU g mod N)",
                       "mode": "rayTraceAndClassify + occupied/free update (carve), one launch" if fusion else "rayTraceAndGetPoints -> visibility rows",
                       "grid_format": args.grid, "n_occupied": n_occ, "l2": "flushed between timed iterations" if flush is not None else "not flushed",
                       "parallelism": f"C-ABI group of {world} GPU(s), volume replicated GPU to GPU, exchange: {rig.info['exchange_name']}", "host": rig.numa},
            "voxel_updates_per_s": inbounds_total / (dev_ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (hot * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / (hot * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "k_forward_line<CLASSIFY, carve>" if fusion else "k_forward_line<POINTS>", "kernel_ms": hot, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "SURVEY 8(d) reference-equivalent bytes; not an HBM utilisation (see the config 2 line and DESIGN.md section 5)"},
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": V * 48, "d2h_bytes_per_step": V * (4 + (0 if fusion else vw * 8)), "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "bytes_are": "per rank", "result": ("found flag per view every step; after the last step the marks and observed grids are fused over the GPUs and Voxel::view / Voxel::good (5 B per occupied voxel) "
                                                         "and the observed counts come back, inside the timed region") if fusion else "visibility row + found flag per view, own rows to pinned host memory every step"},
            "parity_in_run": parity, "volume_prepare_ms": rig.prepare, "gpu_launches": int(launches_total),
            "probes": {"reference_equivalent_per_step": samples_total / args.steps, "in_bounds_per_step": inbounds_total / args.steps, "skipped_as_provably_empty_per_step": skipped_total / args.steps},
            "clocks": clocks,
        }
        line.update(extra)
        full_set = 10000 if fusion else 4096
        line["extrapolated_s_for_the_full_pose_set"] = full_set / (V * world) * (dev_ms / args.steps) * 1e-3
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, r, cvol = _cpu_one_view(B, args, rig, 2 if fusion else 0)
                if fusion:
                    ctx.clear_observed(); check(lib.dmf_clear_marks(ctx.h))
                    rig.eng.forward_views(rig.vol, rig.all_poses[:1], D.MODE_CLASSIFY, sc.zdelta, False, view_id0=1, want=())
                    v1, g1 = _marks(rig)
                    rv, rg = cvol.marks()
                    cb["marks_equal_gpu"] = bool(np.array_equal(rv, v1) and np.array_equal(rg, g1))
                    line["parity_in_run"]["reference_marks_equal_gpu"] = cb["marks_equal_gpu"]
                else:
                    ids = rig.eng.rayTraceAndGetPoints(rig.vol, rig.all_poses[0], sc.zdelta, False)[1]
                    cb["ids_equal_gpu"] = bool(np.array_equal(ids, r["ids"]))
                    line["parity_in_run"]["reference_ids_equal_gpu"] = cb["ids_equal_gpu"]
                line["cpu_baseline"] = cb
            except Exception as e:                                        # noqa: BLE001 - the line must not be lost
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    for p in (p0, p1, p2):
        lib.dmf_host_free(p)
    rig.close()
