"""One-GPU throughput on BASELINE.json configs[1..4] (bounded samples of the big ones; the sample is stated per row).

    python tools/config_table.py > gpurun_out/config_table.json

Every row goes through the host-buffer calls of the Python mirror (poses from host memory, results back to host memory),
so `wall_ms` is end to end; `kernel_ms` is the march kernel(s) alone (CUDA events inside the library).  Parity at these
sizes is the business of tests/ (test_config1_*, test_config3_*, test_config4_*, test_full_size_properties_*)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np

import dmf_b200 as D

ctx = D.Context.default(0)
K = D.scenes.REFERENCE_K
rows = []


def volume(name):
    sc = D.scenes.scene(name)
    gv = D.VoxelVolume(ctx)
    gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    return sc, gv


def timed(fn, reps=3):
    fn()                                    # warm-up: uploads, tables, distance bytes, scratch
    best, hot = 1e30, 0.0
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        if dt < best:
            best, hot = dt, ctx.last_hot_kernel_ms()
    return 1e3 * best, hot


def row(config, what, n_views, H, W, wall_ms, kernel_ms, extra=None):
    r = {"config": config, "what": what, "views": n_views, "camera": f"{W}x{H}", "wall_ms": wall_ms, "kernel_ms": kernel_ms,
         "rays_per_s_e2e": n_views * H * W / (wall_ms * 1e-3), "rays_per_s_kernel": n_views * H * W / (kernel_ms * 1e-3) if kernel_ms else None}
    r.update(extra or {})
    rows.append(r)
    print(json.dumps(r), file=sys.stderr, flush=True)


# ---- configs[1]: one 640x480 view into the 512^3 grid: simulated depth cloud + occupied/free update -----------------------
sc, gv = volume("S512")
eng = D.RayTracingEngine(D.Camera(K, 480, 640), ctx, D.GRID_BYTE)
p1 = D.scenes.poses_sphere_lookat(1.0, 64)[37:38]
w, k = timed(lambda: eng.forward_views(gv, p1, D.MODE_POINTS, sc.zdelta, False, want=("depth", "points", "visibility")))
row(1, "one view: depth image + simulated point cloud + visibility to the host", 1, 480, 640, w, k)
w, k = timed(lambda: eng.forward_views(gv, p1, D.MODE_CLASSIFY, sc.zdelta, False, want=(), carve=True))
ctx.reset_counters(); eng.forward_views(gv, p1, D.MODE_CLASSIFY, sc.zdelta, False, want=(), carve=True)
inb = ctx.counters()["inbounds"]
row(1, "one view: rayTraceAndClassify marks + occupied/free update (carve)", 1, 480, 640, w, k, {"voxel_updates": inb, "voxel_updates_per_s_kernel": inb / (k * 1e-3)})

# ---- configs[2]: the 1024-view candidate sweep -> per-view visibility bitsets -> greedy set cover ------------------------
poses = D.scenes.bench_poses(1.0, 1024)
res = {}
def sweep_forward():
    res["f"] = eng.forward_views(gv, poses, D.MODE_POINTS, sc.zdelta, False, want=("visibility",))
w, k = timed(sweep_forward)
row(2, "1024-view forward sweep, visibility bitsets to the host", 1024, 480, 640, w, k)
def sweep_reverse():
    res["r"] = eng.reverse_views(gv, poses, fast=True, want=("visibility",))
w, k = timed(sweep_reverse)
row(2, "1024-view reverseRayTraceFast sweep (what tests/SetCover.cpp runs), visibility bitsets to the host", 1024, 480, 640, w, k,
    {"voxel_rays_per_s_e2e": 1024 * len(gv.occupied_cells_) / (w * 1e-3)})
t0 = time.perf_counter(); sel = D.greedySetCover(res["r"]["visibility"], ctx); dt = time.perf_counter() - t0
rows.append({"config": 2, "what": "greedy set cover over the 1024 reverse bitsets (host bitsets in, selection out)", "wall_ms": 1e3 * dt, "selected_views": int(len(sel))})

# ---- configs[3]: moving-camera fusion, 1024^3 grid, bounded sample of the 10k-pose helix ---------------------------------
sc3, gv3 = volume("S1024")
eng3 = D.RayTracingEngine(D.Camera(K, 480, 640), ctx, D.GRID_BYTE)
helix = np.ascontiguousarray(D.scenes.poses_helix(1.0, 10000)[::40])              # 250 of the 10 000 poses
w, k = timed(lambda: eng3.forward_views(gv3, helix, D.MODE_CLASSIFY, sc3.zdelta, False, want=()), reps=2)
row(3, f"{len(helix)} of the 10k helix poses (every 40th): rayTraceAndClassify marks (Voxel::view = first view, Voxel::good)", len(helix), 480, 640, w, k,
    {"extrapolated_s_for_10k_poses": 10000 / len(helix) * w * 1e-3})
gv3._commit(ctx); ctx.clear_observed()
w, k = timed(lambda: eng3.forward_views(gv3, helix, D.MODE_CLASSIFY, sc3.zdelta, False, want=(), carve=True), reps=2)
ctx.reset_counters(); eng3.forward_views(gv3, helix, D.MODE_CLASSIFY, sc3.zdelta, False, want=(), carve=True)
inb = ctx.counters()["inbounds"]
oc = ctx.observed_counts()
row(3, f"the same {len(helix)} poses with the occupied/free update (carve) as well", len(helix), 480, 640, w, k,
    {"voxel_updates": inb, "voxel_updates_per_s_e2e": inb / (w * 1e-3), "voxel_updates_per_s_kernel": inb / (k * 1e-3),
     "observed_voxels": oc["observed"], "free_voxels": oc["free"], "extrapolated_s_for_10k_poses": 10000 / len(helix) * w * 1e-3})

# ---- configs[4]: 1920x1080 camera into the 1024^3 grid, bounded sample of the 4096-view Fibonacci sweep ------------------
K3 = D.scenes.scaled_K(3.0)
eng4 = D.RayTracingEngine(D.Camera(K3, 1080, 1920), ctx, D.GRID_BYTE)
fib = np.ascontiguousarray(D.scenes.poses_fibonacci(1.0, 4096)[::64])               # 64 of the 4096 views
w, k = timed(lambda: eng4.forward_views(gv3, fib, D.MODE_POINTS, sc3.zdelta, False, want=("visibility",)), reps=2)
row(4, f"{len(fib)} of the 4096 Fibonacci views (every 64th), 1080p: visibility bitsets to the host", len(fib), 1080, 1920, w, k,
    {"extrapolated_s_for_4096_views": 4096 / len(fib) * w * 1e-3})
w, k = timed(lambda: eng4.forward_views(gv3, fib, D.MODE_POINTS, sc3.zdelta, False, want=("depth16", "visibility")), reps=2)
row(4, "the same with the uint16 depth maps copied back too", len(fib), 1080, 1920, w, k, {"d2h_bytes": int(len(fib)) * 1080 * 1920 * 2})

print(json.dumps({"gpu": "B200", "rows": rows}, indent=1))
