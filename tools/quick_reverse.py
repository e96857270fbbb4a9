"""Ad-hoc timing of reverseRayTraceFast batches (kernel ms from CUDA events).  (The CPU side is timed by bench.py only.)"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D
name = sys.argv[1] if len(sys.argv) > 1 else "S512"
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 64
sc = D.scenes.scene(name)
ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
L = float(sc.bounds[1])
poses = D.scenes.bench_poses(L, nv)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx)
for fmt in (D.GRID_BYTE, D.GRID_BIT):
    ctx.set_reverse_format(fmt)
    for it in range(3):
        ctx.reset_counters()
        t = time.time(); r = eng.reverse_views(gv, poses, fast=True, want=("visibility",)); wall = time.time() - t
        ms = ctx.last_hot_kernel_ms(); c = ctx.counters()
        print(f"{name} reverse fmt={fmt} views={nv} n_occ={len(gv.occupied_cells_)} hot_ms={ms:.3f} wall_ms={wall*1e3:.1f} voxel-rays/s={nv*len(gv.occupied_cells_)/ms*1e3:.3e} steps/s={c['samples']/ms*1e3:.3e} skipped={c['skipped']/max(c['samples'],1):.3f} visible={c['hits']}")
