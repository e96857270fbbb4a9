"""March kernel time per view vs grid size (same camera, same pose set): isolates memory-hierarchy effects."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, dmf_b200 as D
ctx = D.Context.default(0)
for name in ("S128d", "S256", "S512", "S1024"):
    sc = D.scenes.scene(name)
    gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    poses = D.scenes.bench_poses(1.0, 128)
    for fmt in (D.GRID_BYTE, D.GRID_BIT):
        eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, fmt)
        ms = []
        for _ in range(3):
            ctx.reset_counters()
            eng.forward_views(gv, poses, 0, sc.zdelta, False, want=())
            ms.append(ctx.last_hot_kernel_ms())
        c = ctx.counters()
        ev = (c["inbounds"] - c["skipped"]) / (128 * 480 * 640)
        print(f"{name} fmt={fmt} zdelta={sc.zdelta}: {min(ms)/128*1e3:.2f} us/view, exact in-bounds probes per ray {ev:.2f}, samples/ray {c['samples']/(128*480*640):.1f}")
