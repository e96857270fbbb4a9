"""Ad-hoc timing of the forward march through the host-buffer C ABI (kernel ms from CUDA events)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D

name = sys.argv[1] if len(sys.argv) > 1 else "S512"
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sc = D.scenes.scene(name)
ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
L = float(sc.bounds[1])
poses = D.scenes.bench_poses(L, nv)
for fmt in (D.GRID_BIT, D.GRID_BYTE):
    eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, fmt)
    for it in range(3):
        ctx.reset_counters()
        t = time.time(); r = eng.forward_views(gv, poses, 0, sc.zdelta, False, want=("depth", "visibility")); wall = time.time() - t
        ms = ctx.last_kernel_ms(); c = ctx.counters()
        print(f"{name} fmt={fmt} views={nv} kernel_ms={ms:.3f} wall_ms={wall*1e3:.1f} rays/s={nv*480*640/ms*1e3:.3e} samples/s={c['samples']/ms*1e3:.3e} inb/s={c['inbounds']/ms*1e3:.3e} exact={c['exact_div']} hits={c['hits']}")
