"""How much of k_forward_dist's time is the march and how much the per-ray outputs (16-view chunks, hot-kernel events)."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, dmf_b200 as D
sc = D.scenes.scene("S512"); ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
poses = D.scenes.bench_poses(1.0, 128)
for want in ((), ("visibility",), ("depth",), ("depth", "visibility"), ("depth", "points", "voxel", "visibility")):
    ms = []
    for _ in range(4):
        eng.forward_views(gv, poses, 0, 2, False, want=want)
        ms.append(ctx.last_hot_kernel_ms())
    print(f"want={want}: march kernel of the last 16-view chunk {min(ms)*1e3:.1f} us  ({min(ms)/16*1e3:.2f} us/view)")
