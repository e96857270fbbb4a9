import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, dmf_b200 as D
sc = D.scenes.scene("S512"); ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
poses = D.scenes.poses_sphere_lookat(1.0, 1024)[::8]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for i in range(3):
    eng.rayTraceAndGetPoints(gv, poses[i], 2, False); eng.reverseRayTraceFast(gv, poses[i], False)
print("MARK forward", flush=True)
for i in range(n): eng.rayTraceAndGetPoints(gv, poses[i], 2, False)
print("MARK reverse", flush=True)
for i in range(n): eng.reverseRayTraceFast(gv, poses[i], False)
