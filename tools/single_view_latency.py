import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, dmf_b200 as D
sc = D.scenes.scene("S512"); ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
poses = np.ascontiguousarray(D.scenes.poses_sphere_lookat(1.0, 1024)[::8][:32]) if len(sys.argv) > 1 and sys.argv[1] == "spread" else D.scenes.bench_poses(1.0, 32)
for want in (("ids",), ("visibility",), ("depth",), ()):
    for _ in range(3): eng.forward_views(gv, poses[:1], 0, 2, False, want=want)
    t = time.perf_counter()
    for i in range(20): eng.forward_views(gv, poses[i:i+1], 0, 2, False, want=want)
    dt = (time.perf_counter() - t) / 20
    print(f"forward 1 view want={want}: {dt*1e3:.3f} ms/call, kernel-span {ctx.last_kernel_ms():.3f} ms, march kernel {ctx.last_hot_kernel_ms():.3f} ms")
for want in (("ids",), ("visibility",)):
    for _ in range(3): eng.reverse_views(gv, poses[:1], want=want)
    t = time.perf_counter()
    for i in range(20): eng.reverse_views(gv, poses[i:i+1], want=want)
    dt = (time.perf_counter() - t) / 20
    print(f"reverse 1 view want={want}: {dt*1e3:.3f} ms/call, kernel-span {ctx.last_kernel_ms():.3f} ms, march kernel {ctx.last_hot_kernel_ms():.3f} ms")
