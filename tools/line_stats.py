"""Diagnostic (build with DMF_NVCC_EXTRA=-DDMF_LINE_STATS): probes per ray of k_forward_line by kind, and the per-SM
block counts / durations it logs (stderr of dmf_counters)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D

name = sys.argv[1] if len(sys.argv) > 1 else "S512"
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 128
want = tuple(w for w in (sys.argv[3] if len(sys.argv) > 3 else "depth").split(",") if w)
sc = D.scenes.scene(name)
ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
poses = np.ascontiguousarray(D.scenes.poses_sphere_lookat(1.0, 1024)[::1024 // nv]) if 1024 % nv == 0 else D.scenes.bench_poses(float(sc.bounds[1]), nv)   # the bench step's view mix
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
eng.forward_views(gv, poses, 0, sc.zdelta, False, want=want)
ctx.reset_counters()
eng.forward_views(gv, poses, 0, sc.zdelta, False, want=want)
print(name, nv, want, "hot ms (last chunk)", ctx.last_hot_kernel_ms(), flush=True)
c = ctx.counters(); rays = nv * 480 * 640
print({k: round(v / rays, 2) for k, v in c.items() if k in ("samples", "inbounds", "hits", "skipped", "exact_div", "f64_path")})
