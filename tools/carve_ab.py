"""A/B of carve-mode kernel builds: timing on the bench workload plus a checksum of the observed grid, so that a variant
(DMF_B200_LIB=path/to/variant.so) can be held against the default build at full size.  No torch: the library's own CUDA-event
timer (dmf_last_kernel_ms) brackets the march launches.
usage: [DMF_B200_LIB=...] python tools/carve_ab.py [SCENE=S512] [VIEWS=128] [STEPS=5]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D

name = sys.argv[1] if len(sys.argv) > 1 else "S512"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc = D.scenes.scene(name)
ctx = D.Context.default(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
poses = D.scenes.bench_poses(float(sc.bounds[1]), V)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
tag = os.path.basename(D.LIB_PATH)


def run(p, want=()):
    eng.forward_views(gv, p, D.MODE_POINTS, sc.zdelta, False, want=want, carve=True)
    return ctx.last_kernel_ms()


run(poses[:2]); ctx.clear_observed()                      # warm-up (allocations, tables)
one = run(poses[:1])                                      # configs[1]: one view into a cleared grid
h1 = hashlib.sha1(ctx.observed_words().tobytes()).hexdigest()[:16]
ctx.clear_observed(); ctx.reset_counters()
first = run(poses)
c = ctx.counters()
steady = sorted(run(poses) for _ in range(steps))
words = ctx.observed_words()
print(f"{tag}: {name} V={V} one_view_first_pass={one:.4f} ms  first_pass={first:.3f} ms  steady median={steady[len(steady) // 2]:.3f} min={steady[0]:.3f} ms  "
      f"updates/s={c['inbounds'] / steady[len(steady) // 2] * 1e3:.3e}  inbounds={c['inbounds']} observed={ctx.observed_counts()}  "
      f"sha1(one view)={h1} sha1({V} views)={hashlib.sha1(words.tobytes()).hexdigest()[:16]}", flush=True)
