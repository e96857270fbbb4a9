"""kernel ms of the reverse sweep (128 views of the bench step, S512 and S1024): run once with and once without DMF_REVERSE_NO_POOL=1"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D
for name in ("S512", "S1024"):
    sc = D.scenes.scene(name)
    ctx = D.Context(0)
    gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
    poses = np.ascontiguousarray(D.scenes.poses_sphere_lookat(1.0, 1024)[::8])
    eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE)
    ms = []
    for it in range(6):
        ctx.reset_counters()
        r = eng.reverse_views(gv, poses, fast=True, want=("visibility",))
        ms.append(ctx.last_hot_kernel_ms())
    c = ctx.counters()
    import hashlib
    print(f"{name} pool={'no' if os.environ.get('DMF_REVERSE_NO_POOL') else 'yes'} kernel_ms best {min(ms):.3f} median {sorted(ms)[3]:.3f}  samples {c['samples']} inbounds {c['inbounds']} hits {c['hits']} skipped {c['skipped']} sha {hashlib.sha1(r['visibility'].tobytes()).hexdigest()[:12]}", flush=True)
    ctx.close()
