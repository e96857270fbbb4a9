"""kernel ms of the forward bench step (128 views P1024[::8], S512, all per-pixel outputs) for A/B of build or env variants"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, torch
import dmf_b200 as D
from dmf_b200._lib import ForwardOut, ForwardParams, check
name = sys.argv[1] if len(sys.argv) > 1 else "S512"
sc = D.scenes.scene(name); ctx = D.Context(0)
gv = D.VoxelVolume(ctx); gv.setDimensions(*sc.bounds); gv.setVolumeSize(*sc.dims); gv.constructVolume(); gv.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K), ctx, D.GRID_BYTE); eng._prepare(gv)
H, W, V = 480, 640, 128
poses = np.ascontiguousarray(D.scenes.poses_sphere_lookat(1.0, 1024)[::8])
dev = torch.device("cuda", 0)
d_poses = torch.from_numpy(poses).to(dev)
vw = (len(gv.occupied_cells_) + 63) // 64
bufs = dict(depth=torch.empty((V, H, W), dtype=torch.int32, device=dev), pts=torch.empty((V, H, W, 3), dtype=torch.float32, device=dev),
            vox=torch.empty((V, H, W), dtype=torch.int64, device=dev), vis=torch.zeros((V, vw), dtype=torch.int64, device=dev), found=torch.zeros(V, dtype=torch.int32, device=dev))
o = ForwardOut(); o.depth_mm, o.points, o.hit_voxel, o.visibility, o.found_any = (bufs[k].data_ptr() for k in ("depth", "pts", "vox", "vis", "found"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for flags, tag in ((D.FWD_NO_COUNTERS, "no counters"), (0, "counters"), (D.FWD_NO_COUNTERS | D.FWD_CARVE, "carve")):
    p = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, D.GRID_BYTE, flags)
    ms = []
    for it in range(8):
        flush.fill_(1)
        check(ctx.lib.dmf_forward_dev(ctx.h, C.byref(p), C.c_void_p(d_poses.data_ptr()), V, C.byref(o), C.c_void_p(st.cuda_stream)))
        torch.cuda.synchronize()
        ms.append(ctx.last_hot_kernel_ms())
    import hashlib
    print(f"{name} {tag:12s} kernel_ms best {min(ms[2:]):.3f} median {sorted(ms[2:])[3]:.3f}  depth sha {hashlib.sha1(bufs['depth'].cpu().numpy().tobytes()).hexdigest()[:10]} env L2={os.environ.get('DMF_L2_PERSIST','-')}", flush=True)
