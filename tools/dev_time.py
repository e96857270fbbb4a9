"""Device-resident timing of dmf_forward_dev (the march kernel + its bookkeeping) for any scene / output set.
usage: dev_time.py SCENE VIEWS OUTPUTS [FLAGS] [STEPS]   OUTPUTS = letters of d(epth) p(oints) h(it_voxel) v(isibility) f(ound_any), '-' = none"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np, torch
import dmf_b200 as D
from dmf_b200._lib import ForwardParams, ForwardOut

name = sys.argv[1] if len(sys.argv) > 1 else "S512"
V = int(sys.argv[2]) if len(sys.argv) > 2 else 128
outs = sys.argv[3] if len(sys.argv) > 3 else "dphvf"
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
H, W = 480, 640
ctx = D.Context(0)
sc = D.scenes.scene(name)
vol = D.VoxelVolume(ctx); vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K, H, W), ctx, D.GRID_BYTE)
eng._prepare(vol)
vw = (len(vol.occupied_cells_) + 63) // 64
dev = torch.device("cuda", 0)
d_poses = torch.from_numpy(np.ascontiguousarray(D.scenes.bench_poses(float(sc.bounds[1]), V))).to(dev)
o = ForwardOut()
keep = []
if "d" in outs: t = torch.empty((V, H, W), dtype=torch.int32, device=dev); keep.append(t); o.depth_mm = t.data_ptr()
if "p" in outs: t = torch.empty((V, H, W, 3), dtype=torch.float32, device=dev); keep.append(t); o.points = t.data_ptr()
if "h" in outs: t = torch.empty((V, H, W), dtype=torch.int64, device=dev); keep.append(t); o.hit_voxel = t.data_ptr()
if "v" in outs: t = torch.zeros((V, vw), dtype=torch.int64, device=dev); keep.append(t); o.visibility = t.data_ptr()
if "f" in outs: t = torch.zeros((V,), dtype=torch.int32, device=dev); keep.append(t); o.found_any = t.data_ptr()
params = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, D.GRID_BYTE, flags)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
def step():
    rc = ctx.lib.dmf_forward_dev(ctx.h, C.byref(params), C.c_void_p(d_poses.data_ptr()), V, C.byref(o), C.c_void_p(st.cuda_stream))
    assert rc == 0, D.last_error() if hasattr(D, "last_error") else rc
for _ in range(3): step()
torch.cuda.synchronize()
ms = []
for _ in range(steps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
ms.sort()
print(f"{name} V={V} outs={outs} flags={flags}: median {ms[len(ms)//2]:.3f} ms/step  min {ms[0]:.3f}  = {ms[len(ms)//2]/V*1e3:.2f} us/view")
