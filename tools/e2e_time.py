"""Wall-clock timing of the host-buffer dmf_forward call (pinned buffers) for output variants / chunk counts."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-map-fusion-utils_b200"))
import numpy as np
import dmf_b200 as D
from dmf_b200._lib import ForwardParams, ForwardOut
V = int(sys.argv[1]) if len(sys.argv) > 1 else 128
outs = sys.argv[2] if len(sys.argv) > 2 else "uvf"     # d=depth i32, u=depth u16, v=visibility, f=found
H, W = 480, 640
ctx = D.Context(0); lib = ctx.lib
sc = D.scenes.scene("S512")
vol = D.VoxelVolume(ctx); vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
eng = D.RayTracingEngine(D.Camera(D.scenes.REFERENCE_K, H, W), ctx, D.GRID_BYTE); eng._prepare(vol)
vw = (len(vol.occupied_cells_) + 63) // 64
def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    lib.dmf_host_alloc.restype = C.c_void_p
    p = lib.dmf_host_alloc(C.c_size_t(n))
    return np.ctypeslib.as_array((C.c_ubyte * n).from_address(p)).view(dtype).reshape(shape), p
poses, _ = pinned((V, 12), np.float32); poses[:] = D.scenes.bench_poses(1.0, V)
o = ForwardOut(); keep = []
if "d" in outs: a, p = pinned((V, H, W), np.int32); o.depth_mm = p
if "u" in outs: a, p = pinned((V, H, W), np.uint16); o.depth_u16 = p
if "v" in outs: a, p = pinned((V, vw), np.uint64); o.visibility = p
if "f" in outs: a, p = pinned((V,), np.int32); o.found_any = p
params = ForwardParams(D.MODE_POINTS, sc.zdelta, 0, 1, D.GRID_BYTE, 0)
def call():
    rc = lib.dmf_forward(ctx.h, C.byref(params), poses.ctypes.data_as(C.POINTER(C.c_float)), V, C.byref(o)); assert rc == 0
for _ in range(3): call()
ts = []
for _ in range(10):
    t = time.perf_counter(); call(); ts.append((time.perf_counter() - t) * 1e3)
ts.sort()
print(f"V={V} outs={outs} chunks={os.environ.get('DMF_FWD_CHUNKS','default')}: median {ts[5]:.3f} ms  min {ts[0]:.3f} ms   kernel-span {ctx.last_kernel_ms():.3f} ms")
