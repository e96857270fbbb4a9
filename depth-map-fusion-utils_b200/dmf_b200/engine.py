"""Host-side mirror of the reference's Camera / VoxelVolume / RayTracingEngine interface over the C ABI.

Same names, argument meaning and defaults as include/Camera.hpp:17-86, include/Volume.hpp:50-78 and
include/RayTracingEngine.hpp:27-40, so a parity test reads like a reference driver:

    cam = Camera(K)                                  # Camera cam(K);
    volume = VoxelVolume(); volume.setDimensions(...); volume.setVolumeSize(...); volume.constructVolume()
    volume.integratePointCloud(cloud, normals)
    engine = RayTracingEngine(cam)
    found, good_points = engine.rayTraceAndGetPoints(volume, T, zdelta, False)

Everything that computes goes through libdmf_b200.so on the GPU; nothing here is a fallback.  Batched
variants (`forward_views`, `reverse_views`) expose what the reference can only do as a Python/C++ loop over views.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import DmfError, ForwardOut, ForwardParams, ReverseOut, check

MODE_POINTS, MODE_GOOD_POINTS, MODE_CLASSIFY, MODE_MARK, MODE_MINIMUM = range(5)
GRID_BIT, GRID_BYTE, GRID_AUTO = 0, 1, 2
FWD_NO_SKIP, FWD_TWO_PROBE, FWD_CARVE, FWD_NO_COUNTERS = 1, 2, 4, 8
NO_VOXEL = np.uint64(0xFFFFFFFFFFFFFFFF)


def _ptr(a: Optional[np.ndarray], ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def _vptr(a: Optional[np.ndarray]):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One per process and GPU: the device state a by-value RayTracingEngine cannot own (SURVEY 8b 'Ownership')."""

    _default: Optional["Context"] = None

    def __init__(self, device: int = 0):
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.dmf_create(C.byref(h), device))
        self.h, self.lib, self.device = h, lib, device
        self._volume_token = None

    @classmethod
    def default(cls, device: int = 0) -> "Context":
        if cls._default is None or cls._default.device != device:
            cls._default = cls(device)
        return cls._default

    def close(self):
        if getattr(self, "h", None) and getattr(self, "_borrowed", False):
            self.h = None                                       # owned by a Comm group (dmf_comm_init_all)
            return
        if getattr(self, "h", None):
            self.lib.dmf_destroy(self.h)
            self.h = None
            if Context._default is self:
                Context._default = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self) -> dict:
        out = np.zeros(len(_lib.COUNTER_NAMES), np.uint64)
        check(self.lib.dmf_counters(self.h, _ptr(out, C.c_uint64)))
        return dict(zip(_lib.COUNTER_NAMES, (int(x) for x in out)))

    def reset_counters(self):
        check(self.lib.dmf_reset_counters(self.h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.dmf_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def last_hot_kernel_ms(self) -> float:
        ms = C.c_float()
        check(self.lib.dmf_last_hot_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def set_reverse_format(self, grid_format: int):
        """GRID_BYTE (default): the reverse march skips provably empty steps via distance bytes; GRID_BIT: every step."""
        check(self.lib.dmf_set_reverse_format(self.h, int(grid_format)))

    def selftest_div1000(self):
        out = np.zeros(5, np.uint64)
        check(self.lib.dmf_selftest_div1000(self.h, _ptr(out, C.c_uint64)))
        return [int(x) for x in out]

    def synchronize(self):
        check(self.lib.dmf_synchronize(self.h))

    # ---- carve mode (DMF_FWD_CARVE): the observed-voxel bit grid -------------------------------------------
    def clear_observed(self):
        check(self.lib.dmf_clear_observed(self.h))

    def observed_words(self) -> np.ndarray:
        """uint32 words over the padded index space (bit index = (x*(dim_y+1) + y)*(dim_z+1) + z)."""
        out = np.zeros(self.lib.dmf_observed_words(self.h), np.uint32)
        check(self.lib.dmf_download_observed(self.h, _ptr(out, C.c_uint32)))
        return out

    def observed_dev_ptr(self) -> int:
        p = C.c_void_p()
        check(self.lib.dmf_observed_dev(self.h, C.byref(p)))
        return int(p.value)

    def observed_counts(self) -> dict:
        out = np.zeros(3, np.uint64)
        check(self.lib.dmf_observed_counts(self.h, _ptr(out, C.c_uint64)))
        return dict(observed=int(out[0]), hit=int(out[1]), free=int(out[2]))


class Camera:
    """Camera(K, height=480, width=640) -- include/Camera.hpp:23."""

    def __init__(self, K, height: int = 480, width: int = 640):
        self.K_ = np.ascontiguousarray(K, np.float32).reshape(9)
        self.height_, self.width_ = int(height), int(width)

    def getHeight(self):
        return self.height_

    def getWidth(self):
        return self.width_

    def validPixel(self, r, c):
        return 0 <= r < self.height_ and 0 <= c < self.width_


class VoxelVolume:
    """VoxelVolume -- include/Volume.hpp:50-78.  The occupancy itself lives in HBM (bit bricks + rank directory)."""

    def __init__(self, ctx: Optional[Context] = None, integrate_on_gpu: bool = False):
        self.ctx = ctx
        self.integrate_on_gpu = integrate_on_gpu     # True: dmf_volume_from_points_gpu (K0) instead of the host builder
        self.occupied_cells_ = np.zeros(0, np.uint64)
        self._pts = []
        self._nrm = []
        self._has_normals = None
        self._dirty = True
        self._bounds = None
        self._req_dims = None

    @classmethod
    def attach(cls, ctx: "Context") -> "VoxelVolume":
        """A host-side handle on the volume that is ALREADY on ctx's device (replicated from a peer GPU by
        Comm.replicate_volume, or uploaded through the C ABI directly): nothing is uploaded again."""
        v = cls(ctx)
        dims = np.zeros(3, np.int32)
        deltas = np.zeros(3, np.float64)
        vs, no, nn = C.c_double(), C.c_size_t(), C.c_size_t()
        check(ctx.lib.dmf_volume_info(ctx.h, _ptr(dims, C.c_int), _ptr(deltas, C.c_double), C.byref(vs), C.byref(no), C.byref(nn)))
        v.xdim_, v.ydim_, v.zdim_ = (int(d) for d in dims)
        v.xdelta_, v.ydelta_, v.zdelta_ = (float(d) for d in deltas)
        v.voxel_size_, v.n_normals_ = vs.value, nn.value
        v.occupied_cells_ = np.zeros(no.value, np.uint64)
        if no.value:
            check(ctx.lib.dmf_volume_get_occupied(ctx.h, _ptr(v.occupied_cells_, C.c_uint64)))
        v._dirty = False
        ctx._volume_token = v
        return v

    # -- Volume.hpp:89-128
    def setDimensions(self, xmin, xmax, ymin, ymax, zmin, zmax):
        self._bounds = np.array([xmin, xmax, ymin, ymax, zmin, zmax], np.float64)
        self.xmin_, self.xmax_, self.ymin_, self.ymax_, self.zmin_, self.zmax_ = (float(v) for v in self._bounds)
        self.xcenter_ = self.xmin_ + (self.xmax_ - self.xmin_) / 2.0
        self.ycenter_ = self.ymin_ + (self.ymax_ - self.ymin_) / 2.0
        self.zcenter_ = self.zmin_ + (self.zmax_ - self.zmin_) / 2.0

    def setVolumeSize(self, xdim, ydim, zdim):
        self._req_dims = np.array([xdim, ydim, zdim], np.int32)

    def constructVolume(self):
        self._pts, self._nrm, self._has_normals, self._dirty = [], [], None, True
        return True

    # -- Volume.hpp:172-228
    def integratePointCloud(self, cloud, normals=None):
        cloud = np.ascontiguousarray(cloud, np.float32).reshape(-1, 3)
        has = normals is not None
        if self._has_normals is not None and self._has_normals != has:
            raise DmfError("mixing integratePointCloud overloads with and without normals is not supported")
        self._has_normals = has
        self._pts.append(cloud)
        if has:
            self._nrm.append(np.ascontiguousarray(normals, np.float32).reshape(-1, 3))
        self._dirty = True
        return True

    def _commit(self, ctx: Context):
        """Build + upload (dmf_volume_from_points) if anything changed since the last engine call."""
        if self.ctx is None:
            self.ctx = ctx
        if not self._dirty and ctx._volume_token is self:
            return
        if self._bounds is None or self._req_dims is None:
            raise DmfError("VoxelVolume: setDimensions/setVolumeSize/constructVolume must be called first")
        pts = np.concatenate(self._pts) if self._pts else np.zeros((0, 3), np.float32)
        nrm = np.concatenate(self._nrm) if self._nrm else None
        build = ctx.lib.dmf_volume_from_points_gpu if self.integrate_on_gpu else ctx.lib.dmf_volume_from_points
        check(build(ctx.h, _ptr(self._bounds, C.c_double), _ptr(self._req_dims, C.c_int), _ptr(pts, C.c_float), _ptr(nrm, C.c_float), len(pts)))
        dims = np.zeros(3, np.int32)
        deltas = np.zeros(3, np.float64)
        vs, no, nn = C.c_double(), C.c_size_t(), C.c_size_t()
        check(ctx.lib.dmf_volume_info(ctx.h, _ptr(dims, C.c_int), _ptr(deltas, C.c_double), C.byref(vs), C.byref(no), C.byref(nn)))
        self.xdim_, self.ydim_, self.zdim_ = (int(d) for d in dims)
        self.xdelta_, self.ydelta_, self.zdelta_ = (float(d) for d in deltas)
        self.voxel_size_ = vs.value
        self.n_normals_ = nn.value
        self.occupied_cells_ = np.zeros(no.value, np.uint64)
        if no.value:
            check(ctx.lib.dmf_volume_get_occupied(ctx.h, _ptr(self.occupied_cells_, C.c_uint64)))
        self._dirty = False
        ctx._volume_token = self

    def normals_csr(self):
        off = np.zeros(len(self.occupied_cells_) + 1, np.uint32)
        nrm = np.zeros((self.n_normals_, 3), np.float32)
        check(self.ctx.lib.dmf_volume_get_normals(self.ctx.h, _ptr(off, C.c_uint32), _ptr(nrm, C.c_float)))
        return off, nrm

    def marks(self):
        """(view, good) of every occupied voxel, in occupied_cells_ order (Voxel::view / Voxel::good)."""
        n = len(self.occupied_cells_)
        view, good = np.zeros(n, np.int32), np.zeros(n, np.uint8)
        check(self.ctx.lib.dmf_download_marks(self.ctx.h, _ptr(view, C.c_int32), _ptr(good, C.c_uint8), n))
        return view, good

    def clear_marks(self):
        check(self.ctx.lib.dmf_clear_marks(self.ctx.h))

    # -- Volume.hpp:143-170 helpers (pure index arithmetic, kept for drivers that call them)
    @staticmethod
    def getHashId(x, y, z):
        return (int(x) << 40) ^ (int(y) << 20) ^ int(z)

    @staticmethod
    def getVoxelCoords(id_):
        id_ = int(id_)
        return id_ >> 40, (id_ >> 20) & ((1 << 20) - 1), id_ & ((1 << 20) - 1)


def _poses12(T) -> np.ndarray:
    a = np.ascontiguousarray(T, np.float32)
    if a.shape[-2:] == (4, 4):
        a = a[..., :3, :]
    return np.ascontiguousarray(a.reshape(-1, 12))


class RayTracingEngine:
    """RayTracingEngine(cam) -- include/RayTracingEngine.hpp:27-42.  Public member cam_ as in the reference."""

    def __init__(self, cam: Camera, ctx: Optional[Context] = None, grid_format: int = GRID_BIT, skip_empty: bool = True):
        self.cam_ = cam
        self.ctx = ctx or Context.default()
        self.grid_format = grid_format
        self.skip_empty = skip_empty   # False: DMF_FWD_NO_SKIP (evaluate every probe)

    def _prepare(self, volume: VoxelVolume):
        volume._commit(self.ctx)
        check(self.ctx.lib.dmf_set_camera(self.ctx.h, _ptr(self.cam_.K_, C.c_float), self.cam_.height_, self.cam_.width_))

    # ---- batched forward: n views in one call ---------------------------------------------------------------
    def forward_views(self, volume: VoxelVolume, poses, mode: int, zdelta: int, sparse: bool, view_id0: int = 1,
                      want=("depth", "points", "voxel", "visibility", "ids"), carve: bool = False) -> dict:
        """carve=True (DMF_FWD_CARVE): every visited in-bounds sample also marks its voxel in the context's observed grid
        (Context.observed_words / observed_counts)."""
        self._prepare(volume)
        poses = _poses12(poses)
        n = len(poses)
        H, W = self.cam_.height_, self.cam_.width_
        vw = self.ctx.lib.dmf_visibility_words(self.ctx.h)
        res = {}
        o = ForwardOut()
        if "depth" in want:
            res["depth"] = np.empty((n, H, W), np.int32); o.depth_mm = _vptr(res["depth"])
        if "depth16" in want:
            res["depth16"] = np.empty((n, H, W), np.uint16); o.depth_u16 = _vptr(res["depth16"])
        if "points" in want:
            res["points"] = np.empty((n, H, W, 3), np.float32); o.points = _vptr(res["points"])
        if "voxel" in want:
            res["voxel"] = np.empty((n, H, W), np.uint64); o.hit_voxel = _vptr(res["voxel"])
        if "visibility" in want and mode != MODE_MINIMUM:
            res["visibility"] = np.zeros((n, vw), np.uint64); o.visibility = _vptr(res["visibility"])
        res["found_any"] = np.zeros(n, np.int32); o.found_any = _vptr(res["found_any"])
        if mode == MODE_MINIMUM:
            res["min_depth"] = np.zeros(n, np.int32); o.min_depth = _vptr(res["min_depth"])
        ids = offs = None
        if "ids" in want and mode in (MODE_POINTS, MODE_GOOD_POINTS):
            cap = n * min(len(volume.occupied_cells_), H * W) + 1
            ids, offs = np.empty(cap, np.uint64), np.zeros(n + 1, np.int64)
            o.ids, o.ids_offsets, o.ids_capacity = _vptr(ids), _vptr(offs), cap
        p = ForwardParams(mode, int(zdelta), int(bool(sparse)), int(view_id0), self.grid_format, (0 if self.skip_empty else FWD_NO_SKIP) | (FWD_CARVE if carve else 0))
        check(self.ctx.lib.dmf_forward(self.ctx.h, C.byref(p), _ptr(poses, C.c_float), n, C.byref(o)))
        if ids is not None:
            res["ids"] = [ids[offs[i]:offs[i + 1]].copy() for i in range(n)]
        return res

    # ---- the reference's five forward methods (:229-494), same argument order and defaults ---------------------
    def rayTraceAndGetPoints(self, volume, transformation, zdelta=10, sparse=True):
        r = self.forward_views(volume, transformation, MODE_POINTS, zdelta, sparse, want=("ids",))
        return bool(r["found_any"][0]), r["ids"][0]

    def rayTraceAndGetGoodPoints(self, volume, transformation, zdelta=10, sparse=True):
        r = self.forward_views(volume, transformation, MODE_GOOD_POINTS, zdelta, sparse, want=("ids",))
        return bool(r["found_any"][0]), r["ids"][0]

    def rayTraceAndClassify(self, volume, transformation, zdelta=10, view=1, sparse=True):
        self.forward_views(volume, transformation, MODE_CLASSIFY, zdelta, sparse, view_id0=view, want=())

    def rayTrace(self, volume, transformation, zdelta=10, sparse=True):
        self.forward_views(volume, transformation, MODE_MARK, zdelta, sparse, want=())

    def rayTraceAndGetMinimum(self, volume, transformation, zdelta=1, sparse=True):
        return int(self.forward_views(volume, transformation, MODE_MINIMUM, zdelta, sparse, want=())["min_depth"][0])

    # ---- reverse (:45-226) -------------------------------------------------------------------------------------
    def reverse_views(self, volume: VoxelVolume, poses, fast: bool = True, viz: bool = False, want=("visibility", "unoccluded", "ids")) -> dict:
        self._prepare(volume)
        poses = _poses12(poses)
        n = len(poses)
        vw = self.ctx.lib.dmf_visibility_words(self.ctx.h)
        res = {}
        o = ReverseOut()
        if "visibility" in want:
            res["visibility"] = np.zeros((n, vw), np.uint64); o.visibility = _vptr(res["visibility"])
        if "unoccluded" in want:
            res["unoccluded"] = np.zeros((n, vw), np.uint64); o.unoccluded = _vptr(res["unoccluded"])
        res["found_any"] = np.zeros(n, np.int32); o.found_any = _vptr(res["found_any"])
        ids = offs = None
        if "ids" in want:
            cap = n * (2 * len(volume.occupied_cells_) + 64) + 1
            ids, offs = np.empty(cap, np.uint64), np.zeros(n + 1, np.int64)
            o.ids, o.ids_offsets, o.ids_capacity = _vptr(ids), _vptr(offs), cap
        check(self.ctx.lib.dmf_reverse(self.ctx.h, int(bool(fast)), int(bool(viz)), _ptr(poses, C.c_float), n, C.byref(o)))
        if ids is not None:
            res["ids"] = [ids[offs[i]:offs[i + 1]].copy() for i in range(n)]
        return res

    def reverseRayTraceFast(self, volume, transformation, viz, zdelta=1):
        r = self.reverse_views(volume, transformation, True, viz, want=("ids",))
        return bool(r["found_any"][0]), r["ids"][0]

    def reverseRayTrace(self, volume, transformation, viz, zdelta=1):
        r = self.reverse_views(volume, transformation, False, viz, want=("ids",))
        return bool(r["found_any"][0]), r["ids"][0]

    # ---- z-buffer (:498-564) -----------------------------------------------------------------------------------
    def rayTraceVolume(self, volume, transformation, return_depth: bool = False):
        self._prepare(volume)
        pose = _poses12(transformation)[0]
        depth = np.zeros((self.cam_.height_, self.cam_.width_), np.int32)
        n = C.c_int64()
        check(self.ctx.lib.dmf_zbuffer(self.ctx.h, _ptr(pose, C.c_float), _ptr(depth, C.c_int32), C.byref(n)))
        if return_depth:
            return depth, n.value


def willCollide(engine_or_ctx, volume: VoxelVolume, a, b, guard_coords: bool = True) -> np.ndarray:
    """willCollide(volume, a, b) of the reference drivers (tests/CameraPathGen.cpp:128-156) for a batch of segments:
    a, b are (n,3).  guard_coords=False reproduces the copies without the validCoords guard."""
    ctx = engine_or_ctx.ctx if isinstance(engine_or_ctx, RayTracingEngine) else engine_or_ctx
    volume._commit(ctx)
    a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
    b = np.ascontiguousarray(b, np.float32).reshape(-1, 3)
    out = np.zeros(len(a), np.uint8)
    check(ctx.lib.dmf_segments_collide(ctx.h, _ptr(a, C.c_float), _ptr(b, C.c_float), len(a), int(guard_coords), _ptr(out, C.c_uint8)))
    return out.astype(bool)


def optimizeCameraPosition(volume: VoxelVolume, engine: "RayTracingEngine", cameras, low: int = 300, high: int = 600):
    """Algorithms::optimizeCameraPosition(volume, engine, res, Affine3f camera) (Algorithms.hpp:394-421) for a batch of
    cameras.  Returns (mid[n], repositioned poses [n,12])."""
    engine._prepare(volume)
    poses = _poses12(cameras)
    mid = np.zeros(len(poses), np.uint32)
    out = np.zeros_like(poses)
    check(engine.ctx.lib.dmf_optimize_standoff(engine.ctx.h, _ptr(poses, C.c_float), len(poses), low, high, _ptr(mid, C.c_uint32), _ptr(out, C.c_float)))
    return mid, out


def reposition_from_minimum(cameras, nearest_mm) -> np.ndarray:
    """The arithmetic of repositionCamerasSampled (tests/CameraPathGen.cpp:113-121), float for float: with bz = the camera's
    z column and p = its translation, intersection = p + bz * float(double(float(nearest)) / 1000.0), new p = intersection -
    bz * float(0.3); cameras whose cast found nothing (nearest == -1) are returned unchanged (:108-113)."""
    f32 = np.float32
    T = _poses12(cameras).reshape(-1, 3, 4).copy()
    near = np.asarray(nearest_mm, np.int64).reshape(-1)
    s = (near.astype(f32).astype(np.float64) / 1000.0).astype(f32)          # Vector3f * double narrows the scalar to float
    bz, p = T[:, :, 2], T[:, :, 3]
    inter = p + bz * s[:, None]
    new_p = inter - bz * f32(0.3)
    hit = near != -1
    T[hit, :, 3] = new_p[hit]
    return T.reshape(-1, 12)


def repositionCamerasSampled(cameras, volume: VoxelVolume, engine: "RayTracingEngine") -> np.ndarray:
    """repositionCamerasSampled(cameras, volume, cam) of the reference driver (tests/CameraPathGen.cpp:94-126) for a batch:
    one rayTraceAndGetMinimum cast per camera (zdelta = 1, sparse: its defaults), then every camera that hit something is
    moved to 0.3 m in front of the nearest hit along its optical axis.  Returns (n, 12) poses."""
    poses = _poses12(cameras)
    r = engine.forward_views(volume, poses, MODE_MINIMUM, 1, True, want=())
    return reposition_from_minimum(poses, r["min_depth"])


def greedySetCover(candidate_bitsets: np.ndarray, ctx: Optional[Context] = None) -> np.ndarray:
    """Algorithms::greedySetCover (Algorithms.hpp:38-86) over visibility bitsets [n_sets][words]."""
    ctx = ctx or Context.default()
    b = np.ascontiguousarray(candidate_bitsets, np.uint64)
    n, words = b.shape
    sel = np.zeros(max(n, 1), np.int32)
    cnt = C.c_int()
    check(ctx.lib.dmf_greedy_set_cover(ctx.h, _ptr(b, C.c_uint64), n, words, _ptr(sel, C.c_int32), C.byref(cnt)))
    return sel[: cnt.value].copy()


def bits_to_indices(bits: np.ndarray) -> np.ndarray:
    """Indices of the set bits of one uint64 bitset row (little-endian bit order: bit i of word w = element 64w+i)."""
    return np.nonzero(np.unpackbits(bits.view(np.uint8), bitorder="little"))[0]
