"""ctypes binding of the CPU oracle (oracle/liboracle_dmf.so).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Pinned against the reference's own headers compiled here (oracle/_ref,
tests/test_reference_build_cpu.py); PARITY UNPINNED only for Eigen's internal op order (see dmf_oracle.hpp).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMF_ORACLE_LIB: another build of the same library (the ASan/UBSan build of `make -C oracle asan`); default = in-tree
_LIB = os.environ.get("DMF_ORACLE_LIB") or os.path.join(_HERE, "liboracle_dmf.so")

MODE_POINTS, MODE_GOOD_POINTS, MODE_CLASSIFY, MODE_MARK, MODE_MINIMUM = range(5)


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("dmf_oracle_capi.cpp", "dmf_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src)
    if os.environ.get("DMF_ORACLE_LIB"):
        return _LIB                                  # a prebuilt variant: never rebuilt from here
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
        u64p, llp, ucp = C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.POINTER(C.c_ubyte)
        L.orc_volume_new.restype = vp
        L.orc_volume_new.argtypes = [dp, ip, C.c_int]
        L.orc_volume_free.argtypes = [vp]
        L.orc_volume_info.argtypes = [vp, ip, dp, dp]
        L.orc_volume_integrate.restype = C.c_long
        L.orc_volume_integrate.argtypes = [vp, fp, fp, C.c_long]
        L.orc_volume_num_occupied.restype = C.c_long
        L.orc_volume_num_occupied.argtypes = [vp]
        L.orc_volume_get_occupied.argtypes = [vp, u64p]
        L.orc_volume_num_normals.restype = C.c_long
        L.orc_volume_num_normals.argtypes = [vp]
        L.orc_volume_get_normals.argtypes = [vp, C.POINTER(C.c_uint), fp]
        L.orc_volume_get_marks.argtypes = [vp, ip, ucp]
        L.orc_volume_clear_marks.argtypes = [vp]
        L.orc_forward.restype = C.c_long
        L.orc_forward.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  ip, fp, u64p, u64p, C.c_long, ip, ip, llp]
        L.orc_forward_observed.restype = C.c_long
        L.orc_forward_observed.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint), llp]
        L.orc_reverse.restype = C.c_long
        L.orc_reverse.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int,
                                  u64p, C.c_long, ip, ucp, llp]
        L.orc_zbuffer.argtypes = [vp, fp, C.c_int, C.c_int, fp, ip, llp]
        L.orc_affine_inverse.argtypes = [fp, fp]
        L.orc_degree_acosf.restype = C.c_int
        L.orc_degree_acosf.argtypes = [C.c_float]
        L.orc_greedy_set_cover.restype = C.c_long
        L.orc_greedy_set_cover.argtypes = [u64p, llp, C.c_long, u64p]
        L.orc_will_collide.restype = C.c_int
        L.orc_will_collide.argtypes = [vp, fp, fp, C.c_int, llp]
        L.orc_optimize_standoff.restype = C.c_uint
        L.orc_optimize_standoff.argtypes = [vp, fp, C.c_int, C.c_int, fp, fp]
        L.orc_time_views.restype = C.c_double
        L.orc_time_views.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, llp]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_eigen_order.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def set_eigen_order(o: int):
    lib().orc_set_eigen_order(int(o))


class Volume:
    """VoxelVolume of the oracle: setDimensions + setVolumeSize + constructVolume (+ integratePointCloud)."""

    def __init__(self, bounds, dims, flat=True):
        b = np.ascontiguousarray(bounds, np.float64)
        d = np.ascontiguousarray(dims, np.int32)
        self.h = lib().orc_volume_new(_p(b, C.c_double), _p(d, C.c_int), int(flat))
        self.bounds = b
        od = np.zeros(3, np.int32)
        dl = np.zeros(3, np.float64)
        vs = C.c_double()
        lib().orc_volume_info(self.h, _p(od, C.c_int), _p(dl, C.c_double), C.byref(vs))
        self.dims, self.deltas, self.voxel_size = od, dl, vs.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_volume_free(self.h)
            self.h = None

    def integrate(self, pts, normals=None) -> int:
        pts = np.ascontiguousarray(pts, np.float32)
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        return lib().orc_volume_integrate(self.h, _p(pts, C.c_float), _p(nrm, C.c_float), len(pts))

    @property
    def n_occupied(self) -> int:
        return lib().orc_volume_num_occupied(self.h)

    def occupied(self) -> np.ndarray:
        out = np.zeros(self.n_occupied, np.uint64)
        lib().orc_volume_get_occupied(self.h, _p(out, C.c_ulonglong))
        return out

    def normals_csr(self):
        off = np.zeros(self.n_occupied + 1, np.uint32)
        nrm = np.zeros((lib().orc_volume_num_normals(self.h), 3), np.float32)
        lib().orc_volume_get_normals(self.h, _p(off, C.c_uint), _p(nrm, C.c_float))
        return off, nrm

    def marks(self):
        view = np.zeros(self.n_occupied, np.int32)
        good = np.zeros(self.n_occupied, np.uint8)
        lib().orc_volume_get_marks(self.h, _p(view, C.c_int), _p(good, C.c_ubyte))
        return view, good

    def clear_marks(self):
        lib().orc_volume_clear_marks(self.h)


def volume_from_scene(scene, flat=True) -> Volume:
    v = Volume(scene.bounds, scene.dims, flat)
    v.integrate(scene.points, scene.normals)
    return v


def forward(vol: Volume, K, H, W, pose12, mode, zdelta, sparse, view=1, want_pixels=True, want_counters=True):
    """Returns dict(found_any, ids, min_depth, depth, points, voxel, counters)."""
    K = np.ascontiguousarray(K, np.float32)
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    depth = np.zeros((H, W), np.int32) if want_pixels else None
    points = np.zeros((H, W, 3), np.float32) if want_pixels else None
    voxel = np.zeros((H, W), np.uint64) if want_pixels else None
    ids = np.zeros(H * W, np.uint64)
    found, md = C.c_int(), C.c_int()
    cnt = np.zeros(5, np.int64) if want_counters else None
    n = lib().orc_forward(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), mode, zdelta, int(sparse), view,
                          _p(depth, C.c_int), _p(points, C.c_float), _p(voxel, C.c_ulonglong),
                          _p(ids, C.c_ulonglong), len(ids), C.byref(found), C.byref(md), _p(cnt, C.c_longlong))
    return dict(found_any=bool(found.value), ids=ids[:n].copy(), min_depth=md.value, depth=depth, points=points,
                voxel=voxel, counters=None if cnt is None else dict(zip(("samples", "inbounds", "hits", "oob", "runaway"), cnt.tolist())))


def forward_observed(vol: Volume, K, H, W, pose12, mode, zdelta, sparse, view=1, observed=None):
    """Carve-mode extension (PixelOut::observed): returns (observed uint32 words over the padded index space, counters);
    pass the previous `observed` to accumulate over views."""
    K = np.ascontiguousarray(K, np.float32)
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    f = lib().orc_forward_observed
    n = f(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), mode, zdelta, int(sparse), view, None, None)
    if observed is None:
        observed = np.zeros(n, np.uint32)
    assert observed.dtype == np.uint32 and len(observed) == n
    cnt = np.zeros(5, np.int64)
    f(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), mode, zdelta, int(sparse), view, _p(observed, C.c_uint), _p(cnt, C.c_longlong))
    return observed, dict(zip(("samples", "inbounds", "hits", "oob", "runaway"), cnt.tolist()))


def reverse(vol: Volume, K, H, W, pose12, fast=True, viz=False, dead_work=False):
    K = np.ascontiguousarray(K, np.float32)
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    ids = np.zeros(max(vol.n_occupied, 1), np.uint64)
    flags = np.zeros(max(vol.n_occupied, 1), np.uint8)
    found = C.c_int()
    cnt = np.zeros(5, np.int64)
    n = lib().orc_reverse(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), int(fast), int(viz), int(dead_work),
                          _p(ids, C.c_ulonglong), len(ids), C.byref(found), _p(flags, C.c_ubyte), _p(cnt, C.c_longlong))
    return dict(found_any=bool(found.value), ids=ids[:n].copy(), flags=flags[: vol.n_occupied],
                counters=dict(zip(("samples", "inbounds", "hits", "oob", "runaway"), cnt.tolist())))


def zbuffer(vol: Volume, K, H, W, pose12):
    K = np.ascontiguousarray(K, np.float32)
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    depth = np.zeros((H, W), np.int32)
    counter = C.c_longlong()
    lib().orc_zbuffer(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), _p(depth, C.c_int), C.byref(counter))
    return depth, counter.value


def affine_inverse(pose12):
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    out = np.zeros(12, np.float32)
    lib().orc_affine_inverse(_p(pose, C.c_float), _p(out, C.c_float))
    return out


def degree_acosf(d: float) -> int:
    return lib().orc_degree_acosf(float(d))


def greedy_set_cover(sets):
    """sets: list of sorted uint64 arrays.  Returns selected indices in selection order."""
    off = np.zeros(len(sets) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in sets])
    ids = np.concatenate([np.asarray(s, np.uint64) for s in sets]) if len(sets) and off[-1] else np.zeros(1, np.uint64)
    sel = np.zeros(max(len(sets), 1), np.uint64)
    n = lib().orc_greedy_set_cover(_p(ids, C.c_ulonglong), _p(off, C.c_longlong), len(sets), _p(sel, C.c_ulonglong))
    return sel[:n].astype(np.int64)


def will_collide(vol: Volume, a, b, guard_coords=True):
    a = np.ascontiguousarray(a, np.float32).reshape(3)
    b = np.ascontiguousarray(b, np.float32).reshape(3)
    steps = C.c_longlong(0)
    r = lib().orc_will_collide(vol.h, _p(a, C.c_float), _p(b, C.c_float), int(guard_coords), C.byref(steps))
    return bool(r), steps.value


def optimize_standoff(vol: Volume, K, H, W, pose12):
    K = np.ascontiguousarray(K, np.float32)
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    out = np.zeros(12, np.float32)
    mid = lib().orc_optimize_standoff(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), _p(out, C.c_float))
    return int(mid), out


def time_views(vol: Volume, K, H, W, poses, kind, zdelta, sparse, threads=1):
    """Seconds for the reference's own work on n views (see orc_time_views)."""
    K = np.ascontiguousarray(K, np.float32)
    poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    tot = C.c_longlong()
    s = lib().orc_time_views(vol.h, _p(K, C.c_float), H, W, _p(poses, C.c_float), len(poses), kind, zdelta,
                             int(sparse), threads, C.byref(tot))
    return s, tot.value


def max_threads() -> int:
    return lib().orc_max_threads()
