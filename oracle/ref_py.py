"""ctypes binding of oracle/_ref/libref_dmf.so: the reference's OWN hot-path headers (compiled from /root/reference against
oracle/ref_shim by `make -C oracle ref`).  TEST INFRASTRUCTURE: validates the restatement in dmf_oracle.hpp and serves as
the timed CPU baseline (kind "reference").  available() is False where the library has not been built."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("DMF_REF_LIB") or os.path.join(_HERE, "_ref", "libref_dmf.so")      # DMF_REF_LIB: e.g. an ASan build
REF_TREE = os.environ.get("DMF_REFERENCE_TREE", "/root/reference")


def build() -> bool:
    """(Re)build where the reference tree exists; elsewhere keep whatever prebuilt library travelled here."""
    if os.path.exists(os.path.join(REF_TREE, "include", "RayTracingEngine.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REF={REF_TREE}"])
    return available()


def available() -> bool:
    return os.path.exists(_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB)
        vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
        u64p, llp, ucp = C.POINTER(C.c_ulonglong), C.POINTER(C.c_longlong), C.POINTER(C.c_ubyte)
        L.ref_volume_new.restype = vp
        L.ref_volume_new.argtypes = [dp, ip]
        L.ref_volume_free.argtypes = [vp]
        L.ref_volume_info.argtypes = [vp, ip, dp, dp]
        L.ref_volume_integrate.argtypes = [vp, fp, fp, C.c_long]
        L.ref_volume_num_occupied.restype = C.c_long
        L.ref_volume_num_occupied.argtypes = [vp]
        L.ref_volume_get_occupied.argtypes = [vp, u64p]
        L.ref_volume_get_marks.argtypes = [vp, ip, ucp]
        L.ref_volume_clear_marks.argtypes = [vp]
        L.ref_forward.restype = C.c_long
        L.ref_forward.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, u64p, C.c_long, ip, ip]
        L.ref_reverse.restype = C.c_long
        L.ref_reverse.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, u64p, C.c_long, ip]
        L.ref_zbuffer.argtypes = [vp, fp, C.c_int, C.c_int, fp]
        L.ref_time_views.restype = C.c_double
        L.ref_time_views.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, llp]
        L.ref_max_threads.restype = C.c_int
        L.ref_greedy_set_cover.restype = C.c_long
        L.ref_greedy_set_cover.argtypes = [u64p, llp, C.c_long, u64p]
        L.ref_position_cameras.argtypes = [fp, fp, C.c_long, C.c_uint, fp]
        L.ref_generate_sphere.restype = C.c_long
        L.ref_generate_sphere.argtypes = [C.c_double, fp, C.c_double, C.c_double, fp, C.c_long]
        L.ref_reposition_camera.argtypes = [fp, C.c_uint, fp]
        L.ref_optimize_camera_position.argtypes = [vp, fp, C.c_int, C.c_int, fp, fp]
        L.ref_will_collide.restype = C.c_int
        L.ref_will_collide.argtypes = [vp, fp, fp]
        L.ref_reposition_cameras_sampled.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_long, fp]
        L.ref_set_cover.restype = C.c_long
        L.ref_set_cover.argtypes = [vp, fp, C.c_int, C.c_int, fp, C.c_long, u64p]
        L.ref_write_camera_locations.argtypes = [C.c_char_p, fp, C.c_long]
        L.ref_read_camera_locations.restype = C.c_long
        L.ref_read_camera_locations.argtypes = [C.c_char_p, fp, C.c_long]
        _lib = L
    return _lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


class Volume:
    def __init__(self, bounds, dims):
        b = np.ascontiguousarray(bounds, np.float64); d = np.ascontiguousarray(dims, np.int32)
        self.h = lib().ref_volume_new(_p(b, C.c_double), _p(d, C.c_int))
        od = np.zeros(3, np.int32); dl = np.zeros(3, np.float64); vs = C.c_double()
        lib().ref_volume_info(self.h, _p(od, C.c_int), _p(dl, C.c_double), C.byref(vs))
        self.dims, self.deltas, self.voxel_size = od, dl, vs.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_volume_free(self.h); self.h = None

    def integrate(self, pts, normals):
        pts = np.ascontiguousarray(pts, np.float32); nrm = np.ascontiguousarray(normals, np.float32)
        lib().ref_volume_integrate(self.h, _p(pts, C.c_float), _p(nrm, C.c_float), len(pts))

    @property
    def n_occupied(self):
        return lib().ref_volume_num_occupied(self.h)

    def occupied(self):
        out = np.zeros(self.n_occupied, np.uint64)
        lib().ref_volume_get_occupied(self.h, _p(out, C.c_ulonglong))
        return out

    def marks(self):
        view = np.zeros(self.n_occupied, np.int32); good = np.zeros(self.n_occupied, np.uint8)
        lib().ref_volume_get_marks(self.h, _p(view, C.c_int), _p(good, C.c_ubyte))
        return view, good

    def clear_marks(self):
        lib().ref_volume_clear_marks(self.h)


def volume_from_scene(scene) -> Volume:
    v = Volume(scene.bounds, scene.dims)
    v.integrate(scene.points, scene.normals)
    return v


def forward(vol, K, H, W, pose12, mode, zdelta, sparse, view=1):
    K = np.ascontiguousarray(K, np.float32); pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    ids = np.zeros(H * W, np.uint64); found, md = C.c_int(), C.c_int()
    n = lib().ref_forward(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), mode, zdelta, int(sparse), view, _p(ids, C.c_ulonglong), len(ids), C.byref(found), C.byref(md))
    return dict(found_any=bool(found.value), ids=ids[:n].copy(), min_depth=md.value)


def reverse(vol, K, H, W, pose12, fast=True, viz=False):
    K = np.ascontiguousarray(K, np.float32); pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    ids = np.zeros(2 * max(vol.n_occupied, 1) + 64, np.uint64); found = C.c_int()
    n = lib().ref_reverse(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), int(fast), int(viz), _p(ids, C.c_ulonglong), len(ids), C.byref(found))
    return dict(found_any=bool(found.value), ids=ids[:n].copy())


def zbuffer(vol, K, H, W, pose12):
    K = np.ascontiguousarray(K, np.float32); pose = np.ascontiguousarray(pose12, np.float32).reshape(12)
    lib().ref_zbuffer(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float))


def time_views(vol, K, H, W, poses, kind, zdelta, sparse, threads=1):
    K = np.ascontiguousarray(K, np.float32); poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    tot = C.c_longlong()
    s = lib().ref_time_views(vol.h, _p(K, C.c_float), H, W, _p(poses, C.c_float), len(poses), kind, zdelta, int(sparse), threads, C.byref(tot))
    return s, tot.value


def max_threads():
    return lib().ref_max_threads()


# ---- the reference's Algorithms.hpp (compiled unmodified, see ref_capi.cpp) -------------------------------------------
def greedy_set_cover(sets):
    """Algorithms::greedySetCover (Algorithms.hpp:38-86).  sets: list of sorted uint64 arrays -> selected indices in order."""
    off = np.zeros(len(sets) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in sets])
    ids = np.concatenate([np.asarray(s, np.uint64) for s in sets]) if len(sets) and off[-1] else np.zeros(1, np.uint64)
    sel = np.zeros(max(len(sets), 1), np.uint64)
    n = lib().ref_greedy_set_cover(_p(ids, C.c_ulonglong), _p(off, C.c_longlong), len(sets), _p(sel, C.c_ulonglong))
    return sel[:n].astype(np.int64)


def position_cameras(points, normals, distance=300):
    """Algorithms::positionCameras(locations, distance) (Algorithms.hpp:282-298) -> (n, 12) float32 poses."""
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3); nrm = np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
    out = np.zeros((len(pts), 12), np.float32)
    lib().ref_position_cameras(_p(pts, C.c_float), _p(nrm, C.c_float), len(pts), int(distance), _p(out, C.c_float))
    return out


def generate_sphere(radius, transformation12=None, z_threshold=0.1, factor=10.0):
    """Algorithms::generateSphere (Algorithms.hpp:88-112) -> (n, 3) float32 points."""
    T = np.ascontiguousarray(transformation12 if transformation12 is not None else [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32).reshape(12)
    cap = int((2 * factor + 2) * (factor + 2)) + 16
    out = np.zeros((cap, 3), np.float32)
    n = lib().ref_generate_sphere(float(radius), _p(T, C.c_float), float(z_threshold), float(factor), _p(out, C.c_float), cap)
    assert n <= cap
    return out[:n].copy()


def reposition_camera(pose12, distance=300):
    pose = np.ascontiguousarray(pose12, np.float32).reshape(12); out = np.zeros(12, np.float32)
    lib().ref_reposition_camera(_p(pose, C.c_float), int(distance), _p(out, C.c_float))
    return out


def optimize_camera_position(vol, K, H, W, pose12):
    """Algorithms::optimizeCameraPosition(volume, engine, res, camera) (Algorithms.hpp:394-421) -> the camera it returns."""
    K = np.ascontiguousarray(K, np.float32); pose = np.ascontiguousarray(pose12, np.float32).reshape(12); out = np.zeros(12, np.float32)
    lib().ref_optimize_camera_position(vol.h, _p(K, C.c_float), H, W, _p(pose, C.c_float), _p(out, C.c_float))
    return out


# ---- the reference's FileRoutines.hpp: camera-pose text files (:69-112) -------------------------------------------------
def write_camera_locations(filename, poses):
    a = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    lib().ref_write_camera_locations(str(filename).encode(), _p(a, C.c_float), len(a))


def read_camera_locations(filename, cap=4096):
    out = np.zeros((cap, 12), np.float32)
    n = lib().ref_read_camera_locations(str(filename).encode(), _p(out, C.c_float), cap)
    assert n <= cap
    return out[:n].copy()


# ---- the reference's driver tests/CameraPathGen.cpp (compiled with its main renamed away) -------------------------------
def will_collide(vol, a, b):
    a = np.ascontiguousarray(a, np.float32).reshape(3); b = np.ascontiguousarray(b, np.float32).reshape(3)
    return bool(lib().ref_will_collide(vol.h, _p(a, C.c_float), _p(b, C.c_float)))


def reposition_cameras_sampled(vol, K, H, W, poses):
    K = np.ascontiguousarray(K, np.float32); poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    out = np.zeros_like(poses)
    lib().ref_reposition_cameras_sampled(vol.h, _p(K, C.c_float), H, W, _p(poses, C.c_float), len(poses), _p(out, C.c_float))
    return out


def set_cover(vol, K, H, W, poses):
    K = np.ascontiguousarray(K, np.float32); poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
    sel = np.zeros(max(len(poses), 1), np.uint64)
    n = lib().ref_set_cover(vol.h, _p(K, C.c_float), H, W, _p(poses, C.c_float), len(poses), _p(sel, C.c_ulonglong))
    return sel[:n].astype(np.int64)
