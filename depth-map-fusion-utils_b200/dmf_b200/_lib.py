"""Loads libdmf_b200.so and declares the C ABI of include/dmf_b200.h for ctypes.

The library is the only compute path.  If it is missing, or no sm_100 GPU is usable, importing works (so the
CPU-only test tier can check symbols) but every compute call raises DmfError -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DMF_B200_LIB: another build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("DMF_B200_LIB") or os.path.join(os.path.dirname(HERE), "libdmf_b200.so")


class DmfError(RuntimeError):
    pass


class ForwardParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("zdelta", C.c_int), ("sparse", C.c_int), ("view_id0", C.c_int), ("grid_format", C.c_int), ("flags", C.c_int)]


class ForwardOut(C.Structure):
    _fields_ = [
        ("depth_mm", C.c_void_p), ("points", C.c_void_p), ("hit_voxel", C.c_void_p), ("visibility", C.c_void_p),
        ("found_any", C.c_void_p), ("min_depth", C.c_void_p), ("ids", C.c_void_p), ("ids_offsets", C.c_void_p),
        ("ids_capacity", C.c_size_t), ("depth_u16", C.c_void_p),
    ]


class ReverseOut(C.Structure):
    _fields_ = [
        ("visibility", C.c_void_p), ("unoccluded", C.c_void_p), ("found_any", C.c_void_p), ("ids", C.c_void_p),
        ("ids_offsets", C.c_void_p), ("ids_capacity", C.c_size_t),
    ]


class SweepOut(C.Structure):
    _fields_ = [("visibility", C.c_void_p), ("found_any", C.c_void_p), ("rows_to_host", C.c_int)]


# every symbol include/dmf_b200.h declares: name -> (restype, argtypes)
vp, fp, dp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
u64p, u32p, i32p, i64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
SYMBOLS = {
    "dmf_create": (C.c_int, [C.POINTER(vp), C.c_int]),
    "dmf_destroy": (None, [vp]),
    "dmf_last_error": (C.c_char_p, []),
    "dmf_device_count": (C.c_int, []),
    "dmf_version": (C.c_int, []),
    "dmf_host_alloc": (vp, [C.c_size_t]),
    "dmf_host_free": (None, [vp]),
    "dmf_set_camera": (C.c_int, [vp, fp, C.c_int, C.c_int]),
    "dmf_upload_volume": (C.c_int, [vp, dp, dp, ip, u64p, C.c_size_t, u32p, fp]),
    "dmf_volume_from_points": (C.c_int, [vp, dp, ip, fp, fp, C.c_size_t]),
    "dmf_volume_from_points_gpu": (C.c_int, [vp, dp, ip, fp, fp, C.c_size_t]),
    "dmf_volume_info": (C.c_int, [vp, ip, dp, dp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "dmf_volume_get_occupied": (C.c_int, [vp, u64p]),
    "dmf_volume_get_normals": (C.c_int, [vp, u32p, fp]),
    "dmf_clear_marks": (C.c_int, [vp]),
    "dmf_download_marks": (C.c_int, [vp, i32p, u8p, C.c_size_t]),
    "dmf_upload_marks": (C.c_int, [vp, i32p, u8p, C.c_size_t]),
    "dmf_volume_prepare_ms": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "dmf_prepare_grid": (C.c_int, [vp, C.c_int]),
    "dmf_visibility_words": (C.c_size_t, [vp]),
    "dmf_observed_words": (C.c_size_t, [vp]),
    "dmf_clear_observed": (C.c_int, [vp]),
    "dmf_download_observed": (C.c_int, [vp, u32p]),
    "dmf_observed_dev": (C.c_int, [vp, C.POINTER(vp)]),
    "dmf_observed_counts": (C.c_int, [vp, u64p]),
    "dmf_forward": (C.c_int, [vp, C.POINTER(ForwardParams), fp, C.c_int, C.POINTER(ForwardOut)]),
    "dmf_forward_dev": (C.c_int, [vp, C.POINTER(ForwardParams), vp, C.c_int, C.POINTER(ForwardOut), vp]),
    "dmf_reverse": (C.c_int, [vp, C.c_int, C.c_int, fp, C.c_int, C.POINTER(ReverseOut)]),
    "dmf_reverse_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, C.POINTER(ReverseOut), vp]),
    "dmf_optimize_standoff": (C.c_int, [vp, fp, C.c_int, C.c_uint, C.c_uint, u32p, fp]),
    "dmf_segments_collide": (C.c_int, [vp, fp, fp, C.c_int, C.c_int, u8p]),
    "dmf_zbuffer": (C.c_int, [vp, fp, i32p, i64p]),
    "dmf_greedy_set_cover": (C.c_int, [vp, u64p, C.c_int, C.c_size_t, i32p, ip]),
    "dmf_greedy_set_cover_dev": (C.c_int, [vp, vp, C.c_int, C.c_size_t, i32p, ip]),
    "dmf_or_reduce_dev": (C.c_int, [vp, vp, vp, C.c_int, C.c_size_t, vp]),
    "dmf_comm_init_all": (C.c_int, [C.POINTER(vp), C.c_int]),
    "dmf_comm_unique_id": (C.c_int, [vp]),
    "dmf_comm_init_rank": (C.c_int, [C.POINTER(vp), vp, vp, C.c_int, C.c_int]),
    "dmf_comm_destroy": (None, [vp]),
    "dmf_comm_info": (C.c_int, [vp, ip, ip, ip, ip]),
    "dmf_comm_ctx": (vp, [vp, C.c_int]),
    "dmf_comm_set_camera": (C.c_int, [vp, fp, C.c_int, C.c_int]),
    "dmf_comm_replicate_volume": (C.c_int, [vp, C.c_int]),
    "dmf_comm_synchronize": (C.c_int, [vp]),
    "dmf_sweep_forward": (C.c_int, [vp, C.POINTER(ForwardParams), fp, C.c_int, C.POINTER(SweepOut)]),
    "dmf_sweep_reverse": (C.c_int, [vp, C.c_int, fp, C.c_int, C.POINTER(SweepOut)]),
    "dmf_sweep_forward_dev": (C.c_int, [vp, C.POINTER(ForwardParams), C.POINTER(vp), C.c_int, C.POINTER(C.POINTER(ForwardOut)), C.POINTER(vp)]),
    "dmf_sweep_reverse_dev": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]),
    "dmf_sweep_gathered_dev": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), ip]),
    "dmf_sweep_set_cover": (C.c_int, [vp, i32p, ip]),
    "dmf_comm_fuse_observed": (C.c_int, [vp]),
    "dmf_comm_fuse_marks": (C.c_int, [vp, C.c_int]),
    "dmf_host_angle_test": (C.c_int, [fp]),
    "dmf_set_reverse_format": (C.c_int, [vp, C.c_int]),
    "dmf_selftest_div1000": (C.c_int, [vp, u64p]),
    "dmf_counters": (C.c_int, [vp, u64p]),
    "dmf_reset_counters": (C.c_int, [vp]),
    "dmf_last_kernel_ms": (C.c_int, [vp, fp]),
    "dmf_last_hot_kernel_ms": (C.c_int, [vp, fp]),
    "dmf_synchronize": (C.c_int, [vp]),
}

COUNTER_NAMES = ("samples", "inbounds", "hits", "exact_div", "oob", "acos_ties", "launches", "runaway", "f64_path", "skipped", "bounds", "rsv11")

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises DmfError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DmfError(f"{LIB_PATH} is missing: run `python depth-map-fusion-utils_b200/build.py` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != 0:
        raise DmfError(load().dmf_last_error().decode("utf-8", "replace"))
