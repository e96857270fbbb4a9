"""The multi-GPU view sweep of libdmf_b200 (include/dmf_b200.h "multi-GPU") from Python.

    comm = Comm.init_all(n)                     # ONE process drives n GPUs (the shape of the reference's C++ drivers)
    comm = Comm.init_rank(ctx, uid, rank, world)  # one process per GPU (torchrun): uid = Comm.unique_id() of rank 0, shared out of band

Everything that computes or communicates is inside the library (peer-mapped stores from the march kernels, or NCCL loaded
by the library itself); torch.distributed is at most the out-of-band channel for the 128-byte id.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import ForwardParams, SweepOut, check
from .engine import Context, GRID_BYTE, MODE_POINTS, _poses12

EXCHANGE_NAMES = {0: "none (1 GPU)", 1: "peer-mapped stores over NVLink by a push kernel behind the march", 2: "ncclAllGather",
                  3: "peer-mapped stores over NVLink from the march kernels' epilogue (ticket per view)"}
ROWS_OWN, ROWS_ALL = 0, 1


class Comm:
    def __init__(self, h, lib, contexts):
        self.h, self.lib, self.contexts = h, lib, contexts

    # ---- construction -----------------------------------------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        check(lib.dmf_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def init_all(cls, n_gpus: int = 0) -> "Comm":
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.dmf_comm_init_all(C.byref(h), n_gpus))
        self = cls(h, lib, [])
        for i in range(self.info()["n_local"]):
            c = Context.__new__(Context)
            c.h, c.lib, c.device, c._volume_token, c._borrowed = C.c_void_p(lib.dmf_comm_ctx(h, i)), lib, i, None, True
            self.contexts.append(c)
        return self

    @classmethod
    def init_rank(cls, ctx: Context, unique_id: bytes, rank: int, world: int) -> "Comm":
        lib = _lib.load()
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128) if world > 1 else None
        check(lib.dmf_comm_init_rank(C.byref(h), ctx.h, buf, rank, world))
        return cls(h, lib, [ctx])

    def close(self):
        if getattr(self, "h", None):
            self.lib.dmf_comm_destroy(self.h)
            self.h = None
            for c in self.contexts:
                if getattr(c, "_borrowed", False):
                    c.h = None                                   # owned (and just destroyed) by the group

    def info(self) -> dict:
        w, n, f, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.dmf_comm_info(self.h, C.byref(w), C.byref(n), C.byref(f), C.byref(e)))
        return dict(world=w.value, n_local=n.value, first_rank=f.value, exchange=e.value, exchange_name=EXCHANGE_NAMES[e.value])

    # ---- set-up -----------------------------------------------------------------------------------------------
    def set_camera(self, K, height: int, width: int):
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        check(self.lib.dmf_comm_set_camera(self.h, K.ctypes.data_as(C.POINTER(C.c_float)), height, width))

    def replicate_volume(self, root: int = 0):
        check(self.lib.dmf_comm_replicate_volume(self.h, root))
        first = self.info()["first_rank"]
        for i, c in enumerate(self.contexts):
            if first + i != root:
                c._volume_token = None                           # this context now holds the root's volume, not what Python uploaded before

    def synchronize(self):
        check(self.lib.dmf_comm_synchronize(self.h))

    # ---- sweeps -----------------------------------------------------------------------------------------------
    def _out(self, n, want_rows, rows_to_host):
        vw = self.lib.dmf_visibility_words(self.contexts[0].h)
        vis = np.zeros((n, vw), np.uint64) if want_rows else None
        found = np.zeros(n, np.int32)
        o = SweepOut()
        o.visibility = vis.ctypes.data if vis is not None and vis.size else None
        o.found_any = found.ctypes.data
        o.rows_to_host = rows_to_host
        return o, vis, found

    def sweep_forward(self, poses, mode: int = MODE_POINTS, zdelta: int = 10, sparse: bool = False, view_id0: int = 1, grid_format: int = GRID_BYTE,
                      flags: int = 0, want_rows: bool = True, rows_to_host: int = ROWS_ALL) -> dict:
        poses = _poses12(poses)
        o, vis, found = self._out(len(poses), want_rows, rows_to_host)
        p = ForwardParams(mode, int(zdelta), int(bool(sparse)), int(view_id0), grid_format, flags)
        check(self.lib.dmf_sweep_forward(self.h, C.byref(p), poses.ctypes.data_as(C.POINTER(C.c_float)), len(poses), C.byref(o)))
        return dict(visibility=vis, found_any=found)

    def sweep_reverse(self, poses, want_rows: bool = True, rows_to_host: int = ROWS_ALL) -> dict:
        poses = _poses12(poses)
        o, vis, found = self._out(len(poses), want_rows, rows_to_host)
        check(self.lib.dmf_sweep_reverse(self.h, 1, poses.ctypes.data_as(C.POINTER(C.c_float)), len(poses), C.byref(o)))
        return dict(visibility=vis, found_any=found)

    def gathered_dev(self, local_index: int = 0) -> dict:
        p, rw, vw, n = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_int()
        check(self.lib.dmf_sweep_gathered_dev(self.h, local_index, C.byref(p), C.byref(rw), C.byref(vw), C.byref(n)))
        return dict(ptr=int(p.value), row_words=rw.value, vis_words=vw.value, n_views=n.value)

    def set_cover(self) -> np.ndarray:
        n = self.gathered_dev()["n_views"]
        sel = np.zeros(max(n, 1), np.int32)
        cnt = C.c_int()
        check(self.lib.dmf_sweep_set_cover(self.h, sel.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(cnt)))
        return sel[: cnt.value].copy()

    def fuse_observed(self):
        check(self.lib.dmf_comm_fuse_observed(self.h))

    def fuse_marks(self, view_id0: int = 1):
        check(self.lib.dmf_comm_fuse_marks(self.h, int(view_id0)))
