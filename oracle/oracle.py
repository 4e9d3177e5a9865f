"""ctypes loader of the CPU restatement (oracle/hagrid_oracle.c) — TEST INFRASTRUCTURE.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libhagrid_oracle.so"
MAX_LEVELS = 32

CELL_DTYPE = np.dtype([("min", "<i4", 3), ("begin", "<i4"), ("max", "<i4", 3), ("end", "<i4")])
SMALL_CELL_DTYPE = np.dtype([("min", "<u2", 3), ("max", "<u2", 3), ("begin", "<i4")])
HIT_DTYPE = np.dtype([("id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])


class OGrid(C.Structure):
    _fields_ = [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("dims", C.c_int * 3), ("shift", C.c_int),
                ("num_cells", C.c_int), ("num_entries", C.c_int), ("num_refs", C.c_int), ("compressed", C.c_int),
                ("num_offsets", C.c_int), ("offsets", C.c_int * MAX_LEVELS),
                ("entries", C.c_void_p), ("cells", C.c_void_p), ("small_cells", C.c_void_p), ("refs", C.c_void_p)]


def build_library(force: bool = False) -> Path:
    src = HERE / "hagrid_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_dll = None


def dll():
    global _dll
    if _dll is None:
        build_library()
        _dll = C.CDLL(str(LIB))
        P = C.POINTER(OGrid)
        _dll.og_build.restype = P
        _dll.og_build.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        _dll.og_grid_from_arrays.restype = P
        _dll.og_grid_from_arrays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _dll.og_grid_free.argtypes = [P]
        _dll.og_merge.argtypes = [P, C.c_float]
        _dll.og_flatten.argtypes = [P]
        _dll.og_expand.argtypes = [P, C.c_int]
        _dll.og_compress.argtypes = [P]
        _dll.og_compress.restype = C.c_int
        _dll.og_traverse.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _dll.og_gen_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        _dll.og_gen_rays.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        _dll.og_update_surface.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        _dll.og_bounce_rays.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                        C.c_uint32, C.c_void_p]
        _dll.og_traverse_record.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _dll.og_traverse_record_cells.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return _dll


def _arr(ptr, dtype, n):
    if n == 0 or not ptr:
        return np.empty(0, dtype=dtype)
    buf = (C.c_char * (np.dtype(dtype).itemsize * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Grid:
    """Owning wrapper of an og_grid."""

    def __init__(self, ptr):
        self.ptr = ptr

    @classmethod
    def build(cls, tris: np.ndarray, top_density: float, snd_density: float) -> "Grid":
        tris = np.ascontiguousarray(tris)
        return cls(dll().og_build(tris.ctypes.data, tris.shape[0], top_density, snd_density))

    @classmethod
    def from_arrays(cls, info: dict, entries, cells, refs) -> "Grid":
        bmin = np.asarray(info["bbox_min"], np.float32); bmax = np.asarray(info["bbox_max"], np.float32)
        dims = np.asarray(info["dims"], np.int32); offs = np.asarray(info["offsets"], np.int32)
        entries = np.ascontiguousarray(entries, "<u4"); refs = np.ascontiguousarray(refs, "<i4"); cells = np.ascontiguousarray(cells)
        return cls(dll().og_grid_from_arrays(bmin.ctypes.data, bmax.ctypes.data, dims.ctypes.data, info["shift"],
                                             info["num_cells"], info["num_entries"], info["num_refs"], info["compressed"],
                                             len(offs), offs.ctypes.data, entries.ctypes.data, cells.ctypes.data, refs.ctypes.data))

    def __del__(self):
        if getattr(self, "ptr", None):
            dll().og_grid_free(self.ptr)
            self.ptr = None

    def merge(self, alpha=0.995): dll().og_merge(self.ptr, alpha)
    def flatten(self): dll().og_flatten(self.ptr)
    def expand(self, iters=3): dll().og_expand(self.ptr, iters)
    def compress(self) -> bool: return bool(dll().og_compress(self.ptr))

    def info(self) -> dict:
        g = self.ptr.contents
        return {"bbox_min": list(g.bbox_min), "bbox_max": list(g.bbox_max), "dims": list(g.dims), "shift": g.shift,
                "num_cells": g.num_cells, "num_entries": g.num_entries, "num_refs": g.num_refs, "compressed": g.compressed,
                "offsets": list(g.offsets[:g.num_offsets])}

    def arrays(self):
        g = self.ptr.contents
        cells = _arr(g.small_cells, SMALL_CELL_DTYPE, g.num_cells) if g.compressed else _arr(g.cells, CELL_DTYPE, g.num_cells)
        return _arr(g.entries, "<u4", g.num_entries), cells, _arr(g.refs, "<i4", g.num_refs)

    def traverse(self, tris: np.ndarray, rays: np.ndarray, mode: int = 1, threads: int = 1) -> np.ndarray:
        tris = np.ascontiguousarray(tris); rays = np.ascontiguousarray(rays)
        hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        dll().og_traverse(self.ptr, tris.ctypes.data, rays.ctypes.data, hits.ctypes.data, rays.shape[0], mode, threads)
        return hits

    def record(self, tris: np.ndarray, rays: np.ndarray, max_steps: int = 128) -> np.ndarray:
        """Statistics: (num_rays, max_steps) int16, reference count of each visited cell, -1 padded."""
        tris = np.ascontiguousarray(tris); rays = np.ascontiguousarray(rays)
        out = np.empty((rays.shape[0], max_steps), dtype=np.int16)
        dll().og_traverse_record(self.ptr, tris.ctypes.data, rays.ctypes.data, out.ctypes.data, max_steps, rays.shape[0])
        return out

    def record_cells(self, tris: np.ndarray, rays: np.ndarray, max_steps: int = 128) -> np.ndarray:
        """Statistics: (num_rays, max_steps) int32, index of each visited cell, -1 padded."""
        tris = np.ascontiguousarray(tris); rays = np.ascontiguousarray(rays)
        out = np.empty((rays.shape[0], max_steps), dtype=np.int32)
        dll().og_traverse_record_cells(self.ptr, tris.ctypes.data, rays.ctypes.data, out.ctypes.data, max_steps, rays.shape[0])
        return out


RAY_DTYPE = np.dtype([("org", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])


def gen_camera(eye, center, up, fov: float, ratio: float) -> np.ndarray:
    """gen_camera (src/main.cpp:42-50): 12 floats eye, right, up, dir."""
    e, c, u = (np.ascontiguousarray(v, dtype="<f4") for v in (eye, center, up))
    cam = np.empty(12, dtype="<f4")
    dll().og_gen_camera(e.ctypes.data, c.ctypes.data, u.ctypes.data, fov, ratio, cam.ctypes.data)
    return cam


def gen_rays(cam: np.ndarray, clip: float, width: int, height: int) -> np.ndarray:
    """gen_rays (src/main.cpp:52-66)."""
    cam = np.ascontiguousarray(cam, dtype="<f4")
    rays = np.empty(width * height, dtype=RAY_DTYPE)
    dll().og_gen_rays(cam.ctypes.data, clip, width, height, rays.ctypes.data)
    return rays


def update_surface(mode: int, hits: np.ndarray, clip: float, width: int, height: int) -> np.ndarray:
    """update_surface (src/main.cpp:90-111): (height, width, 4) BGRA bytes."""
    hits = np.ascontiguousarray(hits)
    out = np.empty((height, width, 4), dtype=np.uint8)
    dll().og_update_surface(mode, hits.ctypes.data, clip, width, height, out.ctypes.data)
    return out


def bounce_rays(tris: np.ndarray, rays: np.ndarray, hits: np.ndarray, offset: float, tmax: float, seed: int) -> np.ndarray:
    """Second-wave rays as include/hagrid_b200.h defines them (hgb_generate_bounce_rays); `hits` hold primitive ids."""
    tris = np.ascontiguousarray(tris); rays = np.ascontiguousarray(rays); hits = np.ascontiguousarray(hits)
    assert rays.dtype == RAY_DTYPE and hits.dtype.itemsize == 16 and rays.shape[0] == hits.shape[0]
    out = np.empty(rays.shape[0], dtype=RAY_DTYPE)
    dll().og_bounce_rays(tris.ctypes.data, tris.shape[0], rays.ctypes.data, hits.ctypes.data, rays.shape[0],
                         offset, tmax, seed & 0xFFFFFFFF, out.ctypes.data)
    return out
