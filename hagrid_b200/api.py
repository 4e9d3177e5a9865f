"""ctypes binding of include/hagrid_b200.h — the host-side mirror of the
reference's build/traverse interface (src/build.h:17-31, src/traverse.h:11-14,
driven the way src/main.cpp:471-549 drives it).

numpy structured dtypes below are the byte layouts of the reference's structs
(SURVEY.md A.1). Device memory for rays/hits may come from torch tensors
(`tensor.data_ptr()`) or from `Scene.device_alloc`; nothing here computes on
the CPU: every call goes into the CUDA library, and loading fails loudly when
the library is missing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
DEFAULT_LIB = PKG / "libhagrid_b200.so"

HIT_STEPS = 0      # Hit.id = traversal step count (src/traverse.cu:93, reference-verbatim)
HIT_PRIM_ID = 1    # Hit.id = primitive index, -1 = miss (src/ray.h:22)

ARRAY_ENTRIES, ARRAY_CELLS, ARRAY_SMALL_CELLS, ARRAY_REFS, ARRAY_TRIS = range(5)
MAX_LEVELS = 32

TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("nx", "<f4"), ("e1", "<f4", 3), ("ny", "<f4"),
                      ("e2", "<f4", 3), ("nz", "<f4")])                      # src/prims.h:13-16
RAY_DTYPE = np.dtype([("org", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])   # src/ray.h:9-20
HIT_DTYPE = np.dtype([("id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])                  # src/ray.h:23-33
CELL_DTYPE = np.dtype([("min", "<i4", 3), ("begin", "<i4"), ("max", "<i4", 3), ("end", "<i4")])  # src/grid.h:23-33
SMALL_CELL_DTYPE = np.dtype([("min", "<u2", 3), ("max", "<u2", 3), ("begin", "<i4")])            # src/grid.h:36-45
assert TRI_DTYPE.itemsize == 48 and RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16
assert CELL_DTYPE.itemsize == 32 and SMALL_CELL_DTYPE.itemsize == 16


class GridInfo(C.Structure):
    """hgb_grid_info: the host-visible fields of hagrid::Grid (src/grid.h:48-62)."""
    _fields_ = [("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("dims", C.c_int32 * 3),
                ("shift", C.c_int32), ("num_cells", C.c_int32), ("num_entries", C.c_int32),
                ("num_refs", C.c_int32), ("compressed", C.c_int32), ("num_offsets", C.c_int32),
                ("offsets", C.c_int32 * MAX_LEVELS)]

    def as_dict(self):
        return {"bbox_min": list(self.bbox_min), "bbox_max": list(self.bbox_max), "dims": list(self.dims),
                "shift": self.shift, "num_cells": self.num_cells, "num_entries": self.num_entries,
                "num_refs": self.num_refs, "compressed": self.compressed,
                "offsets": list(self.offsets[:self.num_offsets])}


_SIGNATURES = {
    "hgb_impl": (C.c_char_p, []),
    "hgb_last_error": (C.c_char_p, []),
    "hgb_device_count": (C.c_int, []),
    "hgb_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "hgb_tile_costs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "hgb_scene_create": (C.c_void_p, [C.c_int, C.c_int]),
    "hgb_scene_destroy": (None, [C.c_void_p]),
    "hgb_scene_set_tris": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "hgb_scene_load_obj": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "hgb_obj_parse": (C.c_void_p, [C.c_char_p, C.c_int]),
    "hgb_obj_num_vertices": (C.c_int, [C.c_void_p]),
    "hgb_obj_num_tris": (C.c_int, [C.c_void_p]),
    "hgb_obj_vertices": (C.c_void_p, [C.c_void_p]),
    "hgb_obj_indices": (C.c_void_p, [C.c_void_p]),
    "hgb_obj_free": (None, [C.c_void_p]),
    "hgb_scene_num_tris": (C.c_int, [C.c_void_p]),
    "hgb_scene_peak_bytes": (C.c_size_t, [C.c_void_p]),
    "hgb_build_grid": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "hgb_merge_grid": (C.c_int, [C.c_void_p, C.c_float]),
    "hgb_flatten_grid": (C.c_int, [C.c_void_p]),
    "hgb_expand_grid": (C.c_int, [C.c_void_p, C.c_int]),
    "hgb_compress_grid": (C.c_int, [C.c_void_p]),
    "hgb_build_pipeline": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p]),
    "hgb_setup_traversal": (C.c_int, [C.c_void_p]),
    "hgb_traverse_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hgb_traverse_timed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "hgb_traverse_grid_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hgb_make_camera": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "hgb_generate_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    "hgb_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "hgb_generate_bounce_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint,
                                          C.c_void_p]),
    "hgb_generate_bounce_rays_keyed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint,
                                                C.c_void_p, C.c_void_p]),
    "hgb_count_hits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hgb_trace_two_waves": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_uint,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hgb_trace_two_waves_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_uint,
                                          C.c_void_p, C.c_void_p]),
    "hgb_prim_exclusive_scan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hgb_prim_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "hgb_prim_partition": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "hgb_prim_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "hgb_save_image": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "hgb_rays_file_count": (C.c_longlong, [C.c_char_p]),
    "hgb_load_rays": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_float, C.c_float, C.c_void_p]),
    "hgb_save_rays": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "hgb_grid_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "hgb_grid_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "hgb_grid_get_info": (C.c_int, [C.c_void_p, C.POINTER(GridInfo)]),
    "hgb_grid_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "hgb_grid_upload": (C.c_int, [C.c_void_p, C.POINTER(GridInfo), C.c_void_p, C.c_void_p, C.c_void_p]),
    "hgb_device_alloc": (C.c_void_p, [C.c_void_p, C.c_size_t]),
    "hgb_device_free": (None, [C.c_void_p, C.c_void_p]),
    "hgb_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hgb_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "hgb_device_synchronize": (C.c_int, []),
    "hgb_kernel_launches": (C.c_ulonglong, []),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class HagridError(RuntimeError):
    pass


class Library:
    """One loaded implementation of the C ABI (this library by default)."""

    def __init__(self, path: str | Path | None = None):
        self.path = Path(path) if path else DEFAULT_LIB
        if not self.path.exists():
            raise HagridError(
                f"{self.path} is missing: build it with `python -m hagrid_b200.build` "
                "(there is no CPU fallback for this path)")
        self.dll = C.CDLL(str(self.path), mode=C.RTLD_LOCAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args
        self.impl = self.dll.hgb_impl().decode()

    def check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise HagridError(f"{what}: {self.dll.hgb_last_error().decode()}")
        return rc

    def device_count(self) -> int:
        return self.dll.hgb_device_count()

    def set_option(self, key: str, value: int):
        self.check(self.dll.hgb_set_option(key.encode(), int(value)), "set_option")

    def kernel_launches(self) -> int:
        return int(self.dll.hgb_kernel_launches())

    def synchronize(self):
        self.check(self.dll.hgb_device_synchronize(), "synchronize")


def make_camera(eye, center, up, fov: float, ratio: float, lib: "Library | None" = None) -> np.ndarray:
    """gen_camera (src/main.cpp:42-50): 12 floats eye, right, up, dir."""
    lib = lib or library()
    e, c, u = (np.ascontiguousarray(v, dtype="<f4") for v in (eye, center, up))
    cam = np.empty(12, dtype="<f4")
    lib.check(lib.dll.hgb_make_camera(_ptr(e), _ptr(c), _ptr(u), fov, ratio, _ptr(cam)), "make_camera")
    return cam


def save_image(path, bgra: np.ndarray, lib: "Library | None" = None):
    """A frame as returned by Scene.render_frame ((height, width, 4) BGRA bytes) to a binary PPM file."""
    lib = lib or library()
    bgra = np.ascontiguousarray(bgra, dtype=np.uint8)
    assert bgra.ndim == 3 and bgra.shape[2] == 4
    lib.check(lib.dll.hgb_save_image(str(path).encode(), _ptr(bgra), bgra.shape[1], bgra.shape[0]), "save_image")


def parse_obj(path, threads: int = 0, lib: "Library | None" = None):
    """Host half of the ingest: (vertices (n, 3) float32 with the dummy vertex at index 0, indices (t, 3) int32)."""
    lib = lib or library()
    h = lib.dll.hgb_obj_parse(str(path).encode(), threads)
    if not h:
        raise HagridError(lib.dll.hgb_last_error().decode())
    try:
        nv, nt = lib.dll.hgb_obj_num_vertices(h), lib.dll.hgb_obj_num_tris(h)
        verts = np.frombuffer((C.c_char * (12 * nv)).from_address(lib.dll.hgb_obj_vertices(h)), dtype="<f4").reshape(nv, 3).copy()
        idx = (np.frombuffer((C.c_char * (12 * nt)).from_address(lib.dll.hgb_obj_indices(h)), dtype="<i4").reshape(nt, 3).copy()
               if nt else np.empty((0, 3), dtype="<i4"))
        return verts, idx
    finally:
        lib.dll.hgb_obj_free(h)


_default_library: Library | None = None


def library() -> Library:
    global _default_library
    if _default_library is None:
        _default_library = Library()
    return _default_library


def _ptr(x) -> int:
    """Raw address of a numpy array, an int address, or anything with data_ptr() (torch)."""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return int(x.data_ptr())


class Scene:
    """MemManager + device triangles + Grid, as src/main.cpp:471-478 sets them up."""

    def __init__(self, tris, device: int = 0, keep_alive: bool = False, lib: Library | None = None, threads: int = 0):
        """`tris`: (n,) triangle records, or the path of a Wavefront OBJ file (load_model, src/main.cpp:246-275)."""
        self.lib = lib or library()
        self._h = None
        if self.lib.device_count() <= device:
            raise HagridError(f"CUDA device {device} not available (hagrid_b200 has no CPU fallback)")
        if isinstance(tris, (str, Path)):
            self._h = self.lib.dll.hgb_scene_create(device, int(keep_alive))
            if not self._h:
                raise HagridError(self.lib.dll.hgb_last_error().decode())
            self.num_tris = self.lib.check(self.lib.dll.hgb_scene_load_obj(self._h, str(tris).encode(), threads), "load_obj")
            return
        tris = np.ascontiguousarray(tris)
        if tris.dtype != TRI_DTYPE:
            tris = tris.astype("<f4", copy=False).reshape(-1, 12).view(TRI_DTYPE).reshape(-1)
        self.num_tris = int(tris.shape[0])
        self._h = self.lib.dll.hgb_scene_create(device, int(keep_alive))
        if not self._h:
            raise HagridError(self.lib.dll.hgb_last_error().decode())
        self.lib.check(self.lib.dll.hgb_scene_set_tris(self._h, _ptr(tris), self.num_tris), "set_tris")

    def set_tris(self, tris: np.ndarray):
        """Replaces the scene's triangles (dynamic scenes: new geometry, then build_all again)."""
        tris = np.ascontiguousarray(tris)
        if tris.dtype != TRI_DTYPE:
            tris = tris.astype("<f4", copy=False).reshape(-1, 12).view(TRI_DTYPE).reshape(-1)
        self.num_tris = int(tris.shape[0])
        self.lib.check(self.lib.dll.hgb_scene_set_tris(self._h, _ptr(tris), self.num_tris), "set_tris")

    def close(self):
        if self._h:
            self.lib.dll.hgb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- construction stages (src/build.h:17-31) ---------------------------------
    def build_grid(self, top_density=0.12, snd_density=2.4):
        self.lib.check(self.lib.dll.hgb_build_grid(self._h, top_density, snd_density), "build_grid")

    def merge_grid(self, alpha=0.995):
        self.lib.check(self.lib.dll.hgb_merge_grid(self._h, alpha), "merge_grid")

    def flatten_grid(self):
        self.lib.check(self.lib.dll.hgb_flatten_grid(self._h), "flatten_grid")

    def expand_grid(self, iters=3):
        self.lib.check(self.lib.dll.hgb_expand_grid(self._h, iters), "expand_grid")

    def compress_grid(self) -> bool:
        return bool(self.lib.check(self.lib.dll.hgb_compress_grid(self._h), "compress_grid"))

    def build_all(self, top_density=0.12, snd_density=2.4, alpha=0.995, expansion=3, compress=False,
                  warmup=0, iters=1) -> np.ndarray:
        """The timed construction loop of src/main.cpp:480-508; returns ms per timed pass."""
        ms = np.zeros(max(iters, 1), dtype=np.float32)
        self.lib.check(self.lib.dll.hgb_build_pipeline(self._h, top_density, snd_density, alpha, expansion,
                                                       int(compress), warmup, iters, _ptr(ms)), "build_pipeline")
        return ms[:iters]

    # --- traversal (src/traverse.h:11-14) ---------------------------------------
    def setup_traversal(self):
        self.lib.check(self.lib.dll.hgb_setup_traversal(self._h), "setup_traversal")

    def traverse(self, dev_rays, dev_hits, num_rays: int, hit_mode: int = HIT_PRIM_ID):
        self.lib.check(self.lib.dll.hgb_traverse_grid(self._h, _ptr(dev_rays), _ptr(dev_hits), num_rays, hit_mode),
                       "traverse_grid")

    def traverse_timed(self, dev_rays, dev_hits, num_rays: int, hit_mode=HIT_PRIM_ID, warmup=3, iters=10):
        ms = np.zeros(max(iters, 1), dtype=np.float32)
        self.lib.check(self.lib.dll.hgb_traverse_timed(self._h, _ptr(dev_rays), _ptr(dev_hits), num_rays, hit_mode,
                                                       warmup, iters, _ptr(ms)), "traverse_timed")
        return ms[:iters]

    def traverse_host(self, rays: np.ndarray, hit_mode: int = HIT_PRIM_ID, hits: np.ndarray | None = None):
        """Interactive-frame shape (src/main.cpp:599-613): H2D, traverse, D2H inside the call."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        if hits is None:
            hits = np.empty(rays.shape[0], dtype=HIT_DTYPE)
        self.lib.check(self.lib.dll.hgb_traverse_grid_host(self._h, _ptr(rays), _ptr(hits), rays.shape[0], hit_mode),
                       "traverse_grid_host")
        return hits

    # --- camera frames (src/main.cpp:42-111, 591-625) ------------------------------
    def generate_rays(self, cam: np.ndarray, clip: float, width: int, height: int) -> np.ndarray:
        """gen_rays on the device, downloaded: (width*height,) rays in scan-line order."""
        cam = np.ascontiguousarray(cam, dtype="<f4")
        n = width * height
        d_rays = self.device_alloc(n * 32)
        try:
            self.lib.check(self.lib.dll.hgb_generate_rays(self._h, _ptr(cam), clip, width, height, d_rays), "generate_rays")
            return self.to_host(np.empty(n, dtype=RAY_DTYPE), d_rays)
        finally:
            self.device_free(d_rays)

    def render_frame(self, cam: np.ndarray, clip: float, width: int, height: int, mode: int, out: np.ndarray | None = None):
        """One viewer frame: (height, width, 4) BGRA bytes."""
        cam = np.ascontiguousarray(cam, dtype="<f4")
        if out is None:
            out = np.empty((height, width, 4), dtype=np.uint8)
        self.lib.check(self.lib.dll.hgb_render_frame(self._h, _ptr(cam), clip, width, height, mode, _ptr(out)), "render_frame")
        return out

    def bounce_rays_device(self, dev_rays: int, dev_hits: int, num_rays: int, offset: float, tmax: float, seed: int,
                           dev_out: int | None = None):
        """Second wave on the device (hgb_generate_bounce_rays): hits stay in HBM; in place when `dev_out` is None."""
        self.lib.check(self.lib.dll.hgb_generate_bounce_rays(self._h, dev_rays, dev_hits, num_rays, offset, tmax,
                                                             seed & 0xFFFFFFFF, dev_out if dev_out else dev_rays),
                       "generate_bounce_rays")

    def bounce_rays_keyed(self, dev_rays, dev_hits, num_rays: int, offset: float, tmax: float, seed: int, dev_keys, dev_out=None):
        """Second wave of a shard: `dev_keys` = the rays' indices in the whole frame (hgb_generate_bounce_rays_keyed)."""
        self.lib.check(self.lib.dll.hgb_generate_bounce_rays_keyed(self._h, _ptr(dev_rays), _ptr(dev_hits), num_rays, offset, tmax,
                                                                   seed & 0xFFFFFFFF, _ptr(dev_keys),
                                                                   _ptr(dev_out) if dev_out is not None else _ptr(dev_rays)),
                       "generate_bounce_rays_keyed")

    def count_hits(self, dev_hits, num_hits: int, dev_counters):
        """dev_counters (two uint64, device) += [hits with id >= 0, sum of id + 1] (hgb_count_hits)."""
        self.lib.check(self.lib.dll.hgb_count_hits(self._h, _ptr(dev_hits), num_hits, _ptr(dev_counters)), "count_hits")

    def trace_two_waves(self, dev_rays, num_rays: int, dev_keys, offset: float, tmax: float, seed: int, dev_hits_primary,
                        dev_bounce_rays, dev_hits_bounce, dev_counters=None):
        """One two-wave frame, everything resident (hgb_trace_two_waves)."""
        self.lib.check(self.lib.dll.hgb_trace_two_waves(self._h, _ptr(dev_rays), num_rays, _ptr(dev_keys), offset, tmax,
                                                        seed & 0xFFFFFFFF, _ptr(dev_hits_primary), _ptr(dev_bounce_rays),
                                                        _ptr(dev_hits_bounce), _ptr(dev_counters)), "trace_two_waves")

    def trace_two_waves_host(self, host_rays, num_rays: int, dev_keys, offset: float, tmax: float, seed: int, host_hits_primary,
                             host_hits_bounce):
        """One two-wave frame with host buffers (hgb_trace_two_waves_host)."""
        self.lib.check(self.lib.dll.hgb_trace_two_waves_host(self._h, _ptr(host_rays), num_rays, _ptr(dev_keys), offset, tmax,
                                                             seed & 0xFFFFFFFF, _ptr(host_hits_primary), _ptr(host_hits_bounce)),
                       "trace_two_waves_host")

    def bounce_rays(self, rays: np.ndarray, hits: np.ndarray, offset: float, tmax: float, seed: int) -> np.ndarray:
        """Host-array convenience over bounce_rays_device: `hits` are primitive-id hits of `rays`."""
        rays = np.ascontiguousarray(rays); hits = np.ascontiguousarray(hits)
        assert rays.dtype == RAY_DTYPE and hits.dtype == HIT_DTYPE and rays.shape[0] == hits.shape[0]
        n = rays.shape[0]
        if n == 0:
            return rays.copy()
        d_rays, d_hits = self.device_alloc(n * 32), self.device_alloc(n * 16)
        try:
            self.to_device(d_rays, rays); self.to_device(d_hits, hits)
            self.bounce_rays_device(d_rays, d_hits, n, offset, tmax, seed)
            return self.to_host(np.empty(n, dtype=RAY_DTYPE), d_rays)
        finally:
            self.device_free(d_rays); self.device_free(d_hits)

    # --- on-disk formats ------------------------------------------------------------
    def load_rays(self, path, tmin: float = 0.0, tmax: float = float(np.finfo(np.float32).max)) -> np.ndarray:
        """load_rays (src/main.cpp:277-300) to the device, downloaded again: (n,) rays."""
        n = int(self.lib.dll.hgb_rays_file_count(str(path).encode()))
        if n < 0:
            raise HagridError(f"cannot open {path}")
        d_rays = self.device_alloc(max(n, 1) * 32)
        try:
            got = self.lib.check(int(self.lib.dll.hgb_load_rays(self._h, str(path).encode(), tmin, tmax, d_rays)), "load_rays")
            assert got == n
            return self.to_host(np.empty(n, dtype=RAY_DTYPE), d_rays) if n else np.empty(0, dtype=RAY_DTYPE)
        finally:
            self.device_free(d_rays)

    def save_rays(self, path, rays: np.ndarray):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        d_rays = self.device_alloc(max(rays.nbytes, 32))
        try:
            self.to_device(d_rays, rays)
            self.lib.check(self.lib.dll.hgb_save_rays(self._h, str(path).encode(), d_rays, rays.shape[0]), "save_rays")
        finally:
            self.device_free(d_rays)

    def save_grid(self, path):
        self.lib.check(self.lib.dll.hgb_grid_save(self._h, str(path).encode()), "grid_save")

    def load_grid(self, path):
        self.lib.check(self.lib.dll.hgb_grid_load(self._h, str(path).encode()), "grid_load")

    # --- grid inspection / transplant -------------------------------------------
    def info(self) -> GridInfo:
        gi = GridInfo()
        self.lib.check(self.lib.dll.hgb_grid_get_info(self._h, C.byref(gi)), "grid_get_info")
        return gi

    def download(self):
        """Returns (info, entries u32[], cells Cell[]|SmallCell[], refs i32[]) in host memory."""
        gi = self.info()
        entries = np.empty(gi.num_entries, dtype="<u4")
        refs = np.empty(gi.num_refs, dtype="<i4")
        cells = np.empty(gi.num_cells, dtype=SMALL_CELL_DTYPE if gi.compressed else CELL_DTYPE)
        dl = self.lib.dll.hgb_grid_download
        self.lib.check(dl(self._h, ARRAY_ENTRIES, _ptr(entries), entries.nbytes), "download entries")
        self.lib.check(dl(self._h, ARRAY_SMALL_CELLS if gi.compressed else ARRAY_CELLS, _ptr(cells), cells.nbytes),
                       "download cells")
        if gi.num_refs:
            self.lib.check(dl(self._h, ARRAY_REFS, _ptr(refs), refs.nbytes), "download refs")
        return gi, entries, cells, refs

    def upload(self, gi: GridInfo, entries: np.ndarray, cells: np.ndarray, refs: np.ndarray):
        entries = np.ascontiguousarray(entries, dtype="<u4")
        refs = np.ascontiguousarray(refs, dtype="<i4")
        cells = np.ascontiguousarray(cells)
        assert cells.dtype == (SMALL_CELL_DTYPE if gi.compressed else CELL_DTYPE)
        assert entries.shape[0] == gi.num_entries and cells.shape[0] == gi.num_cells and refs.shape[0] == gi.num_refs
        self.lib.check(self.lib.dll.hgb_grid_upload(self._h, C.byref(gi), _ptr(entries), _ptr(cells), _ptr(refs)),
                       "grid_upload")

    def download_tris(self) -> np.ndarray:
        tris = np.empty(self.num_tris, dtype=TRI_DTYPE)
        self.lib.check(self.lib.dll.hgb_grid_download(self._h, ARRAY_TRIS, _ptr(tris), tris.nbytes), "download tris")
        return tris

    def peak_bytes(self) -> int:
        return int(self.lib.dll.hgb_scene_peak_bytes(self._h))

    # --- raw device buffers ------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = self.lib.dll.hgb_device_alloc(self._h, nbytes)
        if not p:
            raise HagridError("device_alloc failed")
        return int(p)

    def device_free(self, ptr: int):
        self.lib.dll.hgb_device_free(self._h, ptr)

    def to_device(self, dev_ptr: int, host: np.ndarray):
        host = np.ascontiguousarray(host)
        self.lib.check(self.lib.dll.hgb_copy_to_device(self._h, dev_ptr, _ptr(host), host.nbytes), "copy_to_device")

    def to_host(self, host: np.ndarray, dev_ptr: int):
        self.lib.check(self.lib.dll.hgb_copy_to_host(self._h, _ptr(host), dev_ptr, host.nbytes), "copy_to_host")
        return host

    def trace(self, rays: np.ndarray, hit_mode: int = HIT_PRIM_ID) -> np.ndarray:
        """Convenience for tests: device buffers from the scene's pool, one launch, hits back."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        n = rays.shape[0]
        d_rays = self.device_alloc(max(rays.nbytes, 32))
        d_hits = self.device_alloc(max(n * 16, 16))
        try:
            self.to_device(d_rays, rays)
            self.traverse(d_rays, d_hits, n, hit_mode)
            return self.to_host(np.empty(n, dtype=HIT_DTYPE), d_hits)
        finally:
            self.device_free(d_rays)
            self.device_free(d_hits)
