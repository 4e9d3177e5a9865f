"""Golden vectors of the reference's front-end host functions (gen_camera, gen_rays, update_surface,
src/main.cpp:42-111), produced by the reference's OWN code: oracle/ref_frontend.cpp includes src/main.cpp
unmodified and oracle/build_ref.sh links it into oracle/_ref/libhagrid_ref.so. Pure CPU work: run it in
the build container (needs /root/reference for the build), commit tests/golden/frontend.npz."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
HIT = np.dtype([("id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
RAY = np.dtype([("org", "<f4", 3), ("tmin", "<f4"), ("dir", "<f4", 3), ("tmax", "<f4")])

dll = C.CDLL(str(ROOT / "oracle" / "_ref" / "libhagrid_ref.so"))
dll.hgb_ref_gen_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
dll.hgb_ref_gen_rays.restype = C.c_void_p
dll.hgb_ref_gen_rays.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
dll.hgb_ref_update_surface.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]

cases = [  # eye, center, up, fov, w, h, clip
    ((278, 273, -800), (278, 273, 0), (0, 1, 0), 60.0, 64, 48, 1900.0),
    ((0.3, 11.7, -0.2), (5.1, 9.0, 4.4), (0, 1, 0), 60.0, 72, 40, 71.25),
    ((-3.5, 0.25, 9.0), (1.0, 2.0, -4.0), (0.1, 0.9, 0.2), 37.5, 40, 24, 33.0),
]
rng = np.random.default_rng(20261017)
out = {"cases": np.array([[*e, *c, *u, f, w, h, clip] for e, c, u, f, w, h, clip in cases], dtype=np.float64)}
for k, (eye, center, up, fov, w, h, clip) in enumerate(cases):
    e, c, u = (np.array(v, dtype="<f4") for v in (eye, center, up))
    cam = np.empty(12, dtype="<f4")
    dll.hgb_ref_gen_camera(e.ctypes.data, c.ctypes.data, u.ctypes.data, fov, w / h, cam.ctypes.data)
    ptr = dll.hgb_ref_gen_rays(cam.ctypes.data, clip, w, h)
    rays = np.frombuffer((C.c_char * (32 * w * h)).from_address(ptr), dtype=RAY).copy()
    hits = np.zeros(w * h, dtype=HIT)
    hits["id"] = rng.integers(0, 140, w * h)                       # step counts, some beyond the 100 / 255 clamps
    hits["id"][:5] = (0, 99, 100, 101, 300)
    hits["t"] = (rng.random(w * h) * clip).astype("<f4")
    hits["t"][:4] = (0.0, clip, clip * 0.5, np.nextafter(np.float32(clip), np.float32(0)))
    out[f"cam{k}"] = cam; out[f"rays{k}"] = rays; out[f"hits{k}"] = hits
    for mode in (0, 1, 2):
        img = np.empty((h, w, 4), dtype=np.uint8)
        dll.hgb_ref_update_surface(mode, hits.ctypes.data, clip, w, h, img.ctypes.data)
        out[f"image{k}_{mode}"] = img
np.savez_compressed(Path(__file__).with_name("frontend.npz"), **out)
print("wrote", Path(__file__).with_name("frontend.npz"))
