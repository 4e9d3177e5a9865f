"""Dynamic scene on one GPU (run under gpurun): every frame the triangles change, the grid is
rebuilt with the keep-alive allocator and the frame is rendered -- upload triangles, build_grid ... expand_grid,
setup_traversal, one viewer frame. Same loop through the same C ABI on the reference (rebuilt for sm_100a) and on
this library; images compared frame by frame; host clock around whole frames."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, make_camera, scenes

W, H, FRAMES, LAPS = 1920, 1080, 8, 4
base = scenes.sponza262k()
poses = [scenes.animate(base, 2 * np.pi * f / FRAMES) for f in range(FRAMES)]
lo, hi = scenes.scene_bbox(base)
center = 0.5 * (lo + hi)
clip = float(np.linalg.norm(hi - lo))
out = {"scene": "sponza262k, all vertices animated", "triangles": int(base.shape[0]), "frame": [W, H], "poses": FRAMES}
images = {}
for label, lib in (("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so")), ("hagrid_b200", Library())):
    cam = make_camera(center, center + np.array([0, 0, 1], np.float32), (0, 1, 0), 60.0, W / H, lib=lib)
    sc = Scene(poses[0], keep_alive=True, lib=lib)
    img = np.empty((H, W, 4), np.uint8)
    part = {"upload": 0.0, "build": 0.0, "frame": 0.0}

    def frame(tris, acc=None):
        t0 = time.perf_counter()
        sc.set_tris(tris)
        t1 = time.perf_counter()
        sc.build_all(0.15, 3.0)
        sc.setup_traversal()
        lib.synchronize()
        t2 = time.perf_counter()
        sc.render_frame(cam, clip, W, H, 2, img)
        t3 = time.perf_counter()
        if acc is not None:
            acc["upload"] += t1 - t0; acc["build"] += t2 - t1; acc["frame"] += t3 - t2

    for tris in poses:                       # warm-up lap, keeps the images for the comparison
        frame(tris)
        images.setdefault(label, []).append(img.copy())
    peak = sc.peak_bytes()
    t0 = time.perf_counter()
    for _ in range(LAPS):
        for tris in poses:
            frame(tris, part)
    dt = (time.perf_counter() - t0) / (LAPS * FRAMES)
    out[label] = {"ms_per_frame": round(dt * 1e3, 3), "fps": round(1 / dt, 1),
                  "upload_ms": round(part["upload"] * 1e3 / (LAPS * FRAMES), 3),
                  "rebuild_ms": round(part["build"] * 1e3 / (LAPS * FRAMES), 3),
                  "render_ms": round(part["frame"] * 1e3 / (LAPS * FRAMES), 3),
                  "pool_grew_after_first_lap": bool(sc.peak_bytes() > peak)}
    sc.close()
out["images_identical"] = all(np.array_equal(a, b) for a, b in zip(images["reference"], images["hagrid_b200"]))
print(json.dumps(out))
