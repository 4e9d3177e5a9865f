"""Interactive-frame comparison (run under gpurun): one viewer frame = primary rays of a camera -> BGRA image in
host memory. Reference = its own loop (CPU gen_rays, upload, traverse_grid, download, CPU update_surface,
src/main.cpp:598-621) through the reference build of hgb_render_frame; ours = one fused launch + 4 B/pixel download."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, make_camera, scenes
W, H = 1920, 1080
tris = scenes.sponza262k()
lo, hi = scenes.scene_bbox(tris); eye = 0.5 * (lo + hi); clip = float(np.linalg.norm(hi - lo))
out = {"frame": f"{W}x{H}, C2 scene, default view"}
images = {}
for label, lib in (("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so")), ("hagrid_b200", Library())):
    sc = Scene(tris, keep_alive=True, lib=lib)
    sc.build_all(0.15, 3.0); sc.setup_traversal()
    cam = make_camera(eye, eye + np.array([0, 0, 1], np.float32), (0, 1, 0), 60.0, W / H, lib=lib)
    img = np.empty((H, W, 4), np.uint8)
    for mode in (0, 2):
        for _ in range(3): sc.render_frame(cam, clip, W, H, mode, img)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter(); sc.render_frame(cam, clip, W, H, mode, img); ts.append((time.perf_counter() - t0) * 1e3)
        out[f"{label}_mode{mode}_ms"] = {"median": round(float(np.median(ts)), 3), "min": round(min(ts), 3)}
        images[(label, mode)] = img.copy()
    sc.close()
for mode in (0, 2):
    out[f"identical_mode{mode}"] = bool(np.array_equal(images[("reference", mode)], images[("hagrid_b200", mode)]))
print(json.dumps(out))
