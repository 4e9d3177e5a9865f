"""Debug aid: host-buffer frames one by one with progress on stderr (run under gpurun + timeout)."""
import sys, time, faulthandler
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes
faulthandler.dump_traceback_later(60, exit=True)
lib = Library()
tris = scenes.sponza262k() if len(sys.argv) < 2 else scenes.small_mixed(20000)
sc = Scene(tris, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
lo, hi = scenes.scene_bbox(tris); eye = 0.5 * (lo + hi)
frames = {"raster720": scenes.primary_rays(eye, eye + np.array([0.3, 0.0, 1.0], np.float32), (0, 1, 0), 60.0, 1280, 720, 1e4),
          "random700001": scenes.random_rays(tris, 700001, seed=12),
          "raster64x8": scenes.primary_rays(eye, eye + np.array([0.0, 0.1, 1.0], np.float32), (0, 1, 0), 60.0, 64, 8, 1e4),
          "random77": scenes.random_rays(tris, 77, seed=13)}
for name, rays in frames.items():
    want = sc.trace(rays, HIT_PRIM_ID)
    for mode in (HIT_PRIM_ID, HIT_STEPS):
        print("frame", name, "mode", mode, "...", file=sys.stderr, flush=True)
        t0 = time.time()
        got = sc.traverse_host(rays, mode)
        print("   done in %.3f s" % (time.time() - t0), "ids equal:", bool(np.array_equal(got["id"], want["id"])) if mode == HIT_PRIM_ID else "-", file=sys.stderr, flush=True)
print("ALL DONE", file=sys.stderr)
