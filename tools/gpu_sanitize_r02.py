"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck) over the kernels added late in
round 2: merge_grid in one cooperative launch, the tile kernel with its ticket list and parts, the side-stream kernels
that make the list. usage (under gpurun): compute-sanitizer --tool memcheck python tools/gpu_sanitize_r02.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes
lib = Library()
tris = scenes.hairball(6000, seed=4)
sc = Scene(tris, keep_alive=True, lib=lib)
sc.build_all(0.12, 2.4, 0.995, 3, False)
sc.setup_traversal()
rays = scenes.default_view(tris, 328, 203)           # ragged raster
n = rays.shape[0]
lib.set_option("traverse_variant", 0)
want = sc.trace(rays, HIT_PRIM_ID)
lib.set_option("traverse_variant", 4)
d_rays, d_hits = sc.device_alloc(n * 32), sc.device_alloc(n * 16)
sc.to_device(d_rays, rays)
for setting in ((8, 256, 2, 0), (8, 64, 5, 0), (4, 16, 1, 50)):
    for key, val in zip(("tile_order", "tile_split", "tile_split_log", "tile_split_share"), setting):
        lib.set_option(key, val)
    for launch in range(5):
        sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
        got = sc.to_host(np.empty(n, dtype=want.dtype), d_hits)
        assert np.array_equal(got["id"], want["id"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), (setting, launch)
sc.close()
print("ok", n, "rays")
