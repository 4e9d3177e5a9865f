import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
which = sys.argv[1] if len(sys.argv) > 1 else "sanmiguel"
tris = scenes.sanmiguel7p8m() if which == "sanmiguel" else scenes.sponza262k()
primary = scenes.default_view(tris)
lib = Library()
sc = Scene(tris, keep_alive=True, lib=lib)
ms = sc.build_all(0.15, 3.0, 0.995, 3, False, warmup=1, iters=6)
out = {"build_ms_each": [round(float(x), 1) for x in ms]}
sc.setup_traversal()
first = sc.trace(primary, HIT_PRIM_ID)
bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
for name, rays in (("primary", primary), ("bounce", bounce)):
    n = rays.shape[0]
    d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16); sc.to_device(d_rays, rays)
    for v in (0, 1, 2, 4, 3):
        lib.set_option("traverse_variant", v)
        t = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=3, iters=15)
        out[f"{name}_v{v}"] = round(float(n * len(t) / (1000.0 * t.sum())), 1)
    sc.device_free(d_rays); sc.device_free(d_hits)
print(json.dumps(out))
