"""Traversal micro-benchmark on a reference-built grid (run under gpurun, optionally under ncu).
usage: gpu_traverse_bench.py [primary|long|random] [iters] [variants csv] [ref]
HGB_COMPRESS=0|1 overrides the default (compressed grid for random rays only); HGB_FRAME=WxH sets the size of the primary view."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "primary"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")]
with_ref = len(sys.argv) > 4 and sys.argv[4] == "ref"

ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
mine = Library()
tris = scenes.sponza262k()
import os
compress = kind == "random" if "HGB_COMPRESS" not in os.environ else os.environ["HGB_COMPRESS"] == "1"
W_, H_ = (int(v) for v in os.environ.get("HGB_FRAME", "1920x1080").split("x"))
rays = {"primary": lambda: scenes.default_view(tris, W_, H_), "long": lambda: scenes.default_view(tris, along_long_axis=True),
        "random": lambda: scenes.random_rays(tris, 4194304)}[kind]()
sr = Scene(tris, lib=ref)
sr.build_grid(0.15, 3.0); sr.merge_grid(0.995); sr.flatten_grid(); sr.expand_grid(3)
if compress:
    sr.compress_grid()
sr.setup_traversal()
gi, e, c, r = sr.download()
sm = Scene(tris, lib=mine)
sm.upload(gi, e, c, r)
sm.setup_traversal()
n = rays.shape[0]
out = {}
ref_hits = None
for label, sc, lib_, vs in (("ref", sr, ref, [0] if with_ref else []), ("mine", sm, mine, variants)):
    if not vs:
        continue
    d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16)
    sc.to_device(d_rays, rays)
    for v in vs:
        lib_.set_option("traverse_variant", v)
        ms = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=min(3, iters), iters=iters)
        hits = sc.to_host(np.empty(n, dtype=np.dtype([("id", "<i4"), ("t", "<u4"), ("u", "<f4"), ("v", "<f4")])), d_hits)
        if label == "ref": ref_hits = hits
        same = None if ref_hits is None else int(((hits["id"] == ref_hits["id"]) & (hits["t"] == ref_hits["t"])).sum())
        out[f"{label}_v{v}"] = {"ms_median": round(float(np.median(ms)), 4), "ms_min": round(float(ms.min()), 4),
                                "mrays_s": round(float(n * len(ms) / (1000.0 * ms.sum())), 1), "identical_hits": same, "n": n}
print(json.dumps(out))
