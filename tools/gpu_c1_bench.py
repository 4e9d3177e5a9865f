"""BASELINE.json C1 (run under gpurun): Cornell-box-scale scene (32 triangles), 256x256 primary rays, default densities."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes
tris = scenes.cornell32(); rays = scenes.cornell_view(256, 256); n = rays.shape[0]
out = {}; hits = {}
for label, lib in (("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so")), ("hagrid_b200", Library())):
    sc = Scene(tris, keep_alive=True, lib=lib)
    ms = sc.build_all(0.12, 2.4, 0.995, 3, False, warmup=3, iters=10)
    sc.setup_traversal()
    d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16); sc.to_device(d_rays, rays)
    if label == "hagrid_b200":
        for v in (0, 2, 4, 1):
            lib.set_option("traverse_variant", v)
            t = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=5, iters=50)
            out[f"variant{v}_us"] = round(float(np.median(t)) * 1e3, 2)
        lib.set_option("traverse_variant", 3)
    for mode, name in ((HIT_STEPS, "steps"), (HIT_PRIM_ID, "ids")):
        t = sc.traverse_timed(d_rays, d_hits, n, mode, warmup=5, iters=50)
        hits[(label, name)] = sc.to_host(np.empty(n, dtype=np.dtype([("id", "<i4"), ("t", "<u4"), ("u", "<f4"), ("v", "<f4")])), d_hits)
        out[f"{label}_{name}"] = {"us_median": round(float(np.median(t)) * 1e3, 2), "mrays_s": round(float(n * len(t) / (1000.0 * t.sum())), 1)}
    out[f"{label}_build_ms"] = round(float(ms.mean()), 3)
    sc.close()
out["identical"] = all(bool(np.array_equal(hits[("reference", k)], hits[("hagrid_b200", k)])) for k in ("steps", "ids"))
print(json.dumps(out))
