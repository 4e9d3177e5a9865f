"""BASELINE.json C5 on one GPU (run under gpurun): 7.8 M-triangle scene, 1920x1080 primary rays + one bounce;
construction and traversal of the reference (rebuilt for sm_100a) and of this library, same buffers."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
tris = scenes.sanmiguel7p8m()
primary = scenes.default_view(tris)
out = {"triangles": int(tris.shape[0])}
bounce = None
hits = {}
for label, lib in (("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so")), ("hagrid_b200", Library())):
    sc = Scene(tris, keep_alive=True, lib=lib)
    ms = sc.build_all(0.15, 3.0, 0.995, 3, False, warmup=2, iters=5)
    out[f"{label}_build_ms"] = {"mean": round(float(ms.mean()), 2), "min": round(float(ms.min()), 2)}
    sc.setup_traversal()
    if bounce is None:
        first = sc.trace(primary, HIT_PRIM_ID)
        bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
        out["grid"] = {k: sc.info().as_dict()[k] for k in ("dims", "shift", "num_cells", "num_entries", "num_refs")}
    for name, rays in (("primary", primary), ("bounce", bounce)):
        n = rays.shape[0]
        d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16); sc.to_device(d_rays, rays)
        t = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=3, iters=20)
        h = sc.to_host(np.empty(n, dtype=np.dtype([("id", "<i4"), ("t", "<u4"), ("u", "<f4"), ("v", "<f4")])), d_hits)
        hits[(label, name)] = h
        out[f"{label}_{name}_mrays_s"] = round(float(n * len(t) / (1000.0 * t.sum())), 1)
        sc.device_free(d_rays); sc.device_free(d_hits)
    sc.close()
for name in ("primary", "bounce"):
    a, b = hits[("reference", name)], hits[("hagrid_b200", name)]
    out[f"identical_{name}"] = bool(np.array_equal(a["id"], b["id"]) and np.array_equal(a["t"], b["t"]))
print(json.dumps(out))
