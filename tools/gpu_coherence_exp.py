"""Experiment: how much does ray ordering buy? (host-side permutations, timing only)"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes

ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so"); mine = Library()
tris = scenes.sponza262k()

def grid(compress):
    sr = Scene(tris, lib=ref)
    sr.build_grid(0.15, 3.0); sr.merge_grid(0.995); sr.flatten_grid(); sr.expand_grid(3)
    if compress: sr.compress_grid()
    sr.setup_traversal()
    gi, e, c, r = sr.download()
    sm = Scene(tris, lib=mine); sm.upload(gi, e, c, r); sm.setup_traversal()
    return sr, sm, gi

def timeit(sc, lib_, rays, variant):
    n = rays.shape[0]
    d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16)
    sc.to_device(d_rays, rays)
    lib_.set_option("traverse_variant", variant)
    ms = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=5, iters=20)
    sc.device_free(d_rays); sc.device_free(d_hits)
    return round(float(np.median(ms)), 4)

def tile_perm(w, h, tw, th):
    idx = np.arange(w * h).reshape(h, w)
    return idx.reshape(h // th, th, w // tw, tw).transpose(0, 2, 1, 3).reshape(-1)

out = {}
sr, sm, gi = grid(False)
for name, rays in (("primary", scenes.default_view(tris)), ("long", scenes.default_view(tris, along_long_axis=True))):
    res = {"raster": {"ref": timeit(sr, ref, rays, 0), "v0": timeit(sm, mine, rays, 0), "v1": timeit(sm, mine, rays, 1)}}
    for tw, th in ((8, 4), (4, 8), (16, 2), (8, 8), (16, 8)):
        p = rays[tile_perm(1920, 1080, tw, th)] if 1080 % th == 0 else None
        if p is None: continue
        res[f"tile{tw}x{th}"] = {"ref": timeit(sr, ref, p, 0), "v0": timeit(sm, mine, p, 0), "v1": timeit(sm, mine, p, 1)}
    out[name] = res
    print(name, json.dumps(res), flush=True)
sr.close(); sm.close()

sr, sm, gi = grid(True)
rays = scenes.random_rays(tris, 4194304)
lo = np.array(gi.bbox_min[:], np.float32); hi = np.array(gi.bbox_max[:], np.float32)
res = {"random": {"ref": timeit(sr, ref, rays, 0), "v0": timeit(sm, mine, rays, 0), "v1": timeit(sm, mine, rays, 1)}}
q = np.clip(((rays["org"] - lo) / (hi - lo) * 32).astype(np.int64), 0, 31)
def morton3(q):
    m = np.zeros(q.shape[0], np.int64)
    for b in range(5):
        for a in range(3):
            m |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return m
octant = (rays["dir"][:, 0] < 0).astype(np.int64) | ((rays["dir"][:, 1] < 0).astype(np.int64) << 1) | ((rays["dir"][:, 2] < 0).astype(np.int64) << 2)
for label, key in (("oct_morton", (octant << 15) | morton3(q)), ("morton_oct", (morton3(q) << 3) | octant), ("morton", morton3(q))):
    p = rays[np.argsort(key, kind="stable")]
    res[label] = {"ref": timeit(sr, ref, p, 0), "v0": timeit(sm, mine, p, 0), "v1": timeit(sm, mine, p, 1)}
out["random"] = res
print("random", json.dumps(res), flush=True)
(ROOT / "gpurun_out/coherence.json").write_text(json.dumps(out, indent=1))
