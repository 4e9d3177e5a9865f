"""How much of a tile launch is fixed cost (ramp-up + tail)? The C2 frame traced once, twice and four times in one launch
(the ray buffer repeated: still a raster of the same width), time per frame = (T(k) - T(1)) / (k - 1). (gpurun)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=30, warmup=4):
    for _ in range(warmup):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]; b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_(); a[i].record(); fn(); b[i].record()
    torch.cuda.synchronize()
    return float(np.mean([x.elapsed_time(y) for x, y in zip(a, b)]))
tris = scenes.sponza262k()
rays = scenes.default_view(tris)
for name, path in (("reference", ROOT / "oracle/_ref/libhagrid_ref.so"), ("hagrid_b200", None)):
    lib = Library(path) if path else Library()
    sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
    res = {}
    for k in (1, 2, 4):
        buf = np.concatenate([rays] * k); n = buf.shape[0]
        d_rays = torch.from_numpy(buf.view(np.float32).reshape(n, 8)).cuda(); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        res[k] = timed(lambda: sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID))
    per = (res[4] - res[1]) / 3
    print(f"{name:12s} T1 {res[1]:.4f} T2 {res[2]:.4f} T4 {res[4]:.4f}  per extra frame {per:.4f}  fixed {res[1] - per:.4f} ms", flush=True)
    sc.close()
