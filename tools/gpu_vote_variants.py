"""Blocks per SM of the voting kernel (build-time variants under hagrid_b200/_build/variants/libhagrid_b200_vote*.so),
on the second wave of C5 and on C3's random rays (run under gpurun)."""
import sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
VAR = ROOT / "hagrid_b200" / "_build" / "variants"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=12, warmup=3):
    for _ in range(warmup):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]; b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_(); a[i].record(); fn(); b[i].record()
    torch.cuda.synchronize()
    ms = np.array([x.elapsed_time(y) for x, y in zip(a, b)])
    return round(float(ms.mean()), 4), round(float(ms.min()), 4)
libs = [("default", Library())] + [(p.stem.replace("libhagrid_b200_", ""), Library(p)) for p in sorted(VAR.glob("libhagrid_b200_vote*.so"))]
for scene_name in ("sanmiguel", "sponza"):
    tris = scenes.sanmiguel7p8m() if scene_name == "sanmiguel" else scenes.sponza262k()
    want = None
    rays = None
    for tag, lib in libs:
        sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0, 0.995, 3, scene_name == "sponza"); sc.setup_traversal()
        if rays is None:
            if scene_name == "sanmiguel":
                primary = scenes.default_view(tris)
                first = sc.trace(primary, HIT_PRIM_ID)
                rays = scenes.bounce_rays(tris, primary, first["id"], first["t"])
            else:
                rays = scenes.random_rays(tris, 1 << 22)
        n = rays.shape[0]
        d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(); d_hits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        m, lo = timed(lambda: sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID))
        got = d_hits.cpu().numpy().view(np.uint32)[:, :2].copy()
        want = got if want is None else want
        print(f"{scene_name:9s} {tag:10s} mean {m:.4f} min {lo:.4f} identical {bool(np.array_equal(got, want))}", flush=True)
        sc.close()
