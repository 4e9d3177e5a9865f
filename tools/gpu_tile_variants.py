"""Build-time variants of the tile kernel (blocks per SM x unroll factor of the reference loop), C2 primary + long view.
  python tools/gpu_tile_variants.py build "12,1 12,2 10,1"     (here: nvcc, no GPU needed; libraries under hagrid_b200/_build/variants/)
  python tools/gpu_tile_variants.py run                          (under gpurun)"""
import json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VAR = ROOT / "hagrid_b200" / "_build" / "variants"

if sys.argv[1] == "build":
    from hagrid_b200 import build as B
    B.build_library()
    VAR.mkdir(parents=True, exist_ok=True)
    for old in VAR.glob("*.so"):
        old.unlink()
    others = [str(o) for o in sorted(B.OBJ.glob("*.o")) if o.stem != "ray_traverse"]
    for cfg in sys.argv[2].split():
        defs = cfg.split(",")
        blocks, unroll = defs[0], defs[1]
        extra = [f"-D{d}" for d in defs[2:]]
        tag = cfg.replace(",", "_").replace("=", "")
        obj = VAR / f"ray_traverse_{tag}.o"
        B._run([B.NVCC] + B.NVCC_FLAGS + [f"-DHGB_TILE_BLOCKS={blocks}", f"-DHGB_REF_UNROLL={unroll}"] + extra +
               ["-c", str(B.CSRC / "ray_traverse.cu"), "-o", str(obj)], VAR / f"{tag}.ptxas.log")
        B._run(["g++", "-shared", "-o", str(VAR / f"libhagrid_b200_{tag}.so"), str(obj)] + others +
               ["-Wl,-Bsymbolic", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
        obj.unlink()
        print("built", tag)
    sys.exit(0)

import numpy as np, torch
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, iters=40, warmup=5):
    for _ in range(warmup):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]; b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_(); a[i].record(); fn(); b[i].record()
    torch.cuda.synchronize()
    ms = np.array([x.elapsed_time(y) for x, y in zip(a, b)])
    return float(ms.mean()), float(ms.min())

scene_names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sponza"]
libs = [("default", Library())] + [(p.stem.replace("libhagrid_b200_", ""), Library(p)) for p in sorted(VAR.glob("*.so"))]
ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
for scene_name in scene_names:
    tris = scenes.sponza262k() if scene_name == "sponza" else scenes.sanmiguel7p8m()
    views = {"primary": scenes.default_view(tris)}
    if scene_name == "sponza":
        views["long"] = scenes.default_view(tris, along_long_axis=True)
    sr = Scene(tris, keep_alive=True, lib=ref); sr.build_all(0.15, 3.0); sr.setup_traversal()
    want = {}
    for vname, rays in views.items():
        n = rays.shape[0]
        d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        m, lo = timed(lambda: sr.traverse(d_rays, d_hits, n, HIT_PRIM_ID))
        want[vname] = (d_hits.cpu().numpy().copy(), m)
        print(f"{scene_name:9s} {vname:8s} reference        mean {m:.4f} min {lo:.4f}", flush=True)
    sr.close()
    for tag, lib in libs:
        sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
        for vname, rays in views.items():
            n = rays.shape[0]
            d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(); d_hits = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
            m, lo = timed(lambda: sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID))
            ok = bool(np.array_equal(d_hits.cpu().numpy().view(np.uint32)[:, :2], want[vname][0].view(np.uint32)[:, :2]))
            print(f"{scene_name:9s} {vname:8s} {tag:16s} mean {m:.4f} min {lo:.4f}  x{want[vname][1] / m:.3f} identical {ok}", flush=True)
        sc.close()
