"""Where a construction goes: each stage of C2 / C4 timed on its own (CUDA events around the C ABI's stage calls, after
warm-up builds in keep-alive mode), kernel launches per stage. (gpurun)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
lib = Library()
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2"]):
    tris, td, sd = {"c2": (scenes.sponza262k, 0.15, 3.0), "c4": (scenes.hairball, 0.12, 2.4), "c5": (scenes.sanmiguel7p8m, 0.15, 3.0)}[name]
    tris = tris()
    sc = Scene(tris, keep_alive=True, lib=lib)
    sc.build_all(td, sd, 0.995, 3, False, warmup=4, iters=0)
    stages = [("build", lambda: sc.build_grid(td, sd)), ("merge", lambda: sc.merge_grid(0.995)), ("flatten", sc.flatten_grid), ("expand", lambda: sc.expand_grid(3))]
    acc = {k: [] for k, _ in stages}; launches = {k: 0 for k, _ in stages}
    for rep in range(8):
        for k, fn in stages:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = lib.kernel_launches()
            a.record(); fn(); b.record(); b.synchronize()
            acc[k].append(a.elapsed_time(b)); launches[k] = lib.kernel_launches() - l0
    print(name, " ".join(f"{k} {np.mean(v[2:]):.3f} ms ({launches[k]} launches)" for k, v in acc.items()), "total %.3f" % sum(np.mean(v[2:]) for v in acc.values()), flush=True)
    sc.close()
