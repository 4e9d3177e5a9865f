"""Construction times (build+merge+flatten+expand, keep-alive, event-timed like src/main.cpp:494-508) of C2, C4 and C5,
kernel launches per build, reference alongside when oracle/_ref is there. (gpurun)"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2", "c4"]
libs = [("hagrid_b200", Library())]
if (ROOT / "oracle/_ref/libhagrid_ref.so").exists() and "noref" not in sys.argv:
    libs.append(("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so")))
out = {}
for name in which:
    tris, td, sd = {"c2": (scenes.sponza262k, 0.15, 3.0), "c4": (scenes.hairball, 0.12, 2.4), "c5": (scenes.sanmiguel7p8m, 0.15, 3.0),
                    "c1": (scenes.cornell32, 0.12, 2.4)}[name]
    tris = tris()
    for label, lib in libs:
        sc = Scene(tris, keep_alive=True, lib=lib)
        sc.build_all(td, sd, 0.995, 3, False, warmup=5, iters=0)
        l0 = lib.kernel_launches()
        ms = sc.build_all(td, sd, 0.995, 3, False, warmup=0, iters=10)
        out[f"{name}_{label}"] = {"mean_ms": round(float(ms.mean()), 3), "min_ms": round(float(ms.min()), 3),
                                  "launches_per_build": (lib.kernel_launches() - l0) / 10.0}
        print(name, label, out[f"{name}_{label}"], flush=True)
        sc.close()
