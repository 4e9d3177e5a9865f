"""Build-time alternatives of construction kernels, measured on a whole build and per kernel under ncu.
  python tools/gpu_build_variants.py build          here: libhagrid_b200_classify_staged.so under hagrid_b200/_build/variants/
  python tools/gpu_build_variants.py run [c4|c2]    under gpurun: build times of the default library and of the alternative
  ncu ... python tools/gpu_build_variants.py profile c4 <tag>   two builds of one library (for a per-kernel capture)"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VAR = ROOT / "hagrid_b200" / "_build" / "variants"
from hagrid_b200 import build as B
if sys.argv[1] == "build":
    B.build_library()
    VAR.mkdir(parents=True, exist_ok=True)
    others = [str(o) for o in sorted(B.OBJ.glob("*.o")) if o.stem != "grid_build"]
    obj = VAR / "grid_build_classify_staged.o"
    B._run([B.NVCC] + B.NVCC_FLAGS + ["-DHGB_CLASSIFY_STAGED", "-c", str(B.CSRC / "grid_build.cu"), "-o", str(obj)], VAR / "classify_staged.ptxas.log")
    B._run(["g++", "-shared", "-o", str(VAR / "libhagrid_b200_classify_staged.so"), str(obj)] + others +
           ["-Wl,-Bsymbolic", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    obj.unlink()
    print("built"); sys.exit(0)
import numpy as np
from hagrid_b200 import Library, Scene, scenes
which = sys.argv[2] if len(sys.argv) > 2 else "c4"
tris, td, sd = (scenes.hairball(), 0.12, 2.4) if which == "c4" else (scenes.sponza262k(), 0.15, 3.0)
libs = {"default": Library(), "classify_staged": Library(VAR / "libhagrid_b200_classify_staged.so")}
if sys.argv[1] == "profile":
    lib = libs[sys.argv[3]]
    sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(td, sd, 0.995, 3, False, warmup=1, iters=1); sc.close(); sys.exit(0)
ref = None
for tag, lib in libs.items():
    sc = Scene(tris, keep_alive=True, lib=lib)
    ms = sc.build_all(td, sd, 0.995, 3, False, warmup=5, iters=10)
    gi, e, c, r = sc.download()
    blob = (e.tobytes(), c.tobytes(), r.tobytes())
    ref = ref or blob
    print(f"{which} {tag:16s} build mean {ms.mean():.3f} ms min {ms.min():.3f}  identical grid {blob == ref}", flush=True)
    sc.close()
