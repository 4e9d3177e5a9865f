"""Build-time alternatives of construction kernels, measured on a whole build and per kernel under ncu.
  python tools/gpu_build_variants.py build          here: libhagrid_b200_classify_staged.so under hagrid_b200/_build/variants/
  python tools/gpu_build_variants.py build grid_merge "r256x4:HGB_ROUND_BLOCK=256,HGB_ROUND_BLOCKS_PER_SM=4 r1024x1:..."   any source, any defines
  python tools/gpu_build_variants.py run [c4|c2]    under gpurun: build times of the default library and of the alternative
  ncu ... python tools/gpu_build_variants.py profile c4 <tag>   two builds of one library (for a per-kernel capture)"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VAR = ROOT / "hagrid_b200" / "_build" / "variants"
from hagrid_b200 import build as B
if sys.argv[1] == "build" and len(sys.argv) > 3:
    B.build_library()
    VAR.mkdir(parents=True, exist_ok=True)
    stem = sys.argv[2]
    others = [str(o) for o in sorted(B.OBJ.glob("*.o")) if o.stem != stem]
    for item in sys.argv[3].split():
        tag, _, defs = item.partition(":")
        obj = VAR / f"{stem}_{tag}.o"
        B._run([B.NVCC] + B.NVCC_FLAGS + [f"-D{d}" for d in defs.split(",") if d] + ["-c", str(B.CSRC / f"{stem}.cu"), "-o", str(obj)], VAR / f"{tag}.ptxas.log")
        B._run(["g++", "-shared", "-o", str(VAR / f"libhagrid_b200_{tag}.so"), str(obj)] + others +
               ["-Wl,-Bsymbolic", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
        obj.unlink()
        print("built", tag)
    sys.exit(0)
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
libs = {"default": Library()}
libs.update({p.stem.replace("libhagrid_b200_", ""): Library(p) for p in sorted(VAR.glob("libhagrid_b200_*.so"))})
if sys.argv[1] == "profile":
    lib = libs[sys.argv[3]]
    sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(td, sd, 0.995, 3, False, warmup=1, iters=1); sc.close(); sys.exit(0)
ref = None
for tag, lib in libs.items():
    if tag == "default" and len(sys.argv) > 3:
        lib.set_option("merge_one_launch_max_cells", int(sys.argv[3]))      # e.g. 0: the pass-per-launch path as the base line
    sc = Scene(tris, keep_alive=True, lib=lib)
    ms = sc.build_all(td, sd, 0.995, 3, False, warmup=5, iters=10)
    gi, e, c, r = sc.download()
    blob = (e.tobytes(), c.tobytes(), r.tobytes())
    ref = ref or blob
    print(f"{which} {tag:16s} build mean {ms.mean():.3f} ms min {ms.min():.3f}  identical grid {blob == ref}", flush=True)
    sc.close()
