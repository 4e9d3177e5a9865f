"""Builds libhagrid_b200.so (hand-written sm_100a kernels + the C ABI) in-tree.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the gpurun snapshot.  Usage: python -m hagrid_b200.build
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "_build"
LIB = PKG / "libhagrid_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "--use_fast_math", "-lineinfo", "--expt-relaxed-constexpr",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
    f"-I{PKG / 'include' / 'hagrid'}", f"-I{CSRC}", f"-I{ROOT / 'include'}",
]
CXX_FLAGS = [
    "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-DHOST=", "-DDEVICE=",
    "-I/usr/local/cuda/include",
    f"-I{PKG / 'include' / 'hagrid'}", f"-I{CSRC}", f"-I{ROOT / 'include'}",
]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps)


def _run(cmd, log: Path | None = None):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("build failed: " + " ".join(map(str, cmd)))
    return res.stdout


def build_library(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list((PKG / "include" / "hagrid").glob("*.h")) + list(CSRC.glob("*.h")) + \
        list(CSRC.glob("*.cuh")) + [ROOT / "include" / "hagrid_b200.h"]
    jobs = []
    objs = []
    for src in sorted(CSRC.glob("*.cu")):
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            jobs.append(([NVCC] + NVCC_FLAGS + ["-c", str(src), "-o", str(obj)], OBJ / (src.stem + ".ptxas.log")))
    for src in sorted(CSRC.glob("*.cpp")):
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            jobs.append((["g++"] + CXX_FLAGS + ["-c", str(src), "-o", str(obj)], None))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(lambda j: _run(*j), jobs):
                if verbose:
                    print(out)
    if jobs or not LIB.exists():
        _run(["g++", "-shared", "-o", str(LIB)] + [str(o) for o in objs] +
             ["-Wl,-Bsymbolic", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
