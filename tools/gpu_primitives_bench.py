"""The hand-written primitives against the toolkit's CUB 2.8 (DeviceScan, DeviceRadixSort, DevicePartition,
DeviceReduce) on the same box, same buffers, results compared (SURVEY.md section 2.3 names CUB's sm_100 tunings as the bar).
  python tools/gpu_primitives_bench.py build      here: nvcc -> hagrid_b200/_build/libcub_bench.so
  python tools/gpu_primitives_bench.py            under gpurun: JSON to stdout and gpurun_out/r02_primitives.json
Times: CUDA events around 20 back-to-back calls after 3 warm-ups, input larger than the L2 for the big sizes."""
import ctypes as C, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
LIB = ROOT / "hagrid_b200" / "_build" / "libcub_bench.so"
if len(sys.argv) > 1 and sys.argv[1] == "build":
    LIB.parent.mkdir(exist_ok=True)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
                    str(ROOT / "tools/cub_bench/cub_bench.cu"), "-o", str(LIB)], check=True)
    print(LIB); sys.exit(0)

import numpy as np, torch
from hagrid_b200 import Library, Scene, scenes
lib = Library()
cub = C.CDLL(str(LIB))
sc = Scene(scenes.cornell32(), keep_alive=True, lib=lib)
P = lambda t: C.c_void_p(t.data_ptr())

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps

out = {}
for n in (1 << 16, 1 << 20, 13_000_000, 1 << 26):
    x = torch.randint(0, 8, (n,), dtype=torch.int32, device="cuda")
    mine = torch.empty(n + 1, dtype=torch.int32, device="cuda"); theirs = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    t_mine = timed(lambda: lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, x.data_ptr(), n, 4, mine.data_ptr()), "scan"))
    t_cub = timed(lambda: cub.cub_exclusive_sum_i32(P(x), P(theirs), n))
    same = bool(torch.equal(mine[:n], theirs[:n]))
    out[f"scan_i32_{n}"] = {"mine_ms": round(t_mine, 4), "cub_ms": round(t_cub, 4), "mine_GBps": round(8 * n / t_mine / 1e6, 1),
                            "cub_GBps": round(8 * n / t_cub / 1e6, 1), "ratio": round(t_cub / t_mine, 3), "same": same}
    x8 = torch.randint(0, 4, (n,), dtype=torch.int64, device="cuda") | (torch.randint(0, 2, (n,), dtype=torch.int64, device="cuda") << 32)
    mine8 = torch.empty(n + 1, dtype=torch.int64, device="cuda"); theirs8 = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    t_mine = timed(lambda: lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, x8.data_ptr(), n, 8, mine8.data_ptr()), "scan"))
    t_cub = timed(lambda: cub.cub_exclusive_sum_u64(P(x8), P(theirs8), n))
    out[f"scan_u64_{n}"] = {"mine_ms": round(t_mine, 4), "cub_ms": round(t_cub, 4), "mine_GBps": round(16 * n / t_mine / 1e6, 1),
                            "cub_GBps": round(16 * n / t_cub / 1e6, 1), "ratio": round(t_cub / t_mine, 3), "same": bool(torch.equal(mine8[:n], theirs8[:n]))}
    for bits in (8, 16, 23):
        keys = torch.randint(0, 1 << bits, (n,), dtype=torch.int32, device="cuda"); vals = torch.arange(n, dtype=torch.int32, device="cuda")
        k1, v1 = keys.clone(), vals.clone()
        k2, v2, k2a, v2a = keys.clone(), vals.clone(), torch.empty_like(keys), torch.empty_like(vals)
        def mine_sort():
            k1.copy_(keys); v1.copy_(vals)
            lib.check(lib.dll.hgb_prim_sort_pairs(sc._h, k1.data_ptr(), v1.data_ptr(), n, bits), "sort")
        def cub_sort():
            k2.copy_(keys); v2.copy_(vals)
            return cub.cub_sort_pairs(P(k2), P(v2), P(k2a), P(v2a), n, bits)
        def copies():
            k1.copy_(keys); v1.copy_(vals)
        t_copy = timed(copies)
        t_mine, t_cub = timed(mine_sort) - t_copy, timed(cub_sort) - t_copy
        where = cub_sort(); torch.cuda.synchronize()
        res = v2a if where == 1 else v2
        out[f"sort_pairs_{n}_bits{bits}"] = {"mine_ms": round(t_mine, 4), "cub_ms": round(t_cub, 4), "mine_Mpairs_s": round(n / t_mine / 1e3, 1),
                                             "cub_Mpairs_s": round(n / t_cub / 1e3, 1), "ratio": round(t_cub / t_mine, 3), "same": bool(torch.equal(v1, res))}
    flags = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda")
    o1, o2, cnt = torch.empty_like(x), torch.empty_like(x), torch.zeros(1, dtype=torch.int32, device="cuda")
    t_mine = timed(lambda: lib.check(lib.dll.hgb_prim_partition(sc._h, x.data_ptr(), flags.data_ptr(), n, o1.data_ptr()), "partition"))
    t_cub = timed(lambda: cub.cub_partition_flagged(P(x), P(flags), P(o2), P(cnt), n))
    out[f"partition_{n}"] = {"mine_ms": round(t_mine, 4), "cub_ms": round(t_cub, 4), "ratio": round(t_cub / t_mine, 3), "same": bool(torch.equal(o1, o2)),
                             "note": "mine returns the count to the host (one sync), like the reference's wrapper; the CUB call here does not"}
    r1, r2 = torch.zeros(4, dtype=torch.int32, device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda")
    t_mine = timed(lambda: lib.check(lib.dll.hgb_prim_reduce(sc._h, x.data_ptr(), n, 0, r1.data_ptr()), "reduce"))
    t_cub = timed(lambda: cub.cub_reduce_sum_i32(P(x), P(r2), n))
    out[f"reduce_sum_{n}"] = {"mine_ms": round(t_mine, 4), "cub_ms": round(t_cub, 4), "ratio": round(t_cub / t_mine, 3), "same": bool(r1[0] == r2[0])}
    del x, mine, theirs, x8, mine8, theirs8
for k, v in out.items():
    print(k, json.dumps(v), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r02_primitives.json").write_text(json.dumps(out, indent=1))
