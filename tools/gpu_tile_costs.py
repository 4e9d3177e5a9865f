"""Distribution of the per-tile times that the history-ordered tile kernel records (gpurun): C2 and C5 primary rays.
Prints, per buffer, the sum, the share of one resident warp, and the top of the distribution in microseconds."""
import ctypes as C, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes

lib = Library()
out = {}
for tag, tris in (("c2", scenes.sponza262k()), ("c5", scenes.sanmiguel7p8m())):
    scene = Scene(tris, keep_alive=True, lib=lib); scene.build_all(0.15, 3.0, 0.995, 3, False); scene.setup_traversal()
    rays = scenes.default_view(tris)
    n = rays.shape[0]
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda()
    d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    lib.set_option("tile_order", 8)
    for split in (0, 256):
        lib.set_option("tile_split", split)
        for _ in range(4):
            scene.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
        torch.cuda.synchronize()
        cost = np.zeros((n + 31) // 32, dtype=np.uint16)
        got = lib.dll.hgb_tile_costs(C.c_void_p(d_rays.data_ptr()), n, C.c_void_p(cost.ctypes.data), cost.size)
        assert got == cost.size, got
        mhz = 1920      # SM clock under load on this pool (bench.py's clocks line)
        us = cost.astype(np.float64) * 64 / mhz
        top = np.sort(us)[::-1]
        warps = 148 * 12 * 4
        res = {"tiles": int(cost.size), "sm_mhz": mhz, "sum_us_per_warp": round(float(us.sum() / warps), 1), "mean_us": round(float(us.mean()), 2),
               "top_us": [round(float(x), 1) for x in top[[0, 1, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095]]],
               "over_half_share": int((us >= 0.5 * us.sum() / warps).sum()), "over_quarter_share": int((us >= 0.25 * us.sum() / warps).sum())}
        out[f"{tag}_split{split}"] = res
        print(tag, split, json.dumps(res), flush=True)
    scene.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "r02_tile_costs.json", "w"), indent=1)
