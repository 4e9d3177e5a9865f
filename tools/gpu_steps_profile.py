"""Distribution of per-ray step counts (Hit.id in the reference-verbatim mode) and of per-tile critical paths for a
primary view: how much of a launch is the tail of its longest rays. usage: gpu_steps_profile.py [sanmiguel|sponza|sponza_long]"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_STEPS, HIT_PRIM_ID, Scene, scenes

which = sys.argv[1] if len(sys.argv) > 1 else "sanmiguel"
tris = scenes.sanmiguel7p8m() if which == "sanmiguel" else scenes.sponza262k()
rays = scenes.default_view(tris, along_long_axis=which.endswith("long"))
sc = Scene(tris, keep_alive=True)
sc.build_all(0.15, 3.0)
sc.setup_traversal()
steps = sc.trace(rays, HIT_STEPS)["id"].astype(np.int64)
ids = sc.trace(rays, HIT_PRIM_ID)["id"]
W, H = 1920, 1080
img = steps.reshape(H, W)
tiles = img.reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(H // 4, W // 8, 32)
tmax, tmean = tiles.max(axis=2), tiles.mean(axis=2)
order = tmax.reshape(-1)                      # tile index order = the order the tile kernel hands them out
total = int(order.sum())
warps = 148 * 40
out = {"scene": which, "hit_fraction": round(float((ids >= 0).mean()), 4),
       "steps_per_ray": {"mean": round(float(steps.mean()), 2), "p50": int(np.percentile(steps, 50)), "p99": int(np.percentile(steps, 99)),
                         "p99.9": int(np.percentile(steps, 99.9)), "max": int(steps.max())},
       "tile_critical_path(max steps of its 32 rays)": {"mean": round(float(order.mean()), 2), "p99": int(np.percentile(order, 99)), "max": int(order.max())},
       "balanced_share_per_warp(sum of tile maxima / 5920 warps)": round(total / warps, 1),
       "longest_tile_over_balanced_share": round(float(order.max()) / (total / warps), 3),
       "rows_of_tiles_with_the_20_longest_tiles": sorted(set(int(i) // (W // 8) for i in np.argsort(order)[-20:])),
       "tile_rows": H // 4}
# list scheduling simulation: warps take tiles in index order (cost = tile maximum); makespan over the balanced share
import heapq
def makespan(costs):
    heap = [0] * warps
    for c in costs:
        t = heapq.heappop(heap); heapq.heappush(heap, t + int(c))
    return max(heap)
out["makespan_over_balanced_share"] = {"index_order": round(makespan(order) / (total / warps), 3),
                                       "reverse_order": round(makespan(order[::-1]) / (total / warps), 3),
                                       "longest_first": round(makespan(np.sort(order)[::-1]) / (total / warps), 3)}
print(json.dumps(out))
