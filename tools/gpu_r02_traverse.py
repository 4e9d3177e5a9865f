"""Round-2 traversal measurements on one B200 (run under gpurun): the tile kernel and the host-buffer frame against the
reference rebuilt for sm_100a, same buffers, hits compared bit for bit (the switches of the rejected experiments of
profiles/r02_traverse_experiments.md went with their code).
usage: gpu_r02_traverse.py [c2] [c5] [e2e] [shards]   (default: all)
Timing as in bench.py: CUDA event pair per launch on the legacy stream, 256 MiB memset between launches (L2 flushed)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes, sharding  # noqa: E402

what = set(sys.argv[1:]) or {"c2", "c5", "e2e", "shards"}
ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
mine = Library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}


def timed(fn, iters=30, warmup=5):
    for _ in range(warmup):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_(); a[i].record(); fn(); b[i].record()
    torch.cuda.synchronize()
    ms = np.array([x.elapsed_time(y) for x, y in zip(a, b)])
    return {"ms_mean": round(float(ms.mean()), 4), "ms_median": round(float(np.median(ms)), 4), "ms_min": round(float(ms.min()), 4)}


def dev(rays):
    n = rays.shape[0]
    return torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(), torch.empty((n, 4), dtype=torch.float32, device="cuda")


def hits_of(d_hits):
    return d_hits.cpu().numpy().view(np.dtype([("id", "<i4"), ("t", "<u4"), ("u", "<f4"), ("v", "<f4")])).reshape(-1)


def same(a, b):
    return bool(np.array_equal(a["id"], b["id"]) and np.array_equal(a["t"], b["t"]))


def scene_pair(tris, compress=False):
    sr = Scene(tris, keep_alive=True, lib=ref); sr.build_all(0.15, 3.0, 0.995, 3, compress); sr.setup_traversal()
    sm = Scene(tris, keep_alive=True, lib=mine); sm.build_all(0.15, 3.0, 0.995, 3, compress); sm.setup_traversal()
    return sr, sm


def compare_buffer(tag, sr, sm, rays, settings, iters=30):
    n = rays.shape[0]
    d_rays, d_hits = dev(rays)
    res = {"n": n}
    r = timed(lambda: sr.traverse(d_rays, d_hits, n, HIT_PRIM_ID), iters)
    want = hits_of(d_hits)
    res["reference"] = r
    for name, opts in settings.items():
        for k, v in opts.items():
            mine.set_option(k, v)
        d_hits.zero_()
        t = timed(lambda: sm.traverse(d_rays, d_hits, n, HIT_PRIM_ID), iters)
        t["identical"] = same(hits_of(d_hits), want)
        t["speedup"] = round(r["ms_mean"] / t["ms_mean"], 3)
        res[name] = t
    out[tag] = res
    print(tag, json.dumps(res), flush=True)
    return want


SETTINGS = {"default": {}}

if "c2" in what or "e2e" in what or "shards" in what:
    tris = scenes.sponza262k()
    sr, sm = scene_pair(tris)
    primary = scenes.default_view(tris)
    if "c2" in what:
        compare_buffer("c2_primary", sr, sm, primary, SETTINGS, 50)
        compare_buffer("c2_long", sr, sm, scenes.default_view(tris, along_long_axis=True), SETTINGS, 30)
    if "shards" in what:
        # one frame split over 2, 4, 8 ranks: rank 0's share, bands dealt round-robin
        W = 1920
        for world in (2, 4, 8):
            idx = sharding.interleaved_bands(primary.shape[0], 0, world, sharding.raster_granule(W))
            part = np.ascontiguousarray(primary[idx])
            compare_buffer(f"c2_shard_1of{world}", sr, sm, part,
                           {"per_thread": {"tile_min_rays": 1 << 30}, "tiles": {"tile_min_rays": 0}}, 30)
        mine.set_option("tile_min_rays", -1)
    if "e2e" in what:
        n = primary.shape[0]
        want = sm.trace(primary, HIT_PRIM_ID)
        h_rays = torch.from_numpy(primary.view(np.float32).reshape(n, 8)).pin_memory()
        h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        res = {}
        r = timed(lambda: ref.check(ref.dll.hgb_traverse_grid_host(sr._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f"), 20, 3)
        res["reference"] = r
        for mode in (0,):
            for stage in (0,):
                for chunk in (224, 256, 288, 384):
                    mine.set_option("host_frame_chunk_rays", chunk * 1024)
                    h_hits.zero_()
                    t = timed(lambda: mine.check(mine.dll.hgb_traverse_grid_host(sm._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f"), 20, 3)
                    t["identical"] = bool(np.array_equal(h_hits.numpy().view(np.int32)[:, 0], want["id"]) and
                                          np.array_equal(h_hits.numpy()[:, 1], want["t"]))
                    t["speedup"] = round(r["ms_mean"] / t["ms_mean"], 3)
                    res[f"mode{mode}_stage{stage}_chunk{chunk}K"] = t
        mine.set_option("host_frame_chunk_rays", 0)
        out["c2_e2e"] = res
        print("c2_e2e", json.dumps(res), flush=True)
    sr.close(); sm.close()

if "c5" in what:
    tris = scenes.sanmiguel7p8m()
    sr, sm = scene_pair(tris)
    primary = scenes.default_view(tris)
    first = compare_buffer("c5_primary", sr, sm, primary, SETTINGS, 30)
    bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"].view(np.float32))
    compare_buffer("c5_bounce", sr, sm, bounce, {"default": {}}, 10)
    sr.close(); sm.close()

Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r02_traverse.json").write_text(json.dumps(out, indent=1))
