"""Round-1 GPU probe (run under gpurun): reference pipeline stats, golden
fixtures, traversal parity of every kernel variant on reference-built grids,
and first timings reference vs. this library.  Writes gpurun_out/probe.json
and gpurun_out/golden_*.npz."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
mine = Library()
report = {}


def stage_dump(sc):
    gi, e, c, r = sc.download()
    return {"info": gi.as_dict(), "entries": e, "cells": c, "refs": r}


def ref_pipeline(tris, td, sd, alpha=0.995, exp=3, compress=False, keep_stages=False):
    sc = Scene(tris, lib=ref)
    stages = {}
    sc.build_grid(td, sd);      stages["build"] = stage_dump(sc) if keep_stages else None
    sc.merge_grid(alpha);       stages["merge"] = stage_dump(sc) if keep_stages else None
    sc.flatten_grid();          stages["flatten"] = stage_dump(sc) if keep_stages else None
    sc.expand_grid(exp);        stages["expand"] = stage_dump(sc) if keep_stages else None
    if compress:
        assert sc.compress_grid()
        stages["compress"] = stage_dump(sc) if keep_stages else None
    return sc, stages


def transplant(src_scene, tris):
    gi, e, c, r = src_scene.download()
    dst = Scene(tris, lib=mine)
    dst.upload(gi, e, c, r)
    dst.setup_traversal()
    return dst


def compare(name, tris, rays, td, sd, compress, time_it=False, golden=None):
    res = {}
    sc_ref, stages = ref_pipeline(tris, td, sd, compress=compress, keep_stages=golden is not None)
    sc_ref.setup_traversal()
    info = sc_ref.info().as_dict()
    res["grid"] = info
    sc_mine = transplant(sc_ref, tris)
    hits_ref = {m: sc_ref.trace(rays, m) for m in (HIT_STEPS, HIT_PRIM_ID)}
    for variant in (0, 1):
        mine.set_option("traverse_variant", variant)
        for m in (HIT_STEPS, HIT_PRIM_ID):
            h = sc_mine.trace(rays, m)
            same_id = int((h["id"] == hits_ref[m]["id"]).sum())
            same_t = int((h["t"].view(np.uint32) == hits_ref[m]["t"].view(np.uint32)).sum())
            res[f"v{variant}_mode{m}"] = {"n": int(rays.shape[0]), "same_id": same_id, "same_t_bits": same_t}
    steps = hits_ref[HIT_STEPS]["id"].astype(np.int64)
    res["steps_mean"] = float(steps.mean()); res["steps_max"] = int(steps.max())
    res["hit_fraction"] = float((hits_ref[HIT_PRIM_ID]["id"] >= 0).mean())
    if time_it:
        n = rays.shape[0]
        for label, sc, lib_, variants in (("ref", sc_ref, ref, (0,)), ("mine", sc_mine, mine, (0, 1))):
            d_rays = sc.device_alloc(rays.nbytes); d_hits = sc.device_alloc(n * 16)
            sc.to_device(d_rays, rays)
            for v in variants:
                lib_.set_option("traverse_variant", v)
                ms = sc.traverse_timed(d_rays, d_hits, n, HIT_PRIM_ID, warmup=5, iters=20)
                res[f"time_{label}_v{v}"] = {"ms_median": float(np.median(ms)), "ms_min": float(ms.min()),
                                             "mrays_s": float(n * len(ms) / (1000.0 * ms.sum()))}
            sc.device_free(d_rays); sc.device_free(d_hits)
    if golden is not None:
        arrs = {"tris": tris, "rays": rays, "hits_steps": hits_ref[HIT_STEPS], "hits_ids": hits_ref[HIT_PRIM_ID],
                "params": np.array([td, sd, 0.995, 3, int(compress)], dtype=np.float64)}
        for st, d in stages.items():
            if d is None: continue
            arrs[f"{st}_info"] = np.frombuffer(json.dumps(d["info"]).encode(), dtype=np.uint8)
            arrs[f"{st}_entries"] = d["entries"]; arrs[f"{st}_cells"] = d["cells"]; arrs[f"{st}_refs"] = d["refs"]
        np.savez_compressed(OUT / f"golden_{golden}.npz", **arrs)
    report[name] = res
    print(name, json.dumps(res)[:1500], flush=True)
    sc_ref.close(); sc_mine.close()


def time_ref_build(name, tris, td, sd, compress, keep, warmup, iters):
    sc = Scene(tris, keep_alive=keep, lib=ref)
    ms = sc.build_all(td, sd, 0.995, 3, compress, warmup=warmup, iters=iters)
    report[name] = {"ms_mean": float(ms.mean()), "ms_min": float(ms.min()), "grid": sc.info().as_dict(),
                    "peak_mb": sc.peak_bytes() / 2**20}
    print(name, json.dumps(report[name]), flush=True)
    sc.close()


t0 = time.time()
corn = scenes.cornell32()
compare("c1_cornell", corn, scenes.cornell_view(256, 256), 0.12, 2.4, False, golden="cornell")
compare("c1_cornell_small", corn, scenes.cornell_view(64, 64), 0.12, 2.4, True, golden="cornell_small")
mixed = scenes.small_mixed(3000)
compare("mixed3k", mixed, scenes.random_rays(mixed, 20000, tmax=10.0), 0.12, 2.4, False, golden="mixed3k")
compare("mixed3k_small", mixed, scenes.random_rays(mixed, 20000, tmax=10.0), 0.12, 2.4, True)
sp = scenes.sponza262k()
compare("c2_sponza_primary", sp, scenes.default_view(sp), 0.15, 3.0, False, time_it=True)
compare("c2_sponza_primary_long", sp, scenes.default_view(sp, along_long_axis=True), 0.15, 3.0, False, time_it=True)
compare("c3_sponza_random_compressed", sp, scenes.random_rays(sp, 4194304), 0.15, 3.0, True, time_it=True)
time_ref_build("build_ref_c2", sp, 0.15, 3.0, False, False, 2, 5)
time_ref_build("build_ref_c2_keep", sp, 0.15, 3.0, False, True, 2, 5)
hair = scenes.hairball()
time_ref_build("build_ref_c4_keep", hair, 0.12, 2.4, False, True, 3, 10)
report["wall_s"] = time.time() - t0
(OUT / "probe.json").write_text(json.dumps(report, indent=1))
print("DONE", time.time() - t0)
