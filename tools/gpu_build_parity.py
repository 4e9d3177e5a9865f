"""Construction parity (run under gpurun): every stage of this library against the
reference rebuilt for sm_100a, (a) stage by stage on the reference's own input for
that stage, (b) end to end; then build timings of both."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes

ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so"); mine = Library()
STAGES = ["build", "merge", "flatten", "expand", "compress"]

def run_stage(sc, stage, td, sd):
    if stage == "build": sc.build_grid(td, sd)
    elif stage == "merge": sc.merge_grid(0.995)
    elif stage == "flatten": sc.flatten_grid()
    elif stage == "expand": sc.expand_grid(3)
    elif stage == "compress": assert sc.compress_grid()

def diff(a, b):
    gi_a, e_a, c_a, r_a = a; gi_b, e_b, c_b, r_b = b
    out = {}
    da, db = gi_a.as_dict(), gi_b.as_dict()
    if da != db: out["info"] = [da, db]
    for name, x, y in (("entries", e_a, e_b), ("cells", c_a, c_b), ("refs", r_a, r_b)):
        if x.shape != y.shape: out[name] = f"shape {x.shape} vs {y.shape}"
        else:
            xb, yb = x.view(np.uint8).reshape(x.shape[0], -1) if x.size else x, y.view(np.uint8).reshape(y.shape[0], -1) if y.size else y
            bad = np.nonzero((xb != yb).any(axis=1))[0] if x.size else []
            if len(bad): out[name] = f"{len(bad)} of {x.shape[0]} differ, first at {int(bad[0])}: {x[bad[0]]} vs {y[bad[0]]}"
    return out

def parity(name, tris, td, sd):
    res = {}
    sr = Scene(tris, lib=ref); sm = Scene(tris, lib=mine); st = Scene(tris, lib=mine)
    for stage in STAGES:
        before = sr.download() if stage != "build" else None
        run_stage(sr, stage, td, sd)
        want = sr.download()
        # (a) isolated: my stage on the reference's input
        if before is not None: st.upload(*before)
        run_stage(st, stage, td, sd)
        d = diff(want, st.download())
        res[f"{stage}_isolated"] = d or "identical"
        # (b) chained: my stage on my own previous output
        run_stage(sm, stage, td, sd)
        d = diff(want, sm.download())
        res[f"{stage}_chained"] = d or "identical"
    res["grid"] = sr.info().as_dict()
    print(name, json.dumps(res)[:3000], flush=True)
    sr.close(); sm.close(); st.close()
    return res

def timing(name, tris, td, sd, keep, warmup, iters, compress=False):
    out = {}
    for label, lib_ in (("ref", ref), ("mine", mine)):
        sc = Scene(tris, keep_alive=keep, lib=lib_)
        ms = sc.build_all(td, sd, 0.995, 3, compress, warmup=warmup, iters=iters)
        out[label] = {"ms_mean": round(float(ms.mean()), 3), "ms_min": round(float(ms.min()), 3), "ms_median": round(float(np.median(ms)), 3),
                      "peak_mb": round(sc.peak_bytes() / 2**20, 1)}
        sc.close()
    print("time", name, json.dumps(out), flush=True)
    return out

report = {}
which = sys.argv[1] if len(sys.argv) > 1 else "all"
report["cornell"] = parity("cornell", scenes.cornell32(), 0.12, 2.4)
report["mixed3k"] = parity("mixed3k", scenes.small_mixed(3000), 0.12, 2.4)
report["mixed50k"] = parity("mixed50k", scenes.small_mixed(50000, seed=11), 0.15, 3.0)
if which != "small":
    sp = scenes.sponza262k()
    report["sponza"] = parity("sponza", sp, 0.15, 3.0)
    hair = scenes.hairball(200000)
    report["hair200k"] = parity("hair200k", hair, 0.12, 2.4)
    report["t_cornell"] = timing("cornell", scenes.cornell32(), 0.12, 2.4, False, 2, 5)
    report["t_sponza"] = timing("sponza", sp, 0.15, 3.0, False, 2, 5)
    report["t_sponza_keep"] = timing("sponza_keep", sp, 0.15, 3.0, True, 3, 10)
    hair2 = scenes.hairball()
    report["hair2m"] = parity("hair2m", hair2, 0.12, 2.4)
    report["t_hair2m_keep"] = timing("hair2m_keep", hair2, 0.12, 2.4, True, 5, 10)
(ROOT / "gpurun_out/build_parity.json").write_text(json.dumps(report, indent=1))
print("DONE")
