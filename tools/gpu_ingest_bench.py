"""Scene ingest timing (run under gpurun): OBJ file -> device Tri array, the reference's single-threaded
load_model + upload vs this library's parallel parse + device triangle setup."""
import json, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tris = scenes.hairball(n)
out = {"triangles": n}
with tempfile.TemporaryDirectory() as d:
    path = Path(d) / "scene.obj"
    t0 = time.perf_counter(); scenes.write_obj(path, tris); out["write_obj_s"] = round(time.perf_counter() - t0, 2)
    out["file_mb"] = round(path.stat().st_size / 1e6, 1)
    res = {}
    for label, lib, threads in (("reference", Library(ROOT / "oracle/_ref/libhagrid_ref.so"), 0), ("hagrid_b200_1thread", Library(), 1),
                                ("hagrid_b200_all_threads", Library(), 0)):
        ts = []
        for _ in range(2):
            t0 = time.perf_counter(); sc = Scene(path, lib=lib, threads=threads); ts.append(time.perf_counter() - t0)
            got = sc.download_tris(); sc.close()
        res[label] = got
        out[label + "_s"] = round(min(ts), 3)
    out["identical"] = bool(res["reference"].tobytes() == res["hagrid_b200_all_threads"].tobytes() == res["hagrid_b200_1thread"].tobytes())
print(json.dumps(out))
