"""The reference's own executable against the drop-in on one GPU (run under gpurun): src/main.cpp + src/load_obj.cpp
compiled unmodified, linked once with the reference's kernels (oracle/_ref/hagrid_ref) and once with this
repository's (oracle/_ref/hagrid_dropin). Same .obj, same .rays files, the front end's own timing loops and its own
report (src/main.cpp:414-446, 494-515) -- nothing of this repository's host code is involved. `traverse_grid` here is
the reference-verbatim mode (Hit.id = step count, src/traverse.cu:93)."""
import json, re, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import scenes

tmp = Path(tempfile.mkdtemp())
tris = scenes.sponza262k()
lo, hi = scenes.scene_bbox(tris)
diag = float(np.linalg.norm(hi - lo))
scenes.write_obj(tmp / "scene.obj", tris)
scenes.write_rays(tmp / "primary.rays", scenes.default_view(tris))
scenes.write_rays(tmp / "random.rays", scenes.random_rays(tris, 1 << 22))
out = {"scene": "sponza262k (262 267 triangles) as .obj, rays as .rays files"}


def run(exe, rays, extra):
    cmd = [str(ROOT / "oracle" / "_ref" / exe), "-td", "0.15", "-sd", "3.0", "-k", "-nb", "10", "-wb", "5", "-r", str(tmp / rays),
           "-n", "50", "-w", "5", "-tmax", str(diag)] + extra + [str(tmp / "scene.obj")]
    t0 = time.perf_counter()
    text = subprocess.run(cmd, capture_output=True, text=True, timeout=600, check=True).stdout
    wall = time.perf_counter() - t0
    grid = re.search(r"Grid built in ([\d.e+-]+) ms \((\S+), (\d+) cells, (\d+) references\)", text)
    return {"build_ms": float(grid.group(1)), "grid": f"{grid.group(2)}, {grid.group(3)} cells, {grid.group(4)} references",
            "mrays_s": float(re.search(r"([\d.e+-]+) Mrays/sec", text).group(1)),
            "median_ms": float(re.search(r"# Median: ([\d.e+-]+) ms", text).group(1)),
            "intersections": int(re.search(r"(\d+) intersection", text).group(1)), "process_wall_s": round(wall, 2)}


for name, rays, extra in (("primary", "primary.rays", []), ("random_compressed", "random.rays", ["--compress"])):
    a, b = run("hagrid_ref", rays, extra), run("hagrid_dropin", rays, extra)
    out[name] = {"hagrid_ref": a, "hagrid_dropin": b, "same_grid_line": a["grid"] == b["grid"],
                 "same_intersections": a["intersections"] == b["intersections"]}
print(json.dumps(out))
