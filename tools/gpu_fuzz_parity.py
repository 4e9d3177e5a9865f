"""Differential fuzzing on one GPU (run under gpurun): random scenes, random construction parameters and random
ray buffers through the reference (rebuilt for sm_100a) and through this library; every stage's grid must be
byte-identical and every hit buffer bit-identical (step counts and primitive ids, Cell and SmallCell grids).
usage: gpu_fuzz_parity.py [--large] [seconds] [first_seed]     prints one JSON summary line; failures are listed in full."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes

large = "--large" in sys.argv          # scenes of 50-200 K triangles, ray buffers big enough for the resident-warp kernels
sys.argv = [a for a in sys.argv if a != "--large"]
worker = len(sys.argv) > 1 and sys.argv[1] == "--worker"
args = sys.argv[2:] if worker else sys.argv[1:]
budget = float(args[0]) if len(args) > 0 else 120.0
seed0 = int(args[1]) if len(args) > 1 else 1

if not worker:
    # Driver: the cases run in a child process, because the reference aborts the process on inputs it cannot
    # handle (CUDA error -> abort(), src/common.h:101-108); such a seed is recorded and the child restarted behind it.
    import subprocess
    total = {"cases": 0, "failures": 0, "stage_grids_compared": 0, "hit_buffers_compared": 0, "scene_kinds": {},
             "flat_box_cases": 0, "flat_box_cases_identical": 0, "failed": [], "aborted": []}
    t_end = time.time() + budget
    seed = seed0
    while time.time() < t_end - 3:
        res = subprocess.run([sys.executable, __file__, "--worker", str(t_end - time.time()), str(seed)] + (["--large"] if large else []),
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        started = None
        for line in res.stdout.splitlines():
            if line.startswith("START "):
                started = json.loads(line[6:])
            elif line.startswith("DONE "):
                d = json.loads(line[5:])
                started = None
                total["cases"] += 1
                total["stage_grids_compared"] += d["stages"]; total["hit_buffers_compared"] += d["buffers"]
                total["scene_kinds"][d["kind"]] = total["scene_kinds"].get(d["kind"], 0) + 1
                if d["flat_box"]:
                    total["flat_box_cases"] += 1
                    total["flat_box_cases_identical"] += not d["problems"]
                elif d["problems"]:
                    total["failures"] += 1; total["failed"].append(d)
                seed = d["seed"] + 1
        if res.returncode != 0 and started is not None:
            total["aborted"].append({"seed": started["seed"], "kind": started["kind"], "tris": started["tris"],
                                     "stderr": res.stderr.strip().splitlines()[-1:]})
            seed = started["seed"] + 1
        elif res.returncode != 0:
            total["aborted"].append({"seed": seed, "stderr": res.stderr.strip().splitlines()[-3:]})
            break
    total["seeds"] = [seed0, seed - 1]
    print(json.dumps(total))
    sys.exit(0)

ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so"); mine = Library()


def soup(rng, n):
    """Triangles of wildly different sizes, some axis-aligned, some degenerate, some duplicated."""
    c = rng.uniform(-1, 1, size=(n, 3))
    size = np.select([rng.uniform(size=n) < 0.03, rng.uniform(size=n) < 0.5], [0.7, 0.08], 0.01)[:, None, None]
    V = c[:, None, :] + rng.normal(0, 1, size=(n, 3, 3)) * size
    flat = rng.uniform(size=n) < 0.15                       # axis-aligned: one coordinate shared by the three corners
    axis = rng.integers(0, 3, size=n)
    for k in range(3):
        m = flat & (axis == k)
        V[m, :, k] = V[m, :1, k]
    thin = rng.uniform(size=n) < 0.02                       # zero area: two corners coincide
    V[thin, 2] = V[thin, 1]
    dup = rng.uniform(size=n) < 0.05                        # exact duplicates of the previous triangle
    dup[0] = False
    V[dup] = V[np.nonzero(dup)[0] - 1]
    if rng.uniform() < 0.3:                                 # snap to a lattice: many coplanar faces on cell planes
        V = np.round(V * 8) / 8
    return scenes.make_tris(V[:, 0], V[:, 1], V[:, 2])


def make_scene(rng):
    kind = rng.integers(0, 5)
    n = int(rng.choice([50000, 120000, 200000])) if large else int(rng.choice([1, 2, 7, 33, 150, 800, 3000, 12000]))
    if kind == 0:
        return "soup", soup(rng, n)
    if kind == 1:
        return "atrium", scenes.atrium(max(n, 300), seed=int(rng.integers(1 << 30)))
    if kind == 2:
        return "hairball", scenes.hairball(max(100, n // 100 * 100), seed=int(rng.integers(1 << 30)))
    if kind == 3:
        return "mixed", scenes.small_mixed(n, seed=int(rng.integers(1 << 30)))
    t = soup(rng, n)                                        # far from the origin, anisotropic extent
    scale = np.array([rng.uniform(0.01, 50), rng.uniform(0.01, 50), rng.uniform(0.01, 50)], np.float32)
    shift = rng.uniform(-500, 500, size=3).astype(np.float32)
    v0, v1, v2 = scenes.tri_vertices(t)
    return "stretched", scenes.make_tris(v0 * scale + shift, v1 * scale + shift, v2 * scale + shift)


def make_rays(rng, tris):
    lo, hi = scenes.scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo)) or 1.0
    out = []
    n = int(rng.choice([140000, 300000])) if large else int(rng.choice([1, 31, 257, 4096, 20000]))
    r = scenes.random_rays(tris, n, seed=int(rng.integers(1 << 30)))
    if rng.uniform() < 0.5:                                 # origins outside the box, finite range
        r["org"] = (lo + (hi - lo) * rng.uniform(-1.0, 2.0, size=(n, 3))).astype(np.float32)
        r["tmax"] = np.float32(rng.uniform(0.2, 3.0) * diag)
        r["tmin"] = np.float32(rng.uniform(0.0, 0.1) * diag)
    if rng.uniform() < 0.3:                                 # axis-parallel directions (zero components)
        k = rng.integers(0, 3)
        r["dir"][::3, k] = 0.0
        r["dir"][1::7, (k + 1) % 3] = 0.0
    out.append(("random", r))
    w, h = [(64, 64), (128, 36), (200, 52), (96, 40)][int(rng.integers(0, 4))]
    if large:
        w, h = [(1280, 720), (1024, 512), (960, 540)][int(rng.integers(0, 3))]
    center = 0.5 * (lo + hi)
    eye = center + (hi - lo) * rng.uniform(-0.9, 0.9, size=3).astype(np.float32)
    target = center + (hi - lo) * rng.uniform(-0.2, 0.2, size=3).astype(np.float32)
    if np.linalg.norm(target - eye) < 1e-6 * diag:
        target = eye + np.array([0, 0, 1], np.float32)
    out.append(("camera", scenes.primary_rays(eye, target, (0, 1, 0), float(rng.uniform(20, 100)), w, h, 2 * diag)))
    return out


def arrays_differ(a, b):
    ga, ea, ca, ra = a; gb, eb, cb, rb = b
    if ga.as_dict() != gb.as_dict():
        return f"info {ga.as_dict()} vs {gb.as_dict()}"
    for name, x, y in (("entries", ea, eb), ("cells", ca, cb), ("refs", ra, rb)):
        if x.shape != y.shape or x.tobytes() != y.tobytes():
            return f"{name} differ ({x.shape} vs {y.shape})"
    return None


t_end = time.time() + budget
seed = seed0
while time.time() < t_end:
    rng = np.random.default_rng(seed)
    kind, tris = make_scene(rng)
    td = float(rng.choice([0.05, 0.12, 0.15, 0.3, 0.6]))
    sd = float(rng.choice([0.5, 1.0, 2.4, 3.0, 5.0]))
    alpha = float(rng.choice([0.0, 0.9, 0.995, 0.999]))
    expansion = int(rng.integers(0, 5))
    what = {"seed": seed, "kind": kind, "tris": int(tris.shape[0]), "td": td, "sd": sd, "alpha": alpha, "expansion": expansion}
    print("START " + json.dumps(what), flush=True)
    stages_checked = buffers_checked = 0
    a, b = Scene(tris, lib=ref), Scene(tris, lib=mine)
    problems = []
    for stage, run in (("build", lambda s: s.build_grid(td, sd)), ("merge", lambda s: s.merge_grid(alpha)),
                       ("flatten", lambda s: s.flatten_grid()), ("expand", lambda s: s.expand_grid(expansion))):
        run(a); run(b)
        if stage == "build":
            # a scene box without volume (all corners in one axis-aligned plane, or in one point) makes the reference
            # divide by zero; this library gives such a box a thickness (flat-box rule) -- counted apart
            gi = a.info().as_dict()
            flat_box = any(lo == hi for lo, hi in zip(gi["bbox_min"], gi["bbox_max"]))
        d = arrays_differ(a.download(), b.download())
        stages_checked += 1
        if d:
            problems.append(f"{stage}: {d}")
            break
    rays = make_rays(rng, tris)
    for compressed in (False, True):
        if problems:
            break
        if compressed:
            ok_a, ok_b = a.compress_grid(), b.compress_grid()
            if ok_a != ok_b:
                problems.append(f"compress: returned {ok_a} vs {ok_b}")
                break
            if not ok_a:
                continue
            d = arrays_differ(a.download(), b.download())
            stages_checked += 1
            if d:
                problems.append(f"compress: {d}")
                break
        a.setup_traversal(); b.setup_traversal()
        for name, r in rays:
            for mode in (HIT_STEPS, HIT_PRIM_ID):
                ha = a.trace(r, mode)
                # 3 = the library's own choice (small buffers: one thread per ray); the others force each kernel in turn
                # (the tile kernel four times in a row: from its third launch on a buffer it hands the tiles out by what
                # they cost before, the most expensive ones in parts)
                for variant in (3, 0, 1, 2, 4, 4, 4, 4):
                    mine.set_option("traverse_variant", variant)
                    hb = b.trace(r, mode)
                    buffers_checked += 1
                    if ha.tobytes() != hb.tobytes():
                        bad = np.nonzero((ha["id"] != hb["id"]) | (ha["t"].view(np.uint32) != hb["t"].view(np.uint32)))[0]
                        problems.append(f"hits {name} mode {mode} compressed {compressed} variant {variant}: {len(bad)} of {len(r)} differ, "
                                        f"first {int(bad[0]) if len(bad) else -1}")
                mine.set_option("traverse_variant", 3)
    a.close(); b.close()
    print("DONE " + json.dumps({**what, "stages": stages_checked, "buffers": buffers_checked, "problems": problems,
                                "flat_box": bool(flat_box)}), flush=True)
    seed += 1
