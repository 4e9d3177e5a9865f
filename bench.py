#!/usr/bin/env python
"""Headline benchmark of the irregular-grid path (BASELINE.json, config C2):
procedural Sponza-class scene (262 267 triangles), 1920x1080 primary rays,
--top-density 0.15 --snd-density 3.0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one traverse_grid pass over one frame (2 073 600 rays). Every step is
timed with a CUDA event pair on the launching (legacy default) stream, exactly like
the reference's profile() (src/profile.cu:5-18, src/main.cpp:414-441); the L2 is
flushed between steps. Mrays/s = rays * K / (1000 * sum ms), max over ranks.

  value  device-resident: rays and hits stay in HBM (the reference's own metric)
  e2e    the same frame through the C ABI with HOST buffers: pinned H2D of the rays,
         traversal, D2H of the hits, all inside the timed region
N > 1 (torchrun): one process per GPU, every rank builds its replica of the grid and
traces the same frame (weak scaling with equal work per GPU, no data-path collective;
one all-reduce of the timing counters per measurement).

--impl reference runs cg-saarland/hagrid itself: the reference has no CPU path, so
its CUDA sources rebuilt for sm_100a (oracle/_ref, see oracle/build_ref.sh) are driven
through the same C ABI on the same GPU. If that build is absent the CPU oracle port
is timed on a bounded sample instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION = 0.15, 3.0, 0.995, 3
WIDTH, HEIGHT = 1920, 1080
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
               "clocks_event_reasons.sw_power_cap")


def measured_peak_gbs():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""

    def __init__(self, gpu_index: int):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={CLOCK_QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [r.split(",") for r in Path(self.file.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.file.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.strip().lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def camera_path_view(tris, scenes, view: int):
    """View `view` of the N-view camera path: the reference's start view (scene centre, +z), yawed."""
    lo, hi = scenes.scene_bbox(tris)
    eye = 0.5 * (lo + hi)
    ext = hi - lo
    ang = np.deg2rad(7.0 * view)
    target = eye + np.array([np.sin(ang), 0.0, np.cos(ang)], np.float32)
    return scenes.primary_rays(eye, target, (0, 1, 0), 60.0, WIDTH, HEIGHT, float(np.sqrt(np.dot(ext, ext))))


def scene_bytes(info: dict, num_tris: int) -> int:
    """S = 4E + (32|16)C + 4R + 48N, read once per launch."""
    return 4 * info["num_entries"] + (16 if info["compressed"] else 32) * info["num_cells"] + 4 * info["num_refs"] + 48 * num_tris


def cpu_oracle_baseline(tris, info, arrays, rays, seconds=12.0):
    """The CPU restatement traced on a bounded sample of the same rays, on all host cores."""
    from oracle import oracle
    grid = oracle.Grid.from_arrays(info, *arrays)
    cores = os.cpu_count() or 1
    sample = rays[:: max(1, rays.shape[0] // 16384)][:16384]
    t0 = time.perf_counter()
    grid.traverse(tris, sample, 1, cores)
    rate = sample.shape[0] / (time.perf_counter() - t0)
    n = int(min(rays.shape[0], max(16384, rate * seconds)))
    sample = rays[:: max(1, rays.shape[0] // n)][:n]
    passes = int(max(1, min(400, round(seconds * rate / sample.shape[0]))))
    t0 = time.perf_counter()
    for _ in range(passes):
        grid.traverse(tris, sample, 1, cores)
    dt = time.perf_counter() - t0
    return {"value": round(passes * sample.shape[0] / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{passes} passes over {sample.shape[0]} of {rays.shape[0]} primary rays, {dt:.1f} s of CPU work, "
                      f"oracle/hagrid_oracle.c (CPU restatement of src/traverse.cu), {cores} threads"}


def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this rank's threads (and with them the first touch of its page-locked frame buffers)
    on the CPUs NVML reports as local to its GPU. Returns a description for the JSON line."""
    if os.environ.get("HGB_BENCH_NO_AFFINITY"):
        return "not bound (HGB_BENCH_NO_AFFINITY)"
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "not bound (no local CPUs allowed)"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} CPUs local to GPU {local_rank}"
    except Exception as e:                                      # the bench must not depend on NVML
        return f"not bound ({type(e).__name__})"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="hagrid_b200", choices=["hagrid_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    reference = args.impl == "reference"

    affinity = bind_near_gpu(local_rank) if world > 1 else "not bound (single rank)"
    import torch
    import torch.distributed as dist
    from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes

    ref_lib_path = ROOT / "oracle" / "_ref" / "libhagrid_ref.so"
    if reference and not ref_lib_path.exists():
        return reference_on_cpu(args, rank)
    launched_ranks = world
    if reference:
        # cg-saarland/hagrid is a single-process, single-GPU program (SURVEY.md quick facts): its arm runs
        # on rank 0's GPU alone, whatever N was launched; the other ranks leave without work
        if rank != 0:
            return
        world = 1

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hagrid_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = Library(ref_lib_path) if reference else Library()
    tris = scenes.sponza262k()
    rays = camera_path_view(tris, scenes, 0)      # weak scaling: the same frame on every rank, so per-GPU work does not depend on N
    n = rays.shape[0]

    # ---- construction (every rank builds its own replica; reported, not the headline)
    scene = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
    build_ms = scene.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=False, warmup=3, iters=10)
    scene.setup_traversal()
    info = scene.info().as_dict()

    # ---- device-resident frames: torch owns the ray / hit buffers
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda()
    d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def step():
        scene.traverse(d_rays, d_hits, n, HIT_PRIM_ID)

    for _ in range(args.warmup):
        flush.zero_(); step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = lib.kernel_launches()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        step()
        stops[i].record()
    torch.cuda.synchronize()
    launches = lib.kernel_launches() - launches0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(starts, stops)], dtype=np.float64)
    hits = d_hits.cpu().numpy().view(np.dtype([("id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])).reshape(-1)

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
    h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 100))
    for _ in range(2):
        lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "e2e")
    torch.cuda.synchronize()
    e2e_ms = []
    for _ in range(e2e_steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "e2e")
        b.record()
        b.synchronize()
        e2e_ms.append(a.elapsed_time(b))
    e2e_ok = bool(np.array_equal(h_hits.numpy().view(np.int32)[:, 0], hits["id"]))
    clocks = sampler.stop() if sampler else None          # sampled across both timed loops

    # ---- secondary number: one viewer frame (camera -> BGRA image in host memory), src/main.cpp:598-621
    frame = {}
    if rank == 0:
        from hagrid_b200 import make_camera
        lo_, hi_ = scenes.scene_bbox(tris)
        eye_ = 0.5 * (lo_ + hi_)
        cam = make_camera(eye_, eye_ + np.array([0, 0, 1], np.float32), (0, 1, 0), 60.0, WIDTH / HEIGHT, lib=lib)
        clip = float(np.linalg.norm(hi_ - lo_))
        image = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8).pin_memory()
        render = lambda: lib.check(lib.dll.hgb_render_frame(scene._h, cam.ctypes.data, clip, WIDTH, HEIGHT, 0, image.data_ptr()), "frame")
        for _ in range(2):
            render()
        t0 = time.perf_counter()
        reps = 5 if reference else 20
        for _ in range(reps):
            render()
        frame = {"viewer_frame_ms": round((time.perf_counter() - t0) * 1e3 / reps, 3),
                 "viewer_frame": "hgb_render_frame, 1920x1080 depth image to pinned host memory, wall clock; " +
                                 ("reference: CPU gen_rays + upload + traverse_grid + download + CPU update_surface" if reference
                                  else "one fused launch (generate, trace, colour) + 4 B/pixel download")}

    # ---- secondary numbers: incoherent rays on the compressed grid (C3)
    inc = {}
    if rank == 0 and not os.environ.get("HGB_BENCH_SKIP_C3"):
        sc3 = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
        sc3.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=True)
        sc3.setup_traversal()
        r3 = scenes.random_rays(tris, 1 << 22)
        d3 = torch.from_numpy(r3.view(np.float32).reshape(-1, 8)).cuda()
        h3 = torch.empty((r3.shape[0], 4), dtype=torch.float32, device="cuda")
        ms3 = sc3.traverse_timed(d3, h3, r3.shape[0], HIT_PRIM_ID, warmup=3, iters=10)
        inc = {"incoherent_mrays_s": round(r3.shape[0] * len(ms3) / (1000.0 * float(ms3.sum())), 1),
               "incoherent_workload": "C3: 4194304 random rays, --compress"}
        sc3.close()
        scene.setup_traversal()

    # ---- aggregate over ranks: sum of rays, max of time
    total_ms, e2e_total = float(step_ms.sum()), float(np.sum(e2e_ms))
    if world > 1:
        t = torch.tensor([total_ms, e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total = float(t[0]), float(t[1])
    if rank == 0:
        value = world * n * args.steps / (1000.0 * total_ms)
        e2e_value = world * n * e2e_steps / (1000.0 * e2e_total)
        peak, peak_src = measured_peak_gbs()
        algo_bytes = 48 * n + scene_bytes(info, tris.shape[0])
        launch_ms = float(np.mean(step_ms))
        achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.loads((ROOT / "profiles" / "r01_traverse_dram.json").read_text())[args.impl]["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "Mrays/s, primary rays (closest hit, bit-exact prim ids)", "value": round(value, 1), "unit": "Mrays/s",
            "n_gpus": launched_ranks, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": args.impl,
            "config": {"workload": "C2: sponza262k stand-in (262267 tris), 1920x1080 primary rays, -td 0.15 -sd 3.0 -a 0.995 -e 3",
                       "rays_per_step_per_gpu": n, "frames": "every rank builds its own replica of the grid and traces the same 1920x1080 frame (fixed work per GPU)",
                       "grid": {k: info[k] for k in ("dims", "shift", "num_cells", "num_entries", "num_refs")},
                       "l2": "256 MiB memset between steps (L2 flushed); scene+rays+hits = %.1f MB" % (algo_bytes / 1e6),
                       "timing": "CUDA event pair per step on the legacy default stream, sum over steps, max over ranks"},
            "e2e": {"value": round(e2e_value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 16 * n,
                    "ms_per_step": round(e2e_total / e2e_steps, 4), "steps": e2e_steps, "hits_match_device_path": e2e_ok,
                    "api": "hgb_traverse_grid_host (pinned host buffers)", "host_affinity_rank0": affinity},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src, "kernel": ("traverse_pid<Cell, Tri>" if reference else "traverse_tiles<Cell, 1>") + " (the only kernel of a step)",
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "gather kernel bound by instruction issue and dependent-load latency (ncu: 70 % of peak issue rate, 22 of 32 lanes active): the scene "
                                 "lives in L2, compulsory HBM traffic is 48 B/ray; see DESIGN.md section 5"},
            "build_ms": {"mean": round(float(build_ms.mean()), 3), "median": round(float(np.median(build_ms)), 3),
                         "min": round(float(build_ms.min()), 3), "iters": int(build_ms.shape[0]),
                         "what": "build+merge+flatten+expand, keep-alive, event-timed like src/main.cpp:494-508"},
            "build_roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                               "algorithmic_bytes": int(48 * tris.shape[0] + 4 * info["num_entries"] + 32 * info["num_cells"] + 4 * info["num_refs"]),
                               "achieved": round((48 * tris.shape[0] + 4 * info["num_entries"] + 32 * info["num_cells"] + 4 * info["num_refs"])
                                                 / (float(build_ms.mean()) * 1e6), 2),
                               "note": "SURVEY 8(d): read the triangles once + write the final grid once, over the whole multi-pass pipeline "
                                       "(about 60 launches at this size, launch- and latency-bound); per-kernel DRAM throughput of the 2 M-triangle "
                                       "build is in profiles/r01_build_hair2m_kernels.csv"},
            "hit_fraction": round(float((hits["id"] >= 0).mean()), 4),
        }
        line.update(inc)
        line.update(frame)
        if reference:
            line["ranks_used"] = 1
            line["cpu_baseline"] = {"value": line["value"], "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                    "sample": "whole workload; cg-saarland/hagrid has no CPU build/traverse path, so the arm runs its "
                                              "CUDA sources rebuilt for sm_100a (oracle/_ref) on the same GPU, one host thread"}
        elif world == 1 and not args.no_cpu_baseline:
            gi, e, c, r = scene.download()
            line["cpu_baseline"] = cpu_oracle_baseline(tris, gi.as_dict(), (e, c, r), rays)
        print(json.dumps(line), flush=True)
    scene.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_on_cpu(args, rank):
    """Fallback of the reference arm when oracle/_ref was not built: the CPU oracle port."""
    if rank != 0:
        return
    from hagrid_b200 import scenes
    from oracle import oracle
    tris = scenes.sponza262k()
    rays = camera_path_view(tris, scenes, 0)
    grid = oracle.Grid.build(tris, TOP_DENSITY, SND_DENSITY)
    grid.merge(ALPHA); grid.flatten(); grid.expand(EXPANSION)
    cores = os.cpu_count() or 1
    sample = rays[::64]
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        grid.traverse(tris, sample, 1, cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
        if sum(times) > 150:
            break
    value = sample.shape[0] * len(times) / sum(times) / 1e6
    print(json.dumps({
        "metric": "Mrays/s, primary rays (closest hit, bit-exact prim ids)", "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": round(1000 * sum(times) / len(times), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "C2: sponza262k stand-in, every 64th of 1920x1080 primary rays per step (bounded sample)"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": f"{sample.shape[0]} rays per step; oracle/_ref absent, CPU oracle port used"},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


if __name__ == "__main__":
    main()
