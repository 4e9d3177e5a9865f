#!/usr/bin/env python
"""Benchmarks of the irregular-grid path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 — config C2, the one BASELINE.json's target is quoted on: procedural Sponza-class scene (262 267 triangles),
1920x1080 primary rays, --top-density 0.15 --snd-density 3.0. A step is one traverse_grid pass over one frame
(2 073 600 rays), timed with a CUDA event pair on the launching (legacy default) stream exactly like the reference's
profile() (src/profile.cu:5-18, src/main.cpp:414-441); the L2 is flushed between steps.
  value  device-resident: rays and hits stay in HBM (the reference's own metric)
  e2e    the same frame through the C ABI with HOST buffers: pinned H2D of the rays, traversal, D2H of the hits, all
         inside the timed region
The line also carries the one-GPU figure of the multi-GPU workload (`c5_frame`), so that a 1 -> N series has its base.

N > 1 (torchrun, one process per GPU, NCCL) — config C5, the one BASELINE.json shards over GPUs: 7.8 M-triangle scene,
one 1920x1080 frame of primary rays + one diffuse bounce per step, STRONG scaling: the frame's 4-row tile bands are
dealt round-robin over the ranks (hagrid_b200.sharding.interleaved_bands), every rank owns a replica of the grid
(construction does not shard: replicas only), traces its share of the first wave, makes its second wave on the device
(keyed by the rays' indices in the whole frame, so the frame does not depend on N), traces it, and joins ONE all-reduce
of the frame's hit counters — all inside the timed region. value = rays of both waves x K / sum of step times, max over
ranks. No data-path collective. Secondary keys: `one_gpu_same_workload` (rank 0 alone, same run), `replicas` (C2 frame
on every rank, the weak-scaling figure of round 1), `pcie_floor_ms` (all ranks copying their shares at once).

--impl reference runs cg-saarland/hagrid itself: it has no CPU build/traverse path and no multi-GPU path, so its CUDA
sources rebuilt for sm_100a (oracle/_ref, see oracle/build_ref.sh) are driven through the same C ABI on rank 0's GPU,
on the same workload as this library's arm at that N (the second wave's rays are pre-generated for it: it has no
bounce stage). If that build is absent the CPU oracle port is timed on a bounded sample instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION = 0.15, 3.0, 0.995, 3
WIDTH, HEIGHT = 1920, 1080
BOUNCE_SEED = 7
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
               "clocks_event_reasons.sw_power_cap")
HIT = np.dtype([("id", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
C2_WORKLOAD = "C2: sponza262k stand-in (262267 tris), 1920x1080 primary rays, -td 0.15 -sd 3.0 -a 0.995 -e 3"
C5_WORKLOAD = ("C5: sanmiguel7p8m stand-in (7800000 tris), one 1920x1080 frame = primary rays + one diffuse bounce "
               "(second wave made on the device), -td 0.15 -sd 3.0 -a 0.995 -e 3")


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


def crc_of(hits: np.ndarray) -> dict:
    """CRC32 of the id column and of the t column (as stored bits) of a hit buffer: lets the two arms' lines be compared offline."""
    hits = np.ascontiguousarray(hits).view(HIT).reshape(-1)
    return {"id": format(zlib.crc32(np.ascontiguousarray(hits["id"]).tobytes()), "08x"),
            "t": format(zlib.crc32(np.ascontiguousarray(hits["t"]).tobytes()), "08x")}


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


class Ctx:
    """What every measurement needs: torch, the process group, the L2 flush buffer."""

    def __init__(self, torch, dist, rank, world, steps, warmup):
        self.torch, self.dist, self.rank, self.world, self.steps, self.warmup = torch, dist, rank, world, steps, warmup
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step, steps, warmup, collective=True):
        """`steps` event-timed steps after `warmup` untimed ones, L2 flushed before each; returns ms per step (this rank)."""
        torch = self.torch
        for _ in range(warmup):
            self.flush.zero_(); step()
        if collective:
            self.barrier()
        else:
            torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            self.flush.zero_()
            starts[i].record()
            step()
            stops[i].record()
        torch.cuda.synchronize()
        if collective:
            self.barrier()
        return np.array([a.elapsed_time(b) for a, b in zip(starts, stops)], dtype=np.float64)

    def max_over_ranks(self, *values):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def device_rays(torch, rays):
    n = rays.shape[0]
    return torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(), torch.empty((n, 4), dtype=torch.float32, device="cuda")


# ----------------------------------------------------------------------------------------------------------------------
# The two-wave frame (C5): this library sharded over the ranks, or the reference on one GPU
# ----------------------------------------------------------------------------------------------------------------------
def two_wave_frames(ctx: Ctx, lib, scene, tris, primary, reference, steps, warmup, rank, world, collective=True):
    """Times `steps` two-wave frames; rank r of `world` traces the tile bands r, r + world, ... of the frame.
    Returns per-rank numpy hit buffers of both waves (shard order), step times and e2e step times in ms."""
    from hagrid_b200 import HIT_PRIM_ID, scenes, sharding
    torch, dist = ctx.torch, ctx.dist
    total = primary.shape[0]
    idx = sharding.interleaved_bands(total, rank, world, sharding.raster_granule(WIDTH)) if world > 1 else np.arange(total, dtype=np.int64)
    mine = np.ascontiguousarray(primary[idx])
    n = mine.shape[0]
    lo, hi = scenes.scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    offset, tmax = 1e-3 * diag, diag
    d_rays, d_hits1 = device_rays(torch, mine)
    d_hits2 = torch.empty_like(d_hits1)
    h_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).pin_memory()
    h_hits1 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    h_hits2 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    launches0 = lib.kernel_launches()
    if reference:
        # the reference has no second-wave stage: its rays are made beforehand on the CPU (bit for bit what this
        # library's device kernel makes, tests/test_bounce.py) and wait in device / page-locked memory
        scene.traverse(d_rays, d_hits1, n, HIT_PRIM_ID)
        first = d_hits1.cpu().numpy().view(HIT).reshape(-1)
        bounce = scenes.bounce_rays_f32(tris, mine, first, offset, tmax, BOUNCE_SEED)
        d_bounce, _ = device_rays(torch, bounce)
        h_bounce = torch.from_numpy(bounce.view(np.float32).reshape(n, 8)).pin_memory()

        def step():
            scene.traverse(d_rays, d_hits1, n, HIT_PRIM_ID)
            scene.traverse(d_bounce, d_hits2, n, HIT_PRIM_ID)

        def host_step():
            lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_rays.data_ptr(), h_hits1.data_ptr(), n, HIT_PRIM_ID), "e2e")
            lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_bounce.data_ptr(), h_hits2.data_ptr(), n, HIT_PRIM_ID), "e2e")
    else:
        d_keys = torch.from_numpy(idx.astype(np.int32)).cuda()
        d_bounce = torch.empty_like(d_rays)
        counters = torch.zeros(2, dtype=torch.int64, device="cuda")

        def step():
            counters.zero_()
            # primary trace, hit count, bounce rays, second trace, hit count: one call, chunked over two streams
            scene.trace_two_waves(d_rays, n, d_keys, offset, tmax, BOUNCE_SEED, d_hits1, d_bounce, d_hits2, counters)
            if world > 1 and collective:
                dist.all_reduce(counters)               # the frame's only collective: 16 bytes

        def host_step():
            scene.trace_two_waves_host(h_rays, n, d_keys, offset, tmax, BOUNCE_SEED, h_hits1, h_hits2)

    step_ms = ctx.timed(step, steps, warmup, collective)
    launches = lib.kernel_launches() - launches0
    reduced = None if reference else [int(v) for v in counters.cpu()]
    hits1 = d_hits1.cpu().numpy().view(HIT).reshape(-1).copy()
    hits2 = d_hits2.cpu().numpy().view(HIT).reshape(-1).copy()
    e2e_steps = max(3, min(steps, 20))
    e2e_ms = ctx.timed(host_step, e2e_steps, 2, collective)
    e2e_ok = bool(np.array_equal(h_hits1.numpy().view(HIT).reshape(-1)["id"], hits1["id"]) and
                  np.array_equal(h_hits2.numpy().view(HIT).reshape(-1)["id"], hits2["id"]))
    return {"idx": idx, "n": n, "step_ms": step_ms, "e2e_ms": e2e_ms, "hits1": hits1, "hits2": hits2, "e2e_ok": e2e_ok,
            "launches_per_step": launches / float(steps + warmup), "counters": reduced}


def gather_frame(ctx: Ctx, part: np.ndarray, idx: np.ndarray, total: int):
    """The whole frame's hit buffer on rank 0 (outside any timed region); None elsewhere."""
    torch, dist = ctx.torch, ctx.dist
    if ctx.world == 1:
        return part
    sizes = [None] * ctx.world
    dist.all_gather_object(sizes, int(part.shape[0]))
    longest = max(sizes)
    padded = torch.zeros((longest, 4), dtype=torch.float32, device="cuda")
    padded[: part.shape[0]] = torch.from_numpy(part.view(np.float32).reshape(-1, 4)).cuda()
    idx_padded = torch.full((longest,), -1, dtype=torch.int64, device="cuda")
    idx_padded[: idx.shape[0]] = torch.from_numpy(idx).cuda()
    parts = [torch.empty_like(padded) for _ in range(ctx.world)] if ctx.rank == 0 else None
    idxs = [torch.empty_like(idx_padded) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(padded, parts, dst=0)
    dist.gather(idx_padded, idxs, dst=0)
    if ctx.rank != 0:
        return None
    whole = np.empty(total, HIT)
    for p, i in zip(parts, idxs):
        i = i.cpu().numpy()
        keep = i >= 0
        whole[i[keep]] = p.cpu().numpy().view(HIT).reshape(-1)[keep]
    return whole


def pcie_floor(ctx: Ctx, n: int, bytes_up_per_ray=32, bytes_down_per_ray=32, reps=5):
    """All ranks copy a frame share's worth of bytes up and down at once (two streams): ms per frame, max over ranks."""
    torch = ctx.torch
    up_h = torch.empty(n * bytes_up_per_ray, dtype=torch.uint8).pin_memory()
    up_d = torch.empty(n * bytes_up_per_ray, dtype=torch.uint8, device="cuda")
    dn_h = torch.empty(n * bytes_down_per_ray, dtype=torch.uint8).pin_memory()
    dn_d = torch.empty(n * bytes_down_per_ray, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(reps + 1):
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a); s2.wait_event(a)
        with torch.cuda.stream(s1):
            up_d.copy_(up_h, non_blocking=True)
        with torch.cuda.stream(s2):
            dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return ctx.max_over_ranks(best)[0]


# ----------------------------------------------------------------------------------------------------------------------
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
    launched = max(world, 1)

    affinity = bind_near_gpu(local_rank) if world > 1 else "not bound (single rank)"
    import torch
    import torch.distributed as dist
    from hagrid_b200 import Library

    ref_lib_path = ROOT / "oracle" / "_ref" / "libhagrid_ref.so"
    if reference and not ref_lib_path.exists():
        return reference_on_cpu(args, rank)
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
    ctx = Ctx(torch, dist, rank, world, args.steps, args.warmup)
    if launched == 1:
        line = bench_single_gpu(ctx, lib, args, reference, local_rank, affinity)
    else:
        line = bench_sharded_frame(ctx, lib, args, reference, local_rank, affinity, launched)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# N = 1: config C2
# ----------------------------------------------------------------------------------------------------------------------
def bench_single_gpu(ctx: Ctx, lib, args, reference, local_rank, affinity):
    from hagrid_b200 import HIT_PRIM_ID, Scene, scenes
    torch = ctx.torch
    tris = scenes.sponza262k()
    rays = camera_path_view(tris, scenes, 0)
    n = rays.shape[0]

    # ---- construction (reported, not the headline)
    scene = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
    scene.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=False, warmup=3, iters=0)
    launches0 = lib.kernel_launches()
    build_ms = scene.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=False, warmup=0, iters=10)
    build_launches = (lib.kernel_launches() - launches0) / 10.0
    scene.setup_traversal()
    info = scene.info().as_dict()

    # ---- device-resident frames: torch owns the ray / hit buffers
    d_rays, d_hits = device_rays(torch, rays)
    sampler = ClockSampler(local_rank)
    launches0 = lib.kernel_launches()
    step_ms = ctx.timed(lambda: scene.traverse(d_rays, d_hits, n, HIT_PRIM_ID), args.steps, args.warmup)
    launches = lib.kernel_launches() - launches0
    hits = d_hits.cpu().numpy().view(HIT).reshape(-1)
    # the same steps with the tiles handed out in buffer order (what the first two launches on a new ray buffer cost:
    # from the third on the kernel orders its tiles by what they cost before) -- reported beside the headline
    unordered = None
    if not reference:
        lib.set_option("tile_order", 0)
        u_steps = max(3, min(args.steps, 50))
        u_ms = ctx.timed(lambda: scene.traverse(d_rays, d_hits, n, HIT_PRIM_ID), u_steps, 3)
        lib.set_option("tile_order", 8)
        unordered = {"value": round(n * u_steps / (1000.0 * float(u_ms.sum())), 1), "unit": "Mrays/s",
                     "ms_per_step": round(float(u_ms.mean()), 5), "steps": u_steps,
                     "hits_identical": bool(np.array_equal(d_hits.cpu().numpy().view(HIT).reshape(-1), hits)),
                     "what": "hgb_set_option('tile_order', 0): tiles in buffer order, no use of earlier launches"}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
    h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 100))
    e2e_ms = ctx.timed(lambda: lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "e2e"),
                       e2e_steps, 2)
    e2e_hits = h_hits.numpy().view(HIT).reshape(-1).copy()
    clocks = sampler.stop()                               # sampled across both timed loops

    # ---- secondary number: one viewer frame (camera -> BGRA image in host memory), src/main.cpp:598-621
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
    if not os.environ.get("HGB_BENCH_SKIP_C3"):
        sc3 = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
        sc3.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=True)
        sc3.setup_traversal()
        r3 = scenes.random_rays(tris, 1 << 22)
        d3, h3 = device_rays(torch, r3)
        ms3 = sc3.traverse_timed(d3, h3, r3.shape[0], HIT_PRIM_ID, warmup=3, iters=10)
        inc = {"incoherent_mrays_s": round(r3.shape[0] * len(ms3) / (1000.0 * float(ms3.sum())), 1),
               "incoherent_workload": "C3: 4194304 random rays, --compress",
               "incoherent_hits_crc": crc_of(h3.cpu().numpy())}
        sc3.close()
        del d3, h3

    # ---- secondary numbers: the multi-GPU workload on this one GPU (the base of a 1 -> N series of `bench.py --gpus N`)
    c5 = {}
    if not os.environ.get("HGB_BENCH_SKIP_C5"):
        tris5 = scenes.sanmiguel7p8m()
        sc5 = Scene(tris5, device=local_rank, keep_alive=True, lib=lib)
        ms5 = sc5.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=False, warmup=2, iters=3)
        sc5.setup_traversal()
        primary5 = scenes.default_view(tris5, WIDTH, HEIGHT)
        k5 = max(3, min(args.steps, 20))
        r5 = two_wave_frames(ctx, lib, sc5, tris5, primary5, reference, k5, 3, 0, 1)
        rays5 = 2 * primary5.shape[0]
        c5 = {"c5_frame": {"workload": C5_WORKLOAD, "value": round(rays5 * k5 / (1000.0 * float(r5["step_ms"].sum())), 1), "unit": "Mrays/s",
                           "ms_per_step": round(float(r5["step_ms"].mean()), 4), "steps": k5,
                           "e2e": {"value": round(rays5 * len(r5["e2e_ms"]) / (1000.0 * float(r5["e2e_ms"].sum())), 1), "unit": "Mrays/s",
                                   "ms_per_step": round(float(r5["e2e_ms"].mean()), 4),
                                   "h2d_bytes_per_step": (64 if reference else 32) * primary5.shape[0], "d2h_bytes_per_step": 32 * primary5.shape[0]},
                           "build_ms": round(float(ms5.mean()), 2),
                           "hits_crc": {"primary": crc_of(r5["hits1"]), "bounce": crc_of(r5["hits2"])},
                           "note": "the workload of `bench.py --gpus N` (N > 1) on one GPU: divide that line's value by N x this value for "
                                   "the strong-scaling efficiency" + ("; second-wave rays pre-generated (the reference has no bounce stage)" if reference else "")}}
        sc5.close()

    total_ms, e2e_total = float(step_ms.sum()), float(e2e_ms.sum())
    value = n * args.steps / (1000.0 * total_ms)
    e2e_value = n * e2e_steps / (1000.0 * e2e_total)
    peak, peak_src = measured_peak_gbs()
    algo_bytes = 48 * n + scene_bytes(info, tris.shape[0])
    launch_ms = float(np.mean(step_ms))
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "r02_traverse_dram.json").read_text())[args.impl]["dram_bytes_per_launch"]
    except Exception:
        pass
    build_bytes = int(48 * tris.shape[0] + 4 * info["num_entries"] + 32 * info["num_cells"] + 4 * info["num_refs"])
    line = {
        "metric": "Mrays/s, primary rays (closest hit, bit-exact prim ids)", "value": round(value, 1), "unit": "Mrays/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": args.impl,
        "config": {"workload": C2_WORKLOAD, "rays_per_step_per_gpu": n,
                   "grid": {k: info[k] for k in ("dims", "shift", "num_cells", "num_entries", "num_refs")},
                   "l2": "256 MiB memset between steps (L2 flushed); scene+rays+hits = %.1f MB" % (algo_bytes / 1e6),
                   "timing": "CUDA event pair per step on the legacy default stream, sum over steps",
                   "tile_order": "steps trace the same ray buffer: from the third launch on the tile kernel hands out the buffer's tiles by what "
                                 "they cost on earlier launches (all rays traced every step, hits unchanged); `tiles_in_buffer_order` is the "
                                 "same measurement without it"},
        "e2e": {"value": round(e2e_value, 1), "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 16 * n,
                "ms_per_step": round(e2e_total / e2e_steps, 4), "steps": e2e_steps,
                "hits_match_device_path": bool(np.array_equal(e2e_hits["id"], hits["id"]) and np.array_equal(e2e_hits["t"], hits["t"])),
                "hits_crc": crc_of(e2e_hits),
                "api": "hgb_traverse_grid_host (pinned host buffers)", "host_affinity_rank0": affinity},
        "hits_crc": crc_of(hits),
        "tiles_in_buffer_order": unordered,
        "gpu_launches": int(launches) if not reference else 0,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "traverse_pid<Cell, Tri> (the only kernel of a step)" if reference else
                               "traverse_tiles<Cell, 1> (the only kernel of a step; every 16th step times its tiles and is followed, on a side "
                               "stream, by the 8 small kernels that sort the tiles by cost for the steps after it)",
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "gather kernel bound by instruction issue and dependent-load latency (ncu, profiles/r02b_traverse_primary_metrics.csv: "
                             "76 % of peak issue rate, 23.0 of 32 lanes active, DRAM 6 % of peak): compulsory HBM traffic is 48 B/ray + the part "
                             "of the scene the view touches; "
                             "traffic = dram read + write bytes of one launch under ncu with caches flushed (profiles/r02_traverse_dram.json); "
                             "see DESIGN.md section 5"},
        "build_ms": {"mean": round(float(build_ms.mean()), 3), "median": round(float(np.median(build_ms)), 3),
                     "min": round(float(build_ms.min()), 3), "iters": int(build_ms.shape[0]),
                     "kernel_launches_per_build": round(build_launches, 1),
                     "what": "build+merge+flatten+expand, keep-alive, event-timed like src/main.cpp:494-508"},
        "build_roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes": build_bytes,
                           "achieved": round(build_bytes / (float(build_ms.mean()) * 1e6), 2),
                           "note": "SURVEY 8(d): read the triangles once + write the final grid once, over the whole multi-pass pipeline "
                                   "(launch- and latency-bound at this size); per-kernel DRAM throughput of the 2 M-triangle build is in profiles/"},
        "hit_fraction": round(float((hits["id"] >= 0).mean()), 4),
    }
    line.update(inc)
    line.update(frame)
    line.update(c5)
    if reference:
        line["ranks_used"] = 1
        line["cpu_baseline"] = {"value": line["value"], "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                "sample": "whole workload; cg-saarland/hagrid has no CPU build/traverse path, so the arm runs its "
                                          "CUDA sources rebuilt for sm_100a (oracle/_ref) on the same GPU, one host thread"}
    elif not args.no_cpu_baseline:
        gi, e, c, r = scene.download()
        line["cpu_baseline"] = cpu_oracle_baseline(tris, gi.as_dict(), (e, c, r), rays)
    scene.close()
    return line


# ----------------------------------------------------------------------------------------------------------------------
# N > 1: config C5, one frame sharded over the ranks (strong scaling)
# ----------------------------------------------------------------------------------------------------------------------
def bench_sharded_frame(ctx: Ctx, lib, args, reference, local_rank, affinity, launched):
    from hagrid_b200 import HIT_PRIM_ID, Scene, scenes
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    tris = scenes.sanmiguel7p8m()
    primary = scenes.default_view(tris, WIDTH, HEIGHT)
    total = primary.shape[0]
    scene = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
    build_ms = scene.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION, compress=False, warmup=2, iters=3)
    scene.setup_traversal()
    info = scene.info().as_dict()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = two_wave_frames(ctx, lib, scene, tris, primary, reference, args.steps, args.warmup, rank, world)
    clocks = sampler.stop() if sampler else None
    total_ms, e2e_total = ctx.max_over_ranks(res["step_ms"].sum(), res["e2e_ms"].sum())
    e2e_steps = len(res["e2e_ms"])
    shard_sizes = [None] * world
    if world > 1:
        ctx.dist.all_gather_object(shard_sizes, int(res["n"]))
    else:
        shard_sizes = [int(res["n"])]
    floor_ms = None if reference else pcie_floor(ctx, max(shard_sizes))
    whole1 = gather_frame(ctx, res["hits1"], res["idx"], total)
    whole2 = gather_frame(ctx, res["hits2"], res["idx"], total)

    # ---- the same frame on rank 0's GPU alone, same run (the base of the strong-scaling figure)
    one, same_as_one = None, None
    if not reference:
        k1 = max(3, min(args.steps, 10))
        if rank == 0:
            r1 = two_wave_frames(ctx, lib, scene, tris, primary, False, k1, 3, 0, 1, collective=False)
            one = {"value": round(2 * total * k1 / (1000.0 * float(r1["step_ms"].sum())), 1), "unit": "Mrays/s",
                   "ms_per_step": round(float(r1["step_ms"].mean()), 4), "steps": k1,
                   "e2e_value": round(2 * total * len(r1["e2e_ms"]) / (1000.0 * float(r1["e2e_ms"].sum())), 1),
                   "e2e_ms_per_step": round(float(r1["e2e_ms"].mean()), 4)}
            same_as_one = bool(whole1.tobytes() == r1["hits1"].tobytes() and whole2.tobytes() == r1["hits2"].tobytes())
        ctx.barrier()

    # ---- the weak-scaling figure of round 1: the C2 frame on every rank (replicas, nothing shared)
    replicas = None
    if not reference and not os.environ.get("HGB_BENCH_SKIP_REPLICAS"):
        tris2 = scenes.sponza262k()
        sc2 = Scene(tris2, device=local_rank, keep_alive=True, lib=lib)
        sc2.build_all(TOP_DENSITY, SND_DENSITY, ALPHA, EXPANSION)
        sc2.setup_traversal()
        rays2 = camera_path_view(tris2, scenes, 0)
        d2, h2 = device_rays(torch, rays2)
        k2 = max(3, min(args.steps, 20))
        ms2 = ctx.timed(lambda: sc2.traverse(d2, h2, rays2.shape[0], HIT_PRIM_ID), k2, 3)
        worst = ctx.max_over_ranks(ms2.sum())[0]
        replicas = {"workload": C2_WORKLOAD + ", the same frame on every rank", "value": round(world * rays2.shape[0] * k2 / (1000.0 * worst), 1),
                    "unit": "Mrays/s", "ms_per_step": round(worst / k2, 5), "scaling": "weak"}
        sc2.close()

    if rank != 0:
        scene.close()
        return None
    rays_per_frame = 2 * total
    value = rays_per_frame * args.steps / (1000.0 * total_ms)
    e2e_value = rays_per_frame * e2e_steps / (1000.0 * e2e_total)
    peak, peak_src = measured_peak_gbs()
    # dominant kernel of a step: the second wave's traversal (incoherent rays). Algorithmic bytes of the whole step per
    # GPU: 48 B per ray and wave + 80 B per ray for the bounce kernel; the scene (1.3 GB) is not amortised over a shard
    algo_bytes = (48 * 2 + 80) * max(shard_sizes)
    achieved = algo_bytes / (total_ms / args.steps * 1e-3) / 1e9
    line = {
        "metric": "Mrays/s, primary + one-bounce rays of one frame (closest hit, bit-exact prim ids)", "value": round(value, 1), "unit": "Mrays/s",
        "n_gpus": launched, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 5),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": args.impl,
        "config": {"workload": C5_WORKLOAD, "rays_per_step": rays_per_frame, "rays_per_step_per_gpu": [2 * s for s in shard_sizes],
                   "sharding": ("the reference has no multi-GPU path: whole frame on rank 0's GPU" if reference else
                                f"4-row tile bands of the 1920x1080 frame dealt round-robin over {world} ranks (hagrid_b200.sharding.interleaved_bands); "
                                "grid replicated (construction does not shard: replicas only); one all-reduce of two 64-bit hit counters per frame, "
                                "inside the timed region; no data-path collective"),
                   "grid": {k: info[k] for k in ("dims", "shift", "num_cells", "num_entries", "num_refs")},
                   "l2": "256 MiB memset between steps (L2 flushed); the scene (1.3 GB) does not fit the L2 anyway",
                   "timing": "CUDA event pair per step on the legacy default stream, sum over steps, max over ranks; barrier + synchronize around the loop",
                   "note": "N = 1 runs config C2 (the one BASELINE.json's target is quoted on) and carries this workload's one-GPU figure as c5_frame"},
        "e2e": {"value": round(e2e_value, 1), "unit": "Mrays/s",
                "h2d_bytes_per_step": (64 if reference else 32) * total, "d2h_bytes_per_step": 32 * total,
                "ms_per_step": round(e2e_total / e2e_steps, 4), "steps": e2e_steps, "hits_match_device_path": res["e2e_ok"],
                "api": ("hgb_traverse_grid_host twice (second-wave rays uploaded)" if reference else
                        "hgb_trace_two_waves_host per rank (pinned host buffers; second wave made on the device)"),
                "pcie_floor_ms": None if floor_ms is None else round(floor_ms, 4),
                "pcie_floor": "all ranks copying their share up (32 B/ray) and down (32 B/ray) at once, best of 5",
                "host_affinity_rank0": affinity},
        "hits_crc": {"primary": crc_of(whole1), "bounce": crc_of(whole2)},
        "gpu_launches": 0 if reference else int(round(res["launches_per_step"] * args.steps)),
        "gpu_launches_per_step": None if reference else round(res["launches_per_step"], 2),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": None, "peak_source": peak_src, "kernel": "whole step of the slowest rank (two traversals + bounce kernel + counters)",
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "latency-bound gathers over a 1.3 GB scene: the fraction says how far a marching kernel is from streaming its rays"},
        "build_ms": {"mean": round(float(build_ms.mean()), 2), "what": "every rank builds its replica (build+merge+flatten+expand, keep-alive)"},
        "hit_fraction": round(float((whole1["id"] >= 0).mean()), 4),
    }
    if not reference:
        line["frame_counters_all_reduced"] = {"hits": res["counters"][0], "checksum": res["counters"][1],
                                              "expected_hits": int((whole1["id"] >= 0).sum() + (whole2["id"] >= 0).sum())}
        line["one_gpu_same_workload"] = one
        line["hits_identical_to_one_gpu"] = same_as_one
        line["replicas"] = replicas
    else:
        line["ranks_used"] = 1
        line["cpu_baseline"] = {"value": line["value"], "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                "sample": "whole workload on one GPU; the reference has no CPU build/traverse path and no second-wave stage "
                                          "(its second-wave rays are pre-generated and resident)"}
    scene.close()
    return line


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
