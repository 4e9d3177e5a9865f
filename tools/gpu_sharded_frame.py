"""One frame sharded over the GPUs of a node (run under torchrun, one rank per GPU, NCCL):
every rank builds its replica of the grid, traces its block of the ray buffer (boundaries on 4-row tile bands) and
joins one all-reduce of three counters; rank 0 gathers the hit blocks and compares them with the unsharded trace.
Prints one JSON line: frame time (max over ranks, CUDA events) device-resident and through host buffers.
usage: torchrun --nproc-per-node N tools/gpu_sharded_frame.py [frames] [blocks|bands]
blocks: rank r traces one contiguous block of the frame; bands: 4-row bands dealt round-robin (sharding.interleaved_bands)."""
import json, os, sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_DTYPE, HIT_PRIM_ID, Library, Scene, scenes, sharding

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 50
mode = sys.argv[2] if len(sys.argv) > 2 else "blocks"
rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
lib = Library()
tris = scenes.sponza262k()
W, H = 1920, 1080
rays = scenes.default_view(tris, W, H)
scene = Scene(tris, device=local_rank, keep_alive=True, lib=lib)
scene.build_all(0.15, 3.0)
scene.setup_traversal()
def indices(r):
    if mode == "bands":
        return sharding.interleaved_bands(rays.shape[0], r, world, sharding.raster_granule(W))
    a, b = sharding.shard_bounds(rays.shape[0], r, world, granule=sharding.raster_granule(W))
    return np.arange(a, b, dtype=np.int64)


mine = np.ascontiguousarray(rays[indices(rank)])
n = mine.shape[0]

d_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).cuda()
d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
h_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).pin_memory()
h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(frames):
        fn()
    t1.record(); t1.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / frames], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0])


resident_ms = timed(lambda: scene.traverse(d_rays, d_hits, n, HIT_PRIM_ID))
host_ms = timed(lambda: lib.check(lib.dll.hgb_traverse_grid_host(scene._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "host frame"))
hits = d_hits.cpu().numpy().view(HIT_DTYPE).reshape(-1)
same_host = bool(np.array_equal(h_hits.numpy().view(HIT_DTYPE).reshape(-1), hits))
counters = sharding.reduce_counters(sharding.frame_counters(hits["id"], None, resident_ms), dist if world > 1 else None)

# gather the blocks on rank 0 (equal-sized padding; the data path itself has no collective)
owned = [indices(r) for r in range(world)]
longest = max(len(o) for o in owned)
padded = torch.zeros((longest, 4), dtype=torch.float32, device="cuda"); padded[:n] = d_hits
if world > 1:
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, parts, dst=0)
else:
    parts = [padded]
if rank == 0:
    whole = np.empty(rays.shape[0], HIT_DTYPE)
    for p, o in zip(parts, owned):
        whole[o] = p.cpu().numpy()[: len(o)].view(HIT_DTYPE).reshape(-1)
    want = scene.trace(rays, HIT_PRIM_ID)
    print(json.dumps({"n_gpus": world, "rays": int(rays.shape[0]), "sharding": mode, "shard_sizes": [len(o) for o in owned],
                      "frame_ms_device_resident": round(resident_ms, 4), "mrays_s_device_resident": round(rays.shape[0] / resident_ms / 1e3, 1),
                      "frame_ms_host_buffers": round(host_ms, 4), "mrays_s_host_buffers": round(rays.shape[0] / host_ms / 1e3, 1),
                      "hits_identical_to_unsharded": bool(whole.tobytes() == want.tobytes()), "host_path_identical": same_host,
                      "hits_counted_by_all_reduce": int(counters[0]), "hits_in_unsharded_trace": int((want["id"] >= 0).sum())}), flush=True)
scene.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
