"""Ray sharding over the GPUs of one node.

Traversal shards embarrassingly: rays are independent, so rank r of W traces a
contiguous block of the ray buffer against its own replica of grid + triangles and
writes the matching block of the hit buffer. There is no collective in the data
path; the only exchange is one all-reduce of a few counters per frame (hit count,
step total, slowest rank's device time). Construction does not shard (merge and
expand cross top-level cells, global scans/sort): every rank builds the same grid
deterministically ("replicas only").

Block boundaries are multiples of `granule` rays; with granule = 4 * W a W-pixel
raster keeps whole 8x4 pixel tile rows per rank, so the tile re-mapping of the
traversal kernel still applies inside every shard.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(num_rays: int, rank: int, world: int, granule: int = 128) -> tuple[int, int]:
    """[begin, end) of rank's block: blocks are contiguous, cover [0, num_rays) exactly once,
    differ by at most one granule, and every boundary but the last is a multiple of `granule`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    units = (num_rays + granule - 1) // granule
    base, extra = divmod(units, world)
    first = rank * base + min(rank, extra)
    last = first + base + (1 if rank < extra else 0)
    return min(first * granule, num_rays), min(last * granule, num_rays)


def interleaved_bands(num_rays: int, rank: int, world: int, granule: int) -> np.ndarray:
    """Indices of rank's rays when bands of `granule` rays are dealt round-robin (band k goes to rank k % world):
    for camera rasters with granule = raster_granule(width) every rank gets 4-row bands from all over the image, which
    evens out cheap and expensive regions (contiguous blocks leave the rank with the expensive part of the image as the
    slowest). The rank's rays, stored in this order, are again a raster of the same width."""
    if world <= 0 or not (0 <= rank < world) or granule <= 0:
        raise ValueError("bad rank/world/granule")
    units = (num_rays + granule - 1) // granule
    parts = [np.arange(u * granule, min((u + 1) * granule, num_rays), dtype=np.int64) for u in range(rank, units, world)]
    return np.concatenate(parts) if parts else np.empty(0, dtype=np.int64)


def raster_granule(width: int, tile_h: int = 4) -> int:
    """Granule that keeps rank boundaries on tile-row boundaries of a `width`-pixel raster."""
    return width * tile_h


def frame_counters(hit_ids: np.ndarray, steps: np.ndarray | None, device_ms: float) -> np.ndarray:
    """[hits, step total, device ms] of one rank, as float64 for the all-reduce."""
    return np.array([float((hit_ids >= 0).sum()), float(steps.sum()) if steps is not None else 0.0, float(device_ms)],
                    dtype=np.float64)


def reduce_counters(local: np.ndarray, dist=None) -> np.ndarray:
    """Sum of hits and steps, max of device time over all ranks (identity without a process group)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return local.copy()
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    sums = torch.tensor(local[:2], dtype=torch.float64, device=dev)
    worst = torch.tensor(local[2:], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return np.concatenate([sums.cpu().numpy(), worst.cpu().numpy()])
