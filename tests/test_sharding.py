"""CPU tier: host-side multi-GPU logic (ray sharding + counter reduction), world_size 2 over gloo."""
import os
import socket
from pathlib import Path

import numpy as np
import pytest

from hagrid_b200 import sharding


@pytest.mark.parametrize("n,world,granule", [(0, 1, 128), (1, 4, 128), (2073600, 8, 1920 * 4), (4194304, 3, 128),
                                             (1000, 8, 128), (65536, 2, 32), (127, 2, 128)])
def test_shard_bounds_partition_exactly(n, world, granule):
    bounds = [sharding.shard_bounds(n, r, world, granule) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (a0, a1), (b0, b1) in zip(bounds, bounds[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in bounds]
    assert max(sizes) - min(sizes) <= granule
    for a, b in bounds[:-1]:
        assert b % granule == 0 or b == n


def test_raster_granule_keeps_tile_rows_whole():
    g = sharding.raster_granule(1920)
    for r in range(8):
        a, b = sharding.shard_bounds(1920 * 1080, r, 8, g)
        assert a % (1920 * 4) == 0 and (b % (1920 * 4) == 0 or b == 1920 * 1080)


@pytest.mark.parametrize("n,world,granule", [(2073600, 8, 1920 * 4), (2073600, 3, 1920 * 4), (1000, 4, 128), (0, 2, 64), (130, 2, 64)])
def test_interleaved_bands_partition_exactly(n, world, granule):
    parts = [sharding.interleaved_bands(n, r, world, granule) for r in range(world)]
    merged = np.sort(np.concatenate(parts))
    assert np.array_equal(merged, np.arange(n))
    for r, p in enumerate(parts):
        assert np.all(np.diff(p) > 0)
        assert np.all((p // granule) % world == r)                      # whole bands, dealt round-robin
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= granule


def test_bad_rank_rejected():
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import oracle
    from util import Golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden("cornell32")
    grid = oracle.Grid.from_arrays(*g.stage["expand"])
    a, b = sharding.shard_bounds(g.rays.shape[0], rank, world, granule=128)
    ids = grid.traverse(g.tris, g.rays[a:b], mode=1)      # the CPU oracle stands in for the GPU tracer here
    steps = grid.traverse(g.tris, g.rays[a:b], mode=0)
    local = sharding.frame_counters(ids["id"], steps["id"], device_ms=1.0 + rank)
    total = sharding.reduce_counters(local, dist)
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), ids["id"])
    np.save(os.path.join(out_dir, f"total_{rank}.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_trace_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from util import Golden
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = Golden("cornell32")
    want = g.hits["hits_cell_ids"]["id"]
    got = np.concatenate([np.load(tmp_path / f"ids_{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
    for r in range(world):
        total = np.load(tmp_path / f"total_{r}.npy")
        assert total[0] == float((want >= 0).sum())
        assert total[1] == float(g.hits["hits_cell_steps"]["id"].sum())
        assert total[2] == 2.0          # max over ranks of the per-rank device time


def _frame_worker(rank, world, port, out_dir):
    """One rank of the sharded two-wave frame of bench.py --gpus N, with the CPU oracle standing in for the GPU:
    bands dealt round-robin, second wave keyed by the rays' indices in the whole frame, one all-reduce of the counters."""
    import torch
    import torch.distributed as dist
    from hagrid_b200 import scenes
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris, primary, grid, offset, tmax = _small_frame(oracle, scenes)
    idx = sharding.interleaved_bands(primary.shape[0], rank, world, sharding.raster_granule(64))
    mine = np.ascontiguousarray(primary[idx])
    first = grid.traverse(tris, mine, mode=1)
    # keyed generation = the rows of the generator run over the whole frame with these hits in place
    everything = np.zeros(primary.shape[0], dtype=first.dtype); everything["id"] = -1
    everything[idx] = first
    bounce = np.ascontiguousarray(oracle.bounce_rays(tris, primary, everything, offset, tmax, 11)[idx])
    second = grid.traverse(tris, bounce, mode=1)
    counters = torch.tensor([int((first["id"] >= 0).sum() + (second["id"] >= 0).sum()),
                             int((first["id"].astype(np.int64) + 1).sum() + (second["id"].astype(np.int64) + 1).sum())])
    dist.all_reduce(counters)
    np.save(os.path.join(out_dir, f"first_{rank}.npy"), first); np.save(os.path.join(out_dir, f"second_{rank}.npy"), second)
    np.save(os.path.join(out_dir, f"idx_{rank}.npy"), idx); np.save(os.path.join(out_dir, f"counters_{rank}.npy"), counters.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _small_frame(oracle, scenes):
    tris = scenes.small_mixed(3000, seed=9)
    grid = oracle.Grid.build(tris, 0.15, 3.0)
    grid.merge(0.995); grid.flatten(); grid.expand(3)
    lo, hi = scenes.scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    primary = scenes.primary_rays(lo - (hi - lo) * 0.5, 0.5 * (lo + hi), (0, 1, 0), 50.0, 64, 32, 4 * diag)
    return tris, primary, grid, 1e-3 * diag, diag


def test_two_rank_sharded_two_wave_frame_equals_the_unsharded_frame(tmp_path):
    import torch.multiprocessing as mp
    from hagrid_b200 import scenes
    from oracle import oracle
    world = 2
    mp.spawn(_frame_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    tris, primary, grid, offset, tmax = _small_frame(oracle, scenes)
    first = grid.traverse(tris, primary, mode=1)
    second = grid.traverse(tris, oracle.bounce_rays(tris, primary, first, offset, tmax, 11), mode=1)
    got1, got2 = np.empty_like(first), np.empty_like(second)
    for r in range(world):
        idx = np.load(tmp_path / f"idx_{r}.npy")
        got1[idx] = np.load(tmp_path / f"first_{r}.npy"); got2[idx] = np.load(tmp_path / f"second_{r}.npy")
    assert got1.tobytes() == first.tobytes() and got2.tobytes() == second.tobytes()
    assert (first["id"] >= 0).sum() > 100
    for r in range(world):
        c = np.load(tmp_path / f"counters_{r}.npy")
        assert c[0] == (first["id"] >= 0).sum() + (second["id"] >= 0).sum()
        assert c[1] == (first["id"].astype(np.int64) + 1).sum() + (second["id"].astype(np.int64) + 1).sum()


def test_reduce_counters_without_process_group_is_identity():
    local = np.array([3.0, 10.0, 0.5])
    assert np.array_equal(sharding.reduce_counters(local, None), local)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["blocks", "bands"])
def test_frame_sharded_over_the_gpus_of_the_node_equals_the_unsharded_frame(mode):
    """NCCL tier (needs at least two GPUs): tools/gpu_sharded_frame.py under torchrun."""
    import json
    import subprocess
    import sys
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("one GPU")
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          str(root / "tools" / "gpu_sharded_frame.py"), "5", mode], capture_output=True, text=True, timeout=400)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == world and sum(line["shard_sizes"]) == line["rays"] and line["sharding"] == mode
    assert line["hits_identical_to_unsharded"] and line["host_path_identical"]
    assert line["hits_counted_by_all_reduce"] == line["hits_in_unsharded_trace"]
