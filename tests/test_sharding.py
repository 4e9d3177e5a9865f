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
