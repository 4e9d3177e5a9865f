"""Second-wave rays (BASELINE.json config C5) and the headless image writer (row f4).

The reference's front end stops at primary rays (src/main.cpp:52-66), so there is no reference output to pin
against: the operation is defined in include/hagrid_b200.h (hgb_generate_bounce_rays). What is checked instead:
CPU tier: the oracle's C restatement against an independent numpy-float32 restatement written here, and against
          the geometric properties the definition promises (unit directions in the hemisphere of the normal that
          faces the ray, origins on the offset surface, cosine-weighted distribution, misses untouched);
GPU tier: the device kernel against the oracle bit for bit, and the two-wave pipeline (hits never leave the
          device) against the host-side procedure on both this library and the reference build."""
import numpy as np
import pytest

from hagrid_b200 import HIT_DTYPE, HIT_PRIM_ID, RAY_DTYPE, Scene, scenes
from hagrid_b200.api import save_image
from oracle import oracle
from util import Golden

f32 = np.float32


def mix32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xFFFFFFFF
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def bounce_one(tri, ray, hit, i, offset, tmax, seed):
    """The definition in include/hagrid_b200.h, one numpy float32 operation per rounding."""
    n = np.array([tri["nx"], tri["ny"], tri["nz"]], f32)
    length = np.sqrt(f32(f32(n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]))
    if not length > 0:
        return ray
    n = n / length
    d, o = ray["dir"], ray["org"]
    if f32(f32(n[0] * d[0] + n[1] * d[1]) + n[2] * d[2]) > 0:
        n = -n
    out = np.zeros((), RAY_DTYPE)
    out["org"] = (o + d * f32(hit["t"])) + n * f32(offset)
    base = mix32(seed ^ mix32(i))
    draw = lambda k: f32(mix32(base + k) >> 8) * f32(2.0 ** -24)
    dx = dy = s = f32(0)
    for k in range(8):
        x, y = f32(2) * draw(2 * k) - f32(1), f32(2) * draw(2 * k + 1) - f32(1)
        q = f32(x * x + y * y)
        if q < 1:
            dx, dy, s = x, y, q
            break
    dz = np.sqrt(f32(f32(1) - s))
    t1 = np.array([-n[2], 0, n[0]], f32) if abs(n[0]) > f32(0.9) else np.array([0, n[2], -n[1]], f32)
    t1 = t1 / np.sqrt(f32(f32(t1[0] * t1[0] + t1[1] * t1[1]) + t1[2] * t1[2]))
    t2 = np.array([n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]], f32)
    out["dir"] = (t1 * dx + t2 * dy) + n * dz
    out["tmin"], out["tmax"] = 0, tmax
    return out


def hits_of(g, key="hits_cell_ids"):
    h = np.ascontiguousarray(g.hits[key])
    return h.view(HIT_DTYPE).reshape(-1) if h.dtype != HIT_DTYPE else h


@pytest.mark.parametrize("name", ["soup800", "strands1500"])
def test_oracle_bounce_equals_the_float32_restatement(name):
    g = Golden(name)
    hits = hits_of(g)
    rays = g.rays.view(RAY_DTYPE).reshape(-1)
    pick = np.concatenate([np.nonzero(hits["id"] >= 0)[0][:150], np.nonzero(hits["id"] < 0)[0][:10]])
    got = oracle.bounce_rays(g.tris, rays, hits, 1e-3, 7.5, 0x48414752)
    for i in pick:
        want = bounce_one(g.tris[hits["id"][i]], rays[i], hits[i], int(i), 1e-3, 7.5, 0x48414752) if hits["id"][i] >= 0 else rays[i]
        assert got[i].tobytes() == np.asarray(want).tobytes(), i


@pytest.mark.parametrize("name", ["soup800", "strands1500", "cornell32"])
def test_host_numpy_restatement_equals_the_oracle(name):
    """scenes.bounce_rays_f32 (vectorised numpy float32, used by the benchmark tools as the host procedure)."""
    g = Golden(name)
    hits, rays = hits_of(g), g.rays.view(RAY_DTYPE).reshape(-1)
    for seed in (0, 7, 0x48414752):
        assert scenes.bounce_rays_f32(g.tris, rays, hits, 2e-3, 9.0, seed).tobytes() == \
            oracle.bounce_rays(g.tris, rays, hits, 2e-3, 9.0, seed).tobytes()


def test_oracle_bounce_properties():
    g = Golden("soup800")
    hits, rays = hits_of(g), g.rays.view(RAY_DTYPE).reshape(-1)
    out = oracle.bounce_rays(g.tris, rays, hits, 0.01, 3.0, 11)
    miss = hits["id"] < 0
    assert miss.any() and (~miss).any()
    assert out[miss].tobytes() == rays[miss].tobytes()
    o, r, h = out[~miss], rays[~miss], hits[~miss]
    tr = g.tris[h["id"]]
    n = np.stack([tr["nx"], tr["ny"], tr["nz"]], -1).astype(np.float64)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[np.sum(n * r["dir"], axis=1) > 0] *= -1
    d = o["dir"].astype(np.float64)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-6
    assert np.sum(d * n, axis=1).min() > -1e-6                      # hemisphere of the facing normal
    p = r["org"] + r["dir"].astype(np.float64) * h["t"][:, None] + n * 0.01
    assert np.abs(o["org"] - p).max() < 1e-4 * max(1.0, np.abs(p).max())
    assert (o["tmin"] == 0).all() and (o["tmax"] == f32(3.0)).all()
    assert out.tobytes() == oracle.bounce_rays(g.tris, rays, hits, 0.01, 3.0, 11).tobytes()
    other = oracle.bounce_rays(g.tris, rays, hits, 0.01, 3.0, 12)
    assert (other[~miss]["dir"] != o["dir"]).any(axis=1).mean() > 0.99
    # out-of-range ids count as misses
    bad = hits.copy(); bad["id"][~miss] = len(g.tris) + 5
    assert oracle.bounce_rays(g.tris, rays, bad, 0.01, 3.0, 11).tobytes() == rays.tobytes()


def test_oracle_bounce_is_cosine_weighted():
    """One triangle facing +z hit by 200 000 rays: E[cos] = 2/3, E[cos^2] = 1/2, azimuth uniform."""
    tris = scenes.make_tris(np.array([[0, 0, 0]], f32), np.array([[1, 0, 0]], f32), np.array([[0, 1, 0]], f32))
    n = 200_000
    rays = np.zeros(n, RAY_DTYPE); rays["org"] = (0.2, 0.2, 1); rays["dir"] = (0, 0, -1); rays["tmax"] = 10
    hits = np.zeros(n, HIT_DTYPE); hits["t"] = 1
    d = oracle.bounce_rays(tris, rays, hits, 0.0, 1.0, 3)["dir"].astype(np.float64)
    nz = abs(float(tris["nz"][0]))
    assert nz > 0
    cos = d[:, 2] * np.sign(d[:, 2].sum())
    assert cos.min() >= 0
    assert abs(cos.mean() - 2 / 3) < 4e-3 and abs((cos ** 2).mean() - 0.5) < 4e-3
    phi = np.arctan2(d[:, 1], d[:, 0])
    hist = np.histogram(phi, bins=16, range=(-np.pi, np.pi))[0]
    assert np.abs(hist / n - 1 / 16).max() < 3e-3


def test_image_writer_round_trip(tmp_path):
    """hgb_save_image is host code: BGRA words in, binary PPM (RGB) out."""
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    path = tmp_path / "frame.ppm"
    save_image(path, img)
    raw = path.read_bytes()
    head = b"P6\n53 37\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + 53 * 37 * 3
    rgb = np.frombuffer(raw[len(head):], np.uint8).reshape(37, 53, 3)
    assert np.array_equal(rgb, img[..., 2::-1])
    from hagrid_b200 import HagridError
    with pytest.raises(HagridError):
        save_image(tmp_path / "no_such_dir" / "x.ppm", img)


# ----------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["soup800", "strands1500", "cornell32"])
def test_device_bounce_is_bit_identical_to_the_oracle(lib, name):
    g = Golden(name)
    hits, rays = hits_of(g), g.rays.view(RAY_DTYPE).reshape(-1)
    sc = Scene(g.tris, lib=lib)
    lo, hi = scenes.scene_bbox(g.tris)
    diag = float(np.linalg.norm(hi - lo))
    for count in (len(rays), 1, 255, 257):
        for seed in (0, 0x48414752):
            want = oracle.bounce_rays(g.tris, rays[:count], hits[:count], 1e-3 * diag, diag, seed)
            got = sc.bounce_rays(rays[:count], hits[:count], 1e-3 * diag, diag, seed)
            assert got.tobytes() == want.tobytes(), (count, seed)
    assert sc.bounce_rays(rays[:0], hits[:0], 0.1, 1.0, 1).shape == (0,)
    sc.close()


@pytest.mark.gpu
def test_two_waves_on_the_device_equal_the_host_procedure(lib, ref_lib):
    """C5 shape: primary hits stay in HBM, the bounce kernel writes the second wave next to them, the second
    traversal reads it. Same hits as tracing the oracle's bounce rays, on this library and on the reference."""
    tris = scenes.atrium(60000, seed=5)
    sc = Scene(tris, lib=lib)
    sc.build_all(0.15, 3.0)
    sc.setup_traversal()
    lo, hi = scenes.scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    primary = scenes.default_view(tris, 640, 360)
    n = len(primary)
    d_rays, d_hits, d_second = sc.device_alloc(n * 32), sc.device_alloc(n * 16), sc.device_alloc(n * 32)
    sc.to_device(d_rays, primary)
    sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
    first = sc.to_host(np.empty(n, HIT_DTYPE), d_hits)
    sc.bounce_rays_device(d_rays, d_hits, n, 1e-3 * diag, diag, 99, d_second)          # out of place
    second = sc.to_host(np.empty(n, RAY_DTYPE), d_second)
    sc.traverse(d_second, d_hits, n, HIT_PRIM_ID)
    got = sc.to_host(np.empty(n, HIT_DTYPE), d_hits)
    sc.to_device(d_hits, first)
    sc.bounce_rays_device(d_rays, d_hits, n, 1e-3 * diag, diag, 99)                     # in place
    assert sc.to_host(np.empty(n, RAY_DTYPE), d_rays).tobytes() == second.tobytes()
    for p in (d_rays, d_hits, d_second):
        sc.device_free(p)

    assert (first["id"] >= 0).mean() > 0.5
    want_rays = oracle.bounce_rays(tris, primary, first, 1e-3 * diag, diag, 99)
    assert second.tobytes() == want_rays.tobytes()
    assert got.tobytes() == sc.trace(want_rays, HIT_PRIM_ID).tobytes()
    ref = Scene(tris, lib=ref_lib)
    ref.build_all(0.15, 3.0)
    ref.setup_traversal()
    want = ref.trace(want_rays, HIT_PRIM_ID)
    assert np.array_equal(got["id"], want["id"])
    assert np.abs(got["t"] - want["t"]).max() <= 1e-5 * diag
    assert got.tobytes() == want.tobytes()
    # the reference build has no such stage and says so
    from hagrid_b200 import HagridError
    with pytest.raises(HagridError):
        ref.bounce_rays(primary[:4], first[:4], 0.1, 1.0, 1)
    ref.close(); sc.close()


@pytest.mark.gpu
def test_a_sharded_frame_gets_the_second_wave_of_the_whole_frame(lib):
    """hgb_generate_bounce_rays_keyed: a rank that passes its rays' indices in the whole frame as keys gets exactly the
    rows of the unsharded second wave (the gathered frame must not depend on the number of GPUs);
    hgb_count_hits gives the counters the ranks all-reduce; hgb_trace_two_waves_host equals the device procedure."""
    import torch
    from hagrid_b200 import sharding
    tris = scenes.atrium(60000, seed=5)
    sc = Scene(tris, lib=lib)
    sc.build_all(0.15, 3.0)
    sc.setup_traversal()
    lo, hi = scenes.scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    W, H = 640, 360
    primary = scenes.default_view(tris, W, H)
    first = sc.trace(primary, HIT_PRIM_ID)
    whole = sc.bounce_rays(primary, first, 1e-3 * diag, diag, 5)
    second = sc.trace(whole, HIT_PRIM_ID)
    world = 3
    for rank in range(world):
        idx = sharding.interleaved_bands(len(primary), rank, world, sharding.raster_granule(W))
        mine = np.ascontiguousarray(primary[idx]); n = len(mine)
        d_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).cuda()
        d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        d_out = torch.empty_like(d_rays)
        d_keys = torch.from_numpy(idx.astype(np.int32)).cuda()
        counters = torch.zeros(2, dtype=torch.int64, device="cuda")
        sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
        sc.count_hits(d_hits, n, counters)
        sc.bounce_rays_keyed(d_rays, d_hits, n, 1e-3 * diag, diag, 5, d_keys, d_out)
        got = d_out.cpu().numpy().view(RAY_DTYPE).reshape(-1)
        assert got.tobytes() == np.ascontiguousarray(whole[idx]).tobytes(), rank
        ids = first["id"][idx]
        assert [int(v) for v in counters.cpu()] == [int((ids >= 0).sum()), int((ids.astype(np.int64) + 1).sum())]
        # the whole frame of this shard in one call (chunk chains on two streams): same buffers, same counters
        one_hits1, one_hits2, one_bounce = torch.zeros_like(d_hits), torch.zeros_like(d_hits), torch.zeros_like(d_rays)
        one_counters = torch.zeros(2, dtype=torch.int64, device="cuda")
        sc.trace_two_waves(d_rays, n, d_keys, 1e-3 * diag, diag, 5, one_hits1, one_bounce, one_hits2, one_counters)
        assert one_hits1.cpu().numpy().view(HIT_DTYPE).reshape(-1).tobytes() == np.ascontiguousarray(first[idx]).tobytes()
        assert one_bounce.cpu().numpy().view(RAY_DTYPE).reshape(-1).tobytes() == np.ascontiguousarray(whole[idx]).tobytes()
        assert one_hits2.cpu().numpy().view(HIT_DTYPE).reshape(-1).tobytes() == np.ascontiguousarray(second[idx]).tobytes()
        ids2 = second["id"][idx]
        assert [int(v) for v in one_counters.cpu()] == [int((ids >= 0).sum() + (ids2 >= 0).sum()),
                                                       int((ids.astype(np.int64) + 1).sum() + (ids2.astype(np.int64) + 1).sum())]
        # the host-buffer frame of this shard
        h_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).pin_memory()
        h1 = torch.empty((n, 4), dtype=torch.float32).pin_memory(); h2 = torch.empty((n, 4), dtype=torch.float32).pin_memory()
        sc.trace_two_waves_host(h_rays, n, d_keys, 1e-3 * diag, diag, 5, h1, h2)
        assert h1.numpy().view(HIT_DTYPE).reshape(-1).tobytes() == np.ascontiguousarray(first[idx]).tobytes()
        assert h2.numpy().view(HIT_DTYPE).reshape(-1).tobytes() == np.ascontiguousarray(second[idx]).tobytes()
    # without keys the stream is the ray's index in the buffer
    d_rays = torch.from_numpy(primary.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.from_numpy(first.view(np.float32).reshape(-1, 4)).cuda()
    d_out = torch.empty_like(d_rays)
    sc.bounce_rays_keyed(d_rays, d_hits, len(primary), 1e-3 * diag, diag, 5, None, d_out)
    assert d_out.cpu().numpy().view(RAY_DTYPE).reshape(-1).tobytes() == whole.tobytes()
    sc.close()
