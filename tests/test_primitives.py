"""Device-wide primitives on their own (the reference wraps CUB for them,
src/parallel.cuh:12-89): exclusive scan (int32 and packed 2 x 32-bit counters, single pass with decoupled look-back),
reduce, flagged partition in CUB's order, stable radix sort of pairs -- each against numpy, bit for bit, at the sizes
where tile logic can go wrong (0, 1, one element around a tile of 2048, 2048 tiles + 1 = one element around the
look-back window) and at construction scale."""
import numpy as np
import pytest

from hagrid_b200 import Scene, scenes

pytestmark = pytest.mark.gpu

SIZES = (0, 1, 2047, 2048, 2049, 65 * 2048 + 7, 2048 * 2048 + 1)


@pytest.fixture(scope="module")
def sc(lib):
    s = Scene(scenes.cornell32(), keep_alive=True, lib=lib)
    yield s
    s.close()


def dev(sc, host):
    host = np.ascontiguousarray(host)
    p = sc.device_alloc(max(host.nbytes, 16))
    if host.nbytes:
        sc.to_device(p, host)
    return p


@pytest.mark.parametrize("n", SIZES)
def test_exclusive_scan_int32(sc, lib, n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 9, size=n, dtype=np.int32)
    d_in, d_out = dev(sc, x), sc.device_alloc(4 * (n + 1))
    lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, d_in, n, 4, d_out), "scan")
    got = sc.to_host(np.empty(n + 1, np.int32), d_out)
    want = np.concatenate([[0], np.cumsum(x, dtype=np.int64)]).astype(np.int32)
    assert np.array_equal(got, want)
    if n:
        # in place: the output may alias the input (the radix sort scans its histogram that way)
        buf = sc.device_alloc(4 * (n + 1)); sc.to_device(buf, np.concatenate([x, [0]]).astype(np.int32))
        lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, buf, n, 4, buf), "scan in place")
        assert np.array_equal(sc.to_host(np.empty(n + 1, np.int32), buf), want)
        sc.device_free(buf)
    sc.device_free(d_in); sc.device_free(d_out)


@pytest.mark.parametrize("n", SIZES)
def test_exclusive_scan_packed_pairs(sc, lib, n):
    """Two 32-bit counters per 64-bit element, the form the build's fused partitions use (kept count, new count)."""
    rng = np.random.default_rng(n + 1)
    lo = rng.integers(0, 5, size=n, dtype=np.uint64)
    hi = rng.integers(0, 2, size=n, dtype=np.uint64)
    x = lo | (hi << np.uint64(32))
    d_in, d_out = dev(sc, x), sc.device_alloc(8 * (n + 1))
    lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, d_in, n, 8, d_out), "scan")
    got = sc.to_host(np.empty(n + 1, np.uint64), d_out)
    want = np.concatenate([[0], np.cumsum(lo)]).astype(np.uint64) | (np.concatenate([[0], np.cumsum(hi)]).astype(np.uint64) << np.uint64(32))
    assert np.array_equal(got, want)
    sc.device_free(d_in); sc.device_free(d_out)


def test_exclusive_scan_repeated_is_stable(sc, lib):
    """The look-back protocol under contention: many scans of a large array, every one identical."""
    n = 13_000_001
    x = np.random.default_rng(5).integers(0, 3, size=n, dtype=np.int32)
    want = np.concatenate([[0], np.cumsum(x, dtype=np.int64)]).astype(np.int32)
    d_in, d_out = dev(sc, x), sc.device_alloc(4 * (n + 1))
    for _ in range(20):
        lib.check(lib.dll.hgb_prim_exclusive_scan(sc._h, d_in, n, 4, d_out), "scan")
        assert np.array_equal(sc.to_host(np.empty(n + 1, np.int32), d_out), want)
    sc.device_free(d_in); sc.device_free(d_out)


@pytest.mark.parametrize("n", (0, 1, 255, 257, 1_000_003))
def test_reduce(sc, lib, n):
    rng = np.random.default_rng(n + 2)
    ints = rng.integers(-1000, 1000, size=n, dtype=np.int32)
    floats = (rng.standard_normal(n) * 100).astype(np.float32)
    d_i, d_f, d_out = dev(sc, ints), dev(sc, floats), sc.device_alloc(16)
    for op, src, want in ((0, d_i, np.int32(ints.sum(dtype=np.int64)) if n else np.int32(0)),
                          (1, d_i, ints.max() if n else np.int32(-2**31)),
                          (2, d_f, floats.min() if n else np.float32(np.inf)),
                          (3, d_f, floats.max() if n else np.float32(-np.inf))):
        lib.check(lib.dll.hgb_prim_reduce(sc._h, src, n, op, d_out), "reduce")
        got = sc.to_host(np.empty(1, np.int32 if op < 2 else np.float32), d_out)[0]
        assert got == want, (op, got, want)
    for p in (d_i, d_f, d_out):
        sc.device_free(p)


@pytest.mark.parametrize("n", (0, 1, 2, 2047, 2049, 300_001))
def test_partition_is_cubs_flagged_order(sc, lib, n):
    """Kept items first in input order, rejected items behind them in REVERSE input order: the order of
    cub::DevicePartition::Flagged that the reference's build relies on (src/build.cu:568-569)."""
    rng = np.random.default_rng(n + 3)
    x = rng.integers(0, 1 << 30, size=n, dtype=np.int32)
    flags = rng.integers(0, 2, size=n, dtype=np.int32)
    d_x, d_f, d_out = dev(sc, x), dev(sc, flags), sc.device_alloc(4 * max(n, 1))
    kept = lib.check(lib.dll.hgb_prim_partition(sc._h, d_x, d_f, n, d_out), "partition")
    got = sc.to_host(np.empty(n, np.int32), d_out)
    want = np.concatenate([x[flags != 0], x[flags == 0][::-1]])
    assert kept == int((flags != 0).sum()) and np.array_equal(got, want)
    for p in (d_x, d_f, d_out):
        sc.device_free(p)


@pytest.mark.parametrize("bits", (1, 8, 9, 23))
@pytest.mark.parametrize("n", (1, 31, 2048, 2049, 1_500_017))
def test_sort_pairs_is_stable(sc, lib, n, bits):
    """Stable on the low `bits` bits: equal keys keep their input order (what gives every cell its reference list in
    the reference's order, src/build.cu:686-695), for one pass, several passes and a partial last digit."""
    rng = np.random.default_rng(n * 31 + bits)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.int32)
    vals = np.arange(n, dtype=np.int32)
    d_k, d_v = dev(sc, keys), dev(sc, vals)
    lib.check(lib.dll.hgb_prim_sort_pairs(sc._h, d_k, d_v, n, bits), "sort")
    got_k, got_v = sc.to_host(np.empty(n, np.int32), d_k), sc.to_host(np.empty(n, np.int32), d_v)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(got_k, keys[order]) and np.array_equal(got_v, vals[order])
    sc.device_free(d_k); sc.device_free(d_v)


def test_sort_ignores_bits_above_the_requested_ones(sc, lib):
    n, bits = 100_000, 10
    rng = np.random.default_rng(9)
    keys = rng.integers(0, 1 << 20, size=n, dtype=np.int32)
    vals = np.arange(n, dtype=np.int32)
    d_k, d_v = dev(sc, keys), dev(sc, vals)
    lib.check(lib.dll.hgb_prim_sort_pairs(sc._h, d_k, d_v, n, bits), "sort")
    order = np.argsort(keys & ((1 << bits) - 1), kind="stable")
    assert np.array_equal(sc.to_host(np.empty(n, np.int32), d_v), vals[order])
    sc.device_free(d_k); sc.device_free(d_v)
