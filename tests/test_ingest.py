"""Scene ingest: OBJ -> Tri array with the reference's semantics
(src/load_obj.cpp:78-239, load_model src/main.cpp:246-275).
CPU tier: the sequential oracle (oracle/obj_oracle.py) and this library's parallel host parser against golden
vectors the reference's own load_model produced (tests/golden/ingest.npz).
GPU tier: the full ingest (parallel parse + triangle setup on the device) against the oracle and, where the
reference build is present, against the reference's loader on a large file."""
from pathlib import Path

import numpy as np
import pytest

from hagrid_b200 import HagridError, Scene, parse_obj, scenes
from oracle import obj_oracle

Z = np.load(Path(__file__).resolve().parent / "golden" / "ingest.npz")
NAMES = sorted({k.rsplit("_", 1)[0] for k in Z.files})
ROOT = Path(__file__).resolve().parent.parent


def write(tmp_path, name):
    path = tmp_path / f"{name}.obj"
    path.write_bytes(Z[f"{name}_obj"].tobytes())
    return path


def tris_from(verts, idx):
    """load_model's arithmetic in numpy float32 (one rounding per operation, like the x86 front end)."""
    v0, v1, v2 = verts[idx[:, 0]], verts[idx[:, 1]], verts[idx[:, 2]]
    e1, e2 = v0 - v1, v2 - v0
    t = np.empty(idx.shape[0], dtype=obj_oracle.TRI_DTYPE)
    t["v0"], t["e1"], t["e2"] = v0, e1, e2
    t["nx"] = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    t["ny"] = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    t["nz"] = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    return t


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_the_reference_loader(tmp_path, name):
    got = obj_oracle.load_model(write(tmp_path, name))
    if not bool(Z[f"{name}_ok"]):
        assert got is None
    else:
        assert got.tobytes() == Z[f"{name}_tris"].tobytes()


@pytest.mark.parametrize("threads", [1, 3, 16])
@pytest.mark.parametrize("name", NAMES)
def test_parallel_host_parser_matches_the_reference_loader(tmp_path, name, threads):
    path = write(tmp_path, name)
    if not bool(Z[f"{name}_ok"]):
        with pytest.raises(HagridError):
            parse_obj(path, threads)
        return
    verts, idx = parse_obj(path, threads)
    assert tris_from(verts, idx).tobytes() == Z[f"{name}_tris"].tobytes()


def test_number_parsing_equals_strtof_on_adversarial_input(tmp_path):
    """The host parser converts plain decimals itself (one exact double operation, then float) and leaves the rest to
    strtof; the oracle calls the C library's strtof for every number like the reference (src/load_obj.cpp:119-122).
    Same bits and same end-of-number positions on: random decimals of every length, floats' exact half-way points
    (where rounding twice would differ from rounding once), range limits, and everything strtof accepts besides."""
    import struct
    rng = np.random.default_rng(12)
    tokens = ["inf", "-inf", "nan", "0x1p3", "-0x1.8p1", "1e", "1e+", "2.5e-", ".", "-.5", "+.5e1", "5.", "1e400", "-1e400",
              "1e-400", "1e-40", "1.1754944e-38", "1.1754942e-38", "3.4028235e38", "3.4028236e38", "3.5e38", "0", "-0", "-0.0e5",
              "00012.5000", "1E5", "1d5", "1_000", "1e0005", "1e99999999999", "12345678901234567890", "1234567890123456789",
              "0.00000000000000000001234", "123456789012345678901234567890e-25", "9007199254740993", "9007199254740992e-3",
              "8388608.5", "8388609.5", "16777217", "16777219", "33554434", "0.1", "0.3", "1e23", "1e22", "1e-22", "1e-23"]
    for _ in range(1500):                                     # random decimals: 1-21 digits, point anywhere, optional exponent
        digits = "".join(rng.choice(list("0123456789"), size=int(rng.integers(1, 22))))
        cut = int(rng.integers(0, len(digits) + 1))
        text = digits[:cut] + "." + digits[cut:] if rng.random() < 0.8 else digits
        if rng.random() < 0.4:
            text += rng.choice(["e", "E"]) + rng.choice(["", "+", "-"]) + str(int(rng.integers(0, 45)))
        tokens.append(("-" if rng.random() < 0.3 else "") + text)
    for _ in range(1500):                                     # exact half-way points between neighbouring floats, and their neighbours
        f = np.float32(rng.uniform(-1, 1) * 10.0 ** rng.integers(-6, 9))
        g = np.nextafter(f, np.float32(np.inf))
        mid = (float(f) + float(g)) / 2                       # exact in double
        tokens += [repr(mid), "%.17g" % mid, "%.9g" % mid, "%.8g" % float(f), "%.9g" % float(g)]
        from decimal import Decimal
        tokens.append(format(Decimal(mid), "f"))              # the half-way point written out in full
    while len(tokens) % 3:
        tokens.append("1")
    path = tmp_path / "numbers.obj"
    with open(path, "w") as f:
        for i in range(0, len(tokens), 3):
            f.write("v %s %s %s\n" % tuple(tokens[i:i + 3]))
        f.write("f 1 2 3\n")
    lines = len(tokens) // 3
    ref = np.array([obj_oracle._strtof3((" %s %s %s" % tuple(tokens[3 * i:3 * i + 3])).encode()) for i in range(lines)],
                   dtype=np.float32).view(np.uint32)
    for threads in (1, 5):
        verts, idx = parse_obj(path, threads)
        got = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)[-lines:].view(np.uint32)
        bad = np.nonzero((got != ref).any(axis=1))[0]
        assert len(bad) == 0, (tokens[3 * bad[0]: 3 * bad[0] + 3], got[bad[0]], ref[bad[0]])


def test_chunk_boundaries_do_not_change_the_result(tmp_path):
    """A file large enough to be cut into many chunks, relative indices reaching across chunk boundaries."""
    rng = np.random.default_rng(11)
    lines, made = [], 0
    for i in range(40000):
        for _ in range(int(rng.integers(0, 3))):
            x, y, z = rng.normal(size=3)
            lines.append(f"v {x:.9g} {y:.9g} {z:.9g}"); made += 1
        if made >= 3:
            k = int(rng.integers(3, 6))
            rel = -rng.integers(1, min(made, 5000) + 1, k)
            lines.append("f " + " ".join(str(int(v)) for v in rel))
    path = tmp_path / "big.obj"
    path.write_text("\n".join(lines) + "\n")
    base = parse_obj(path, 1)
    for threads in (2, 7, 16):
        got = parse_obj(path, threads)
        assert np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1])
    want = obj_oracle.load_model(path)
    assert tris_from(*base).tobytes() == want.tobytes()


def test_missing_file_and_overlong_line(tmp_path):
    with pytest.raises(HagridError):
        parse_obj(tmp_path / "nope.obj")
    path = tmp_path / "long.obj"
    path.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\n# " + "x" * 2000 + "\nf 1 2 3\n")
    with pytest.raises(HagridError):        # the reference's getline gives up here (src/load_obj.cpp:103); we say so
        parse_obj(path)


# ----------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if bool(Z[f"{n}_ok"])])
def test_device_ingest_matches_the_reference_loader(lib, tmp_path, name):
    sc = Scene(write(tmp_path, name), lib=lib)
    assert sc.download_tris().tobytes() == Z[f"{name}_tris"].tobytes()
    sc.close()


@pytest.mark.gpu
def test_ingest_of_a_large_scene_and_the_grid_built_from_it(lib, ref_lib, tmp_path):
    """300 K triangles through both loaders: identical triangle arrays, hence identical grids."""
    tris = scenes.atrium(300000, seed=3)
    path = tmp_path / "atrium.obj"
    scenes.write_obj(path, tris)
    a, b = Scene(path, lib=ref_lib), Scene(path, lib=lib, threads=8)
    ta, tb = a.download_tris(), b.download_tris()
    assert ta.shape[0] == 300000 and ta.tobytes() == tb.tobytes()
    a.build_all(0.15, 3.0); b.build_all(0.15, 3.0)
    ia, ib = a.info().as_dict(), b.info().as_dict()
    assert ia == ib
    a.close(); b.close()
