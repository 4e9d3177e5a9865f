"""On-disk formats: the reference's .rays files (load_rays, src/main.cpp:277-300) and
this library's grid cache."""
import numpy as np
import pytest

from hagrid_b200 import HIT_PRIM_ID, RAY_DTYPE, HagridError, Library, Scene, scenes
from util import Golden, grid_diff


def test_rays_file_count_follows_the_reference(tmp_path):
    """count = file size / 24, trailing bytes ignored (src/main.cpp:282); no device needed."""
    lib = Library()
    path = tmp_path / "r.rays"
    path.write_bytes(b"\0" * (24 * 7 + 5))
    assert lib.dll.hgb_rays_file_count(str(path).encode()) == 7
    assert lib.dll.hgb_rays_file_count(str(tmp_path / "missing.rays").encode()) == -1


@pytest.mark.gpu
def test_rays_files_round_trip_and_match_the_reference_loader(lib, ref_lib, tmp_path):
    g = Golden("soup800")
    rays = np.concatenate([g.rays, scenes.random_rays(g.tris, 300001, seed=5)])
    path = tmp_path / "frame.rays"
    scenes.write_rays(path, rays)                         # the writer the test-suite always used: org, dir as float32
    a, b = Scene(g.tris, lib=ref_lib), Scene(g.tris, lib=lib)
    want = a.load_rays(path, 0.25, 77.0)                 # the reference's own load_rays, uploaded
    got = b.load_rays(path, 0.25, 77.0)                  # 24-byte records uploaded, expanded on the device
    assert want.shape[0] == rays.shape[0] and got.tobytes() == want.tobytes()
    assert np.array_equal(got["org"], rays["org"]) and np.array_equal(got["dir"], rays["dir"])
    assert (got["tmin"] == np.float32(0.25)).all() and (got["tmax"] == np.float32(77.0)).all()
    out = tmp_path / "copy.rays"
    b.save_rays(out, rays)
    assert out.read_bytes() == path.read_bytes()
    empty = tmp_path / "empty.rays"
    empty.write_bytes(b"")
    assert b.load_rays(empty).shape[0] == 0
    with pytest.raises(HagridError):
        b.load_rays(tmp_path / "missing.rays")
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("compress", [False, True])
def test_grid_cache_round_trip(lib, tmp_path, compress):
    g = Golden("strands1500")
    a = Scene(g.tris, lib=lib)
    a.build_all(g.top_density, g.snd_density, g.alpha, g.expansion, compress)
    path = tmp_path / "scene.hgrid"
    a.save_grid(path)
    b = Scene(g.tris, lib=lib)
    b.load_grid(path)
    ia, ea, ca, ra = a.download(); ib, eb, cb, rb = b.download()
    assert grid_diff(ib.as_dict(), (eb, cb, rb), (ia.as_dict(), ea, ca, ra)) == []
    b.setup_traversal()
    key = "hits_small_ids" if compress else "hits_cell_ids"
    got = b.trace(g.rays, HIT_PRIM_ID)
    assert np.array_equal(got["id"], g.hits[key]["id"]) and np.array_equal(got["t"].view(np.uint32), g.hits[key]["t"].view(np.uint32))
    b.build_all(g.top_density, g.snd_density)             # the loaded arrays belong to the scene's pool: a rebuild frees them
    # damaged files are refused, the scene keeps its grid
    blob = path.read_bytes()
    header = 104 + 4 * len(ia.as_dict()["offsets"])
    damaged = [blob[:-3], b"XGRID001" + blob[8:], blob + b"\0"]
    # right size, wrong contents: a voxel-map word that names a cell that does not exist, an inner node that points at
    # itself (an endless look-up), a cell whose reference range leaves the array
    bad_leaf = bytearray(blob); bad_leaf[header:header + 4] = np.uint32(0x7FFFFFF << 2).tobytes()
    self_loop = bytearray(blob); self_loop[header:header + 4] = np.uint32((0 << 2) | 1).tobytes()
    cells_at = header + 4 * ia.num_entries
    bad_cell = bytearray(blob); bad_cell[cells_at + 12:cells_at + 16] = np.int32(2**30).tobytes()
    damaged += [bytes(bad_leaf), bytes(self_loop), bytes(bad_cell)]
    for bad in damaged:
        (tmp_path / "bad.hgrid").write_bytes(bad)
        with pytest.raises(HagridError):
            b.load_grid(tmp_path / "bad.hgrid")
    assert b.info().num_cells > 0
    a.close(); b.close()
