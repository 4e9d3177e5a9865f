"""Camera frames: gen_camera, gen_rays and update_surface of the reference's
front end (src/main.cpp:42-111).
CPU tier: the oracle's restatement against golden vectors produced by the reference's own code
(tests/golden/frontend.npz, made by oracle/ref_frontend.cpp = src/main.cpp included unmodified).
GPU tier: rays generated on the device and fused frames (generate + trace + colour in one launch) against the
oracle and against the reference's host loop driven through the same C ABI."""
from pathlib import Path

import numpy as np
import pytest

from hagrid_b200 import HIT_STEPS, Scene, make_camera, scenes
from oracle import oracle
from util import Golden

Z = np.load(Path(__file__).resolve().parent / "golden" / "frontend.npz")
CASES = Z["cases"]


def case(k):
    c = CASES[k]
    return c[0:3], c[3:6], c[6:9], float(c[9]), int(c[10]), int(c[11]), float(c[12])


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_camera_and_rays_match_the_reference_front_end(k):
    eye, center, up, fov, w, h, clip = case(k)
    cam = oracle.gen_camera(eye, center, up, fov, w / h)
    assert cam.tobytes() == Z[f"cam{k}"].tobytes()
    rays = oracle.gen_rays(cam, clip, w, h)
    assert rays.tobytes() == Z[f"rays{k}"].tobytes()


@pytest.mark.parametrize("k", range(len(CASES)))
@pytest.mark.parametrize("mode", (0, 1, 2))
def test_oracle_pixels_match_the_reference_front_end(k, mode):
    _, _, _, _, w, h, clip = case(k)
    img = oracle.update_surface(mode, Z[f"hits{k}"], clip, w, h)
    assert np.array_equal(img, Z[f"image{k}_{mode}"])


def test_host_camera_of_the_library_matches_the_reference(tmp_path):
    """hgb_make_camera is host arithmetic: it runs without a GPU."""
    for k in range(len(CASES)):
        eye, center, up, fov, w, h, _ = case(k)
        assert make_camera(eye, center, up, fov, w / h).tobytes() == Z[f"cam{k}"].tobytes()


# ----------------------------------------------------------------------------- GPU tier
@pytest.fixture(scope="module")
def atrium(lib):
    tris = scenes.atrium(60000, seed=5)
    sc = Scene(tris, lib=lib)
    sc.build_all(0.15, 3.0)
    yield tris, sc
    sc.close()


def view(tris, w, h, yaw=0.35):
    lo, hi = scenes.scene_bbox(tris)
    eye = 0.5 * (lo + hi)
    target = eye + np.array([np.sin(yaw), -0.05, np.cos(yaw)], np.float32)
    clip = float(np.linalg.norm(hi - lo))
    return oracle.gen_camera(eye, target, (0, 1, 0), 60.0, w / h), clip


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(64, 48), (1920, 1080), (101, 37)])
def test_device_ray_generation_is_bit_identical(lib, atrium, size):
    tris, sc = atrium
    w, h = size
    cam, clip = view(tris, w, h)
    got = sc.generate_rays(cam, clip, w, h)
    assert got.tobytes() == oracle.gen_rays(cam, clip, w, h).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("compressed", [False, True])
@pytest.mark.parametrize("size", [(640, 360), (100, 37)])      # tiled 8x4 and scan-line pixel order
def test_fused_frame_equals_generate_trace_colour(lib, atrium, size, compressed):
    tris, sc0 = atrium
    w, h = size
    sc = sc0
    if compressed:
        sc = Scene(tris, lib=lib)
        sc.build_all(0.15, 3.0, compress=True)
    sc.setup_traversal()
    cam, clip = view(tris, w, h)
    steps = sc.trace(oracle.gen_rays(cam, clip, w, h), HIT_STEPS)         # Hit.id = step count, src/traverse.cu:93
    for mode in (0, 1, 2):
        want = oracle.update_surface(mode, steps, clip, w, h)
        got = sc.render_frame(cam, clip, w, h, mode)
        assert np.array_equal(got, want), mode
    if compressed:
        sc.close()


@pytest.mark.gpu
def test_fused_frame_equals_the_reference_viewer_loop(lib, ref_lib):
    """The reference build of the same entry point runs src/main.cpp's own gen_rays / update_surface around its
    own traverse_grid: the images must be identical byte for byte."""
    g = Golden("strands1500")
    a, b = Scene(g.tris, lib=ref_lib), Scene(g.tris, lib=lib)
    for sc in (a, b):
        sc.build_all(g.top_density, g.snd_density, g.alpha, g.expansion)
    lo, hi = scenes.scene_bbox(g.tris)
    cam = make_camera(lo - (hi - lo) * 0.7, 0.5 * (lo + hi), (0, 1, 0), 55.0, 320 / 200, lib=ref_lib)
    assert cam.tobytes() == make_camera(lo - (hi - lo) * 0.7, 0.5 * (lo + hi), (0, 1, 0), 55.0, 320 / 200, lib=lib).tobytes()
    clip = float(np.linalg.norm(hi - lo)) * 2
    for mode in (0, 1, 2):
        a.setup_traversal()
        want = a.render_frame(cam, clip, 320, 200, mode)
        b.setup_traversal()
        got = b.render_frame(cam, clip, 320, 200, mode)
        assert np.array_equal(got, want), mode
        assert got[..., 3].min() == 255 and (mode == 0 or got[..., :3].max() > 0)
    a.close(); b.close()


@pytest.mark.gpu
def test_dynamic_scene_rebuild_and_render_matches_the_reference(lib, ref_lib):
    """Row f4: every frame new triangles, rebuild with the keep-alive allocator (mem_manager.h:36-42), render.
    Same images as the reference driven through the same loop; the pool stops growing after the first lap."""
    base = scenes.atrium(30000, seed=9)
    poses = [scenes.animate(base, 2 * np.pi * f / 4) for f in range(4)]
    w, h = 320, 184
    images = {}
    for name, L in (("ref", ref_lib), ("new", lib)):
        sc = Scene(poses[0], keep_alive=True, lib=L)
        peaks = []
        for lap in range(3):
            for k, tris in enumerate(poses):
                sc.set_tris(tris)
                sc.build_all(0.15, 3.0)
                sc.setup_traversal()
                cam, clip = view(tris, w, h, yaw=0.2 * k)
                img = sc.render_frame(cam, clip, w, h, 2)
                if lap == 0:
                    images.setdefault(name, []).append(img)
                else:
                    assert np.array_equal(img, images[name][k])
            peaks.append(sc.peak_bytes())
        if name == "new":
            assert peaks[2] == peaks[1]
        sc.close()
    for k in range(4):
        assert np.array_equal(images["new"][k], images["ref"][k]), k
    assert not np.array_equal(images["new"][0], images["new"][2])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["depth", "heat"])
def test_headless_viewer_writes_the_reference_start_up_frame(lib, tmp_path, mode):
    """python -m hagrid_b200.view: OBJ in, PPM out; the first frame is the reference's start-up view
    (eye at the box centre, forward +z, src/main.cpp:572-588) coloured like update_surface."""
    from hagrid_b200 import view as viewer
    tris = scenes.atrium(4000, seed=3)
    obj, ppm = tmp_path / "scene.obj", tmp_path / "frame.ppm"
    scenes.write_obj(obj, tris)
    w, h = 128, 96
    assert viewer.main([str(obj), "-o", str(ppm), "-sx", str(w), "-sy", str(h), "-td", "0.15", "-sd", "3.0", "--mode", mode]) == 0
    raw = ppm.read_bytes()
    head = f"P6\n{w} {h}\n255\n".encode()
    assert raw.startswith(head)
    got = np.frombuffer(raw[len(head):], np.uint8).reshape(h, w, 3)

    sc = Scene(obj, lib=lib)                       # same loader, so the same triangles and scene box
    sc.build_all(0.15, 3.0)
    sc.setup_traversal()
    gi = sc.info()
    lo, hi = np.array(gi.bbox_min, np.float32), np.array(gi.bbox_max, np.float32)
    eye = (lo + hi) * np.float32(0.5)
    clip = float(np.sqrt(np.sum((hi - lo) * (hi - lo), dtype=np.float32)))
    cam = oracle.gen_camera(eye, eye + np.array([0, 0, 100], np.float32), (0, 1, 0), 60.0, w / h)
    hits = sc.trace(oracle.gen_rays(cam, clip, w, h), HIT_STEPS)
    want = oracle.update_surface({"depth": 0, "heat": 2}[mode], hits, clip, w, h)
    assert np.array_equal(got, want[..., 2::-1])
    sc.close()
    # several frames: numbered files, the camera turns
    assert viewer.main([str(obj), "-o", str(ppm), "-sx", "64", "-sy", "48", "--frames", "3"]) == 0
    frames = [(tmp_path / f"frame_{k:04d}.ppm").read_bytes() for k in range(3)]
    assert len({len(f) for f in frames}) == 1 and frames[0] != frames[1]
