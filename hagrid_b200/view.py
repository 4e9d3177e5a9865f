"""Headless stand-in for the reference's interactive mode (src/main.cpp:550-628): loads an OBJ scene, builds
the grid with the reference's options and writes frames as PPM files instead of drawing into an SDL window.

    python -m hagrid_b200.view scene.obj -o frame.ppm [-sx 1024 -sy 1024 -f 60 -c CLIP] [-td 0.12 -sd 2.4 -a 0.995 -e 3 -z]
                               [--mode depth|steps|heat] [--eye x y z --forward x y z --up x y z] [--frames N]

Option names and defaults are the reference's (src/main.cpp:121-140, 181-230). The first frame is the one the
reference shows on start-up: eye at the centre of the scene box, looking down +z, depth display (src/main.cpp:572-588).
With --frames N the camera turns about the up axis, one full turn over N frames, and the files are numbered.
Every frame is one fused launch on the device (generate, trace, colour): Scene.render_frame."""
from __future__ import annotations

import argparse
import sys
import time
from pathlib import Path

import numpy as np

from .api import Scene, library, make_camera, save_image

MODES = {"depth": 0, "steps": 1, "heat": 2}          # DisplayMode, src/main.cpp:34-38


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m hagrid_b200.view", description=__doc__.split("\n\n")[0])
    ap.add_argument("scene", help="Wavefront OBJ file")
    ap.add_argument("-o", "--output", default="frame.ppm")
    ap.add_argument("-sx", "--width", type=int, default=1024)
    ap.add_argument("-sy", "--height", type=int, default=1024)
    ap.add_argument("-c", "--clip", type=float, default=0.0, help="<= 0: length of the scene box diagonal")
    ap.add_argument("-f", "--fov", type=float, default=60.0)
    ap.add_argument("-td", "--top-density", type=float, default=0.12)
    ap.add_argument("-sd", "--snd-density", type=float, default=2.4)
    ap.add_argument("-a", "--alpha", type=float, default=0.995)
    ap.add_argument("-e", "--expansion", type=int, default=3)
    ap.add_argument("-z", "--compress", action="store_true")
    ap.add_argument("--mode", choices=sorted(MODES), default="depth")
    ap.add_argument("--eye", type=float, nargs=3)
    ap.add_argument("--forward", type=float, nargs=3, default=(0.0, 0.0, 1.0))
    ap.add_argument("--up", type=float, nargs=3, default=(0.0, 1.0, 0.0))
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    lib = library()
    t0 = time.perf_counter()
    scene = Scene(Path(args.scene), device=args.device, keep_alive=True, lib=lib)
    t1 = time.perf_counter()
    ms = scene.build_all(args.top_density, args.snd_density, args.alpha, args.expansion, compress=args.compress)
    scene.setup_traversal()
    gi = scene.info()
    lo, hi = np.array(gi.bbox_min, np.float32), np.array(gi.bbox_max, np.float32)
    clip = args.clip if args.clip > 0 else float(np.sqrt(np.sum((hi - lo) * (hi - lo), dtype=np.float32)))
    eye = np.array(args.eye, np.float32) if args.eye else (lo + hi) * np.float32(0.5)
    print(f"{scene.num_tris} triangles loaded in {1e3 * (t1 - t0):.1f} ms, grid built in {float(ms[0]):.2f} ms: "
          f"{gi.num_cells} cells, {gi.num_refs} references, {gi.num_entries} voxel map entries")

    forward, up = np.array(args.forward, np.float64), np.array(args.up, np.float64)
    up /= np.linalg.norm(up)
    out = Path(args.output)
    image = np.empty((args.height, args.width, 4), np.uint8)
    spent = 0.0
    for k in range(max(1, args.frames)):
        angle = 2 * np.pi * k / max(1, args.frames)
        # Rodrigues rotation of the view direction about the up axis
        f = forward * np.cos(angle) + np.cross(up, forward) * np.sin(angle) + up * np.dot(up, forward) * (1 - np.cos(angle))
        cam = make_camera(eye, eye + (f * 100.0).astype(np.float32), up.astype(np.float32), args.fov, args.width / args.height, lib=lib)
        t = time.perf_counter()
        scene.render_frame(cam, clip, args.width, args.height, MODES[args.mode], image)
        spent += time.perf_counter() - t
        path = out if args.frames <= 1 else out.with_name(f"{out.stem}_{k:04d}{out.suffix}")
        save_image(path, image, lib=lib)
    print(f"{max(1, args.frames)} frame(s) of {args.width}x{args.height}, {1e3 * spent / max(1, args.frames):.3f} ms per frame "
          f"(generate + trace + colour + download), written to {out if args.frames <= 1 else out.with_name(out.stem + '_NNNN' + out.suffix)}")
    scene.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
