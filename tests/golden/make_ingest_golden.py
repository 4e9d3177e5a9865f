"""Golden vectors of the reference's scene ingest (ObjLoader::load_obj + load_model, src/load_obj.cpp:78-239,
src/main.cpp:246-275), produced by the reference's OWN code through oracle/ref_frontend.cpp (which includes
src/main.cpp unmodified). Pure CPU work; run where oracle/_ref/libhagrid_ref.so was built, commit ingest.npz."""
import ctypes as C
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
TRI = np.dtype([("v0", "<f4", 3), ("nx", "<f4"), ("e1", "<f4", 3), ("ny", "<f4"), ("e2", "<f4", 3), ("nz", "<f4")])

dll = C.CDLL(str(ROOT / "oracle" / "_ref" / "libhagrid_ref.so"))
dll.hgb_ref_load_model.restype = C.c_void_p
dll.hgb_ref_load_model.argtypes = [C.c_char_p, C.POINTER(C.c_int)]


def reference_tris(text: bytes):
    with tempfile.NamedTemporaryFile(suffix=".obj") as f:
        f.write(text); f.flush()
        n = C.c_int(0)
        ptr = dll.hgb_ref_load_model(f.name.encode(), C.byref(n))
        if not ptr:
            return None
        return np.frombuffer((C.c_char * (48 * n.value)).from_address(ptr), dtype=TRI).copy()


TRICKY = b"""# every syntax the loader accepts
mtllib scene.mtl
o first
g walls
v 0 0 0
v 1.5 0 0\r
v 1.5e0 2.25 0
  v\t0 2.25 -0.125   
vn 0 0 1
vt 0.5 0.5
vt 0.25 0.75
usemtl white
s off
f 1 2 3
f 1/1 3/2 4/1
f 1//1 2//1 3//1 4//1
f -4 -3 -2 -1
f 1/1/1 2/2/1 3/1/1
g second group
v -1 -1 -1
v 1 -1 -1
v 1 1 -1
v -1 1 -1
v -1 -1 1
v 1 -1 1
v 1 1 1
v -1 1 1
f 5 6 7 8 9 10 11 12
f 5 6 7 8 9 10 11 12 1 2
f 9/-1 10/-2 11/-1
o second
f 13 14 15
v 0.1 0.2 0.3
v .4 -.5 +.6
v 7e-1 8E+0 9
f -3 -2 -1
f 2 14 3
"""

cases = {"tricky": TRICKY}
rng = np.random.default_rng(7)
for k, (nv, nf) in enumerate([(50, 120), (400, 900)]):
    lines = []
    verts = rng.normal(size=(nv, 3)).astype(np.float32)
    made = 0
    for i in range(nf):
        while made < nv and (made < 3 or rng.random() < 0.4):
            x, y, z = verts[made]; made += 1
            lines.append(f"v {x:.9g} {y:.9g} {z:.9g}")
            if rng.random() < 0.2: lines.append("vn 0 1 0")
            if rng.random() < 0.2: lines.append("vt 0 1")
        corners = int(rng.integers(3, 7))
        if rng.random() < 0.5:
            idx = rng.integers(1, made + 1, corners)
        else:
            idx = -rng.integers(1, made + 1, corners)
        lines.append("f " + " ".join(str(int(v)) for v in idx))
        if rng.random() < 0.05: lines.append(f"g part{i}")
        if rng.random() < 0.02: lines.append(f"o obj{i}")
    cases[f"random{k}"] = ("\n".join(lines) + "\n").encode()
# files the reference's loader refuses
cases["bad_command"] = b"v 0 0 0\nv 1 0 0\nv 0 1 0\nl 1 2\nf 1 2 3\n"
cases["bad_face"] = b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n"
cases["bad_index"] = b"v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -4\n"

out = {}
for name, text in cases.items():
    tris = reference_tris(text)
    out[f"{name}_obj"] = np.frombuffer(text, dtype=np.uint8)
    out[f"{name}_ok"] = np.array(tris is not None)
    out[f"{name}_tris"] = tris if tris is not None else np.empty(0, dtype=TRI)
    print(name, "refused" if tris is None else f"{tris.shape[0]} triangles")
np.savez_compressed(Path(__file__).with_name("ingest.npz"), **out)
