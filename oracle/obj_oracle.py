"""Sequential restatement of the reference's scene ingest — TEST INFRASTRUCTURE (see oracle/hagrid_oracle.c).
  load_obj   src/load_obj.cpp:78-239   (only what reaches the tracer: positions and faces)
  load_model src/main.cpp:246-275      (fan triangulation, e1 = v0 - v1, e2 = v2 - v0, n = e1 x e2)
Pure-Python loops: for small files only. Pinned by tests/golden/ingest.npz, which the reference's own load_model
produced (oracle/ref_frontend.cpp includes src/main.cpp; tests/golden/make_ingest_golden.py)."""
from __future__ import annotations

import ctypes
import re

import numpy as np

TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("nx", "<f4"), ("e1", "<f4", 3), ("ny", "<f4"), ("e2", "<f4", 3), ("nz", "<f4")])
_libc = ctypes.CDLL(None)
_libc.strtof.restype = ctypes.c_float
_libc.strtof.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p)]
_INT = re.compile(rb"[ \t\n\v\f\r]*[+-]?[0-9]+")
_SPACE = b" \t\n\v\f\r"


def _strtof3(text: bytes):
    """Three consecutive strtof calls like src/load_obj.cpp:119-122 (C locale)."""
    out = []
    buf = ctypes.create_string_buffer(text + b"\0")
    addr = ctypes.addressof(buf)
    pos = 0
    for _ in range(3):
        end = ctypes.c_char_p()
        val = _libc.strtof(ctypes.c_char_p(addr + pos), ctypes.byref(end))
        pos = ctypes.cast(end, ctypes.c_void_p).value - addr
        out.append(val)
    return out


def _strtol(text: bytes, pos: int):
    m = _INT.match(text, pos)
    if not m:
        return 0, pos
    return int(m.group(0)), m.end()


def _skip(text: bytes, pos: int):
    while pos < len(text) and text[pos:pos + 1] in (b" ", b"\t", b"\n", b"\v", b"\f", b"\r"):
        pos += 1
    return pos


def _read_index(text: bytes, pos: int):
    """read_index, src/load_obj.cpp:42-76. Returns (ok, pos, v, t, n)."""
    base = _skip(text, pos)
    c = text[base:base + 1]
    if not (c.isdigit() or c == b"-"):
        return False, pos, 0, 0, 0
    v, base = _strtol(text, base)
    t = n = 0
    base = _skip(text, base)
    if text[base:base + 1] == b"/":
        base += 1
        if text[base:base + 1] != b"/":
            t, base = _strtol(text, base)
        base = _skip(text, base)
        if text[base:base + 1] == b"/":
            base += 1
            n, base = _strtol(text, base)
    return True, base, v, t, n


def load_model(path) -> np.ndarray | None:
    """The Tri array the reference's front end would upload, or None when its loader refuses the file."""
    vertices = [(0.0, 0.0, 0.0)]          # dummy vertex, src/load_obj.cpp:96
    num_normals = num_texcoords = 1
    faces = []
    errors = 0
    data = open(path, "rb").read()
    for raw in data.split(b"\n"):
        if len(raw) >= 1023:              # getline(line, 1024) fails: the loop ends silently (src/load_obj.cpp:103)
            break
        ptr = raw.lstrip(_SPACE)
        if not ptr or ptr[:1] == b"#":
            continue
        ptr = ptr[:1] + ptr[1:].rstrip(_SPACE)      # remove_eol keeps the first character
        if ptr[:1] == b"v":
            kind = ptr[1:2]
            if kind in (b" ", b"\t"):
                vertices.append(tuple(_strtof3(ptr[1:])))
            elif kind == b"n":
                num_normals += 1
            elif kind == b"t":
                num_texcoords += 1
            else:
                errors += 1
        elif ptr[:1] == b"f" and ptr[1:2] in (b" ", b"\t", b"\v", b"\f", b"\r"):
            pos, corners = 2, []
            while len(corners) < 8:
                ok, pos, v, t, n = _read_index(ptr, pos)
                if not ok:
                    break
                corners.append((v, t, n))
            if len(corners) < 3:
                errors += 1
                continue
            fixed = [(len(vertices) + v if v < 0 else v, num_texcoords + t if t < 0 else t, num_normals + n if n < 0 else n)
                     for v, t, n in corners]
            if any(v <= 0 or t < 0 or n < 0 for v, t, n in fixed):
                errors += 1
                continue
            faces.append([v for v, _, _ in fixed])
        elif ptr[:1] in (b"g", b"o", b"s") and ptr[1:2] in (b" ", b"\t", b"\v", b"\f", b"\r"):
            pass
        elif ptr[:6] in (b"usemtl", b"mtllib") and ptr[6:7] in (b" ", b"\t", b"\v", b"\f", b"\r"):
            pass
        else:
            errors += 1
    if errors or any(v >= len(vertices) for f in faces for v in f):
        return None
    V = np.array(vertices, dtype=np.float32)
    idx = np.array([(f[0], f[i + 1], f[i + 2]) for f in faces for i in range(len(f) - 2)], dtype=np.int64).reshape(-1, 3)
    v0, v1, v2 = V[idx[:, 0]], V[idx[:, 1]], V[idx[:, 2]]
    e1, e2 = v0 - v1, v2 - v0
    n = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                  e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], axis=1)
    tris = np.empty(idx.shape[0], dtype=TRI_DTYPE)
    tris["v0"], tris["e1"], tris["e2"] = v0, e1, e2
    tris["nx"], tris["ny"], tris["nz"] = n[:, 0], n[:, 1], n[:, 2]
    return tris
