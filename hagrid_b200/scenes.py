"""Seeded procedural stand-ins for the benchmark scenes and ray buffers.

No scene assets exist offline, so the configurations of
BASELINE.json are reproduced with synthetic geometry of the same triangle
counts and the same "teapot in a stadium" density variation, and with ray
buffers built by the reference's own formulas (gen_camera / gen_rays,
src/main.cpp:42-66). Triangles are set up exactly like load_model does
(src/main.cpp:255-268): e1 = v0 - v1, e2 = v2 - v0, n = cross(e1, e2), all in
float32.  Everything is numpy on the host: this is input generation, not part of
the measured path.
"""
from __future__ import annotations

import numpy as np

from .api import RAY_DTYPE, TRI_DTYPE

SEED = 0x48414752  # "HAGR"


# ----------------------------------------------------------------------------- triangles
def make_tris(v0, v1, v2) -> np.ndarray:
    v0 = np.asarray(v0, dtype=np.float32).reshape(-1, 3)
    v1 = np.asarray(v1, dtype=np.float32).reshape(-1, 3)
    v2 = np.asarray(v2, dtype=np.float32).reshape(-1, 3)
    e1 = v0 - v1
    e2 = v2 - v0
    n = np.empty_like(e1)
    n[:, 0] = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    n[:, 1] = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    n[:, 2] = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    t = np.empty(v0.shape[0], dtype=TRI_DTYPE)
    t["v0"], t["e1"], t["e2"] = v0, e1, e2
    t["nx"], t["ny"], t["nz"] = n[:, 0], n[:, 1], n[:, 2]
    return t


def tri_vertices(tris: np.ndarray):
    v0 = tris["v0"]
    return v0, v0 - tris["e1"], v0 + tris["e2"]


def _grid_tris(P: np.ndarray) -> np.ndarray:
    """(nu, nv, 3) lattice of points -> (2*(nu-1)*(nv-1), 3, 3) triangle vertices."""
    a, b, c, d = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
    t1 = np.stack([a, b, c], axis=-2).reshape(-1, 3, 3)
    t2 = np.stack([a, c, d], axis=-2).reshape(-1, 3, 3)
    return np.concatenate([t1, t2], axis=0)


def _patch(origin, du, dv, nu, nv, bump=None):
    u = np.linspace(0.0, 1.0, nu + 1)[:, None, None]
    v = np.linspace(0.0, 1.0, nv + 1)[None, :, None]
    P = np.asarray(origin, float) + u * np.asarray(du, float) + v * np.asarray(dv, float)
    if bump is not None:
        P = P + bump(u, v)
    return _grid_tris(P)


def _box(lo, hi, n=1):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    return np.concatenate([
        _patch(lo, ex, ey, n, n), _patch(lo + ez, ex, ey, n, n),
        _patch(lo, ex, ez, n, n), _patch(lo + ey, ex, ez, n, n),
        _patch(lo, ey, ez, n, n), _patch(lo + ex, ey, ez, n, n)])


def _cylinder(center, radius, height, nseg, nring, flute=0.0):
    a = np.linspace(0.0, 2 * np.pi, nseg + 1)[:, None]
    h = np.linspace(0.0, 1.0, nring + 1)[None, :]
    r = radius * (1.0 + flute * np.cos(12 * a)) * (1.0 - 0.08 * np.sin(np.pi * h))
    P = np.stack([center[0] + r * np.cos(a), center[1] + height * h + 0 * a, center[2] + r * np.sin(a)], axis=-1)
    return _grid_tris(P)


def _sphere(center, radius, nseg, nring):
    a = np.linspace(0.0, 2 * np.pi, nseg + 1)[:, None]
    b = np.linspace(0.0, np.pi, nring + 1)[None, :]
    P = np.stack([center[0] + radius * np.cos(a) * np.sin(b), center[1] + radius * np.cos(b) + 0 * a,
                  center[2] + radius * np.sin(a) * np.sin(b)], axis=-1)
    return _grid_tris(P)


def _arch(p0, p1, y, rise, thick, nseg, depth):
    """Half-torus arch between two column tops, extruded by `depth` along its normal."""
    p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
    s = np.linspace(0.0, np.pi, nseg + 1)[:, None]
    w = np.linspace(0.0, 2 * np.pi, 9)[None, :]
    mid = 0.5 * (p0 + p1)
    half = 0.5 * (p1 - p0)
    span = np.linalg.norm(half)
    axis = half / span
    side = np.array([-axis[2], 0.0, axis[0]])
    cx = mid[0] - np.cos(s) * half[0]
    cz = mid[2] - np.cos(s) * half[2]
    cy = y + rise * np.sin(s)
    rad = thick * (1.0 + 0 * s)
    P = np.stack([cx + rad * np.cos(w) * side[0] * depth, cy + rad * np.sin(w), cz + rad * np.cos(w) * side[2] * depth],
                 axis=-1)
    return _grid_tris(P)


def _to_tris(parts, count=None, rng=None, filler_box=None) -> np.ndarray:
    V = np.concatenate(parts, axis=0).astype(np.float32)
    if count is not None:
        if V.shape[0] > count:
            V = V[:count]
        elif V.shape[0] < count:
            # pad with tiny scattered "debris" triangles inside filler_box
            k = count - V.shape[0]
            lo, hi = filler_box
            c = rng.uniform(lo, hi, size=(k, 3))
            s = 0.002 * float(np.max(np.asarray(hi) - np.asarray(lo)))
            tri = c[:, None, :] + rng.normal(0.0, s, size=(k, 3, 3))
            V = np.concatenate([V, tri.astype(np.float32)], axis=0)
    return make_tris(V[:, 0], V[:, 1], V[:, 2])


def cornell32() -> np.ndarray:
    """C1: classic 555-unit Cornell box: 5 walls (10) + light (2) + two blocks (10 + 10) = 32 tris."""
    def quad(a, b, c, d):
        return np.array([[a, b, c], [a, c, d]], dtype=np.float64)
    S = 555.0
    parts = [
        quad([0, 0, 0], [S, 0, 0], [S, 0, S], [0, 0, S]),          # floor
        quad([0, S, 0], [0, S, S], [S, S, S], [S, S, 0]),          # ceiling
        quad([0, 0, S], [S, 0, S], [S, S, S], [0, S, S]),          # back
        quad([0, 0, 0], [0, 0, S], [0, S, S], [0, S, 0]),          # right
        quad([S, 0, 0], [S, S, 0], [S, S, S], [S, 0, S]),          # left
        quad([213, S - 1, 227], [343, S - 1, 227], [343, S - 1, 332], [213, S - 1, 332]),  # light
    ]

    def block(cx, cz, w, h, ang):
        c, s = np.cos(ang), np.sin(ang)
        base = [np.array([cx + c * x - s * z, 0.0, cz + s * x + c * z]) for x, z in
                ((-w, -w), (w, -w), (w, w), (-w, w))]
        top = [b + np.array([0, h, 0]) for b in base]
        q = [quad(top[0], top[1], top[2], top[3])]
        for i in range(4):
            j = (i + 1) % 4
            q.append(quad(base[i], base[j], top[j], top[i]))
        return np.concatenate(q)
    parts.append(block(185, 169, 82, 165, 0.29))
    parts.append(block(368, 351, 82, 330, -0.31))
    t = _to_tris(parts)
    assert t.shape[0] == 32
    return t


def atrium(num_tris: int, seed: int = SEED, foliage: float = 0.0) -> np.ndarray:
    """Sponza-like atrium: coarse floor/walls, two storeys of fluted columns and
    arches, finely tessellated hanging drapes, ornament spheres; optional
    instanced foliage clusters. Trimmed/padded to exactly `num_tris`."""
    rng = np.random.default_rng(seed)
    L, W, H = 60.0, 26.0, 24.0
    scale = (num_tris / 262267.0) ** 0.5
    parts = []
    # coarse shell (big triangles: the "stadium")
    parts.append(_patch([-L / 2, 0, -W / 2], [L, 0, 0], [0, 0, W], 12, 6))
    parts.append(_patch([-L / 2, H, -W / 2], [L, 0, 0], [0, 0, W], 2, 2))
    parts.append(_patch([-L / 2, 0, -W / 2], [L, 0, 0], [0, H, 0], 6, 3))
    parts.append(_patch([-L / 2, 0, W / 2], [L, 0, 0], [0, H, 0], 6, 3))
    parts.append(_patch([-L / 2, 0, -W / 2], [0, 0, W], [0, H, 0], 3, 3))
    parts.append(_patch([L / 2, 0, -W / 2], [0, 0, W], [0, H, 0], 3, 3))
    # galleries (upper floors along both long sides)
    for side in (-1, 1):
        z0 = side * (W / 2 - 5.0) - (2.5 if side > 0 else -2.5) - 2.5
        parts.append(_box([-L / 2, 8.0, min(z0, z0 + 5)], [L / 2, 8.6, max(z0, z0 + 5)], 4))
        parts.append(_box([-L / 2, 16.0, min(z0, z0 + 5)], [L / 2, 16.6, max(z0, z0 + 5)], 4))
    # columns + arches, two storeys
    ncol = 12
    seg = max(8, int(28 * scale))
    ring = max(4, int(36 * scale))
    xs = np.linspace(-L / 2 + 4, L / 2 - 4, ncol)
    for storey, (y0, hgt, rad) in enumerate(((0.0, 8.0, 0.55), (8.6, 7.4, 0.42))):
        for side in (-1, 1):
            z = side * (W / 2 - 5.0)
            for i, x in enumerate(xs):
                parts.append(_cylinder((x, y0, z), rad, hgt * 0.8, seg, ring, flute=0.06))
                parts.append(_box([x - rad * 1.4, y0, z - rad * 1.4], [x + rad * 1.4, y0 + 0.4, z + rad * 1.4], 2))
                if i + 1 < ncol:
                    parts.append(_arch((x, 0, z), (xs[i + 1], 0, z), y0 + hgt * 0.8, hgt * 0.2, 0.25,
                                       max(6, int(24 * scale)), 1.0))
    # drapes: very fine wavy sheets hanging across the nave
    nd = 6
    du = max(16, int(150 * scale))
    dv = max(16, int(110 * scale))
    for k in range(nd):
        x = -L / 2 + (k + 1) * L / (nd + 1)
        ph = rng.uniform(0, 6.28)
        parts.append(_patch([x, 10.0, -W / 2 + 6], [0.0, 0, W - 12], [0.0, 9.0, 0.0], du, dv,
                            bump=lambda u, v, ph=ph: np.concatenate(
                                [0.5 * np.sin(14 * u + ph) * (1 - v) + 0.15 * np.sin(40 * v + 3 * u),
                                 -1.2 * np.sin(np.pi * u) * (1 - v) + 0 * v,
                                 0 * u + 0 * v], axis=-1)))
    # ornaments: dense little spheres (the "teapots")
    nsph = 40
    sseg = max(8, int(36 * scale))
    for k in range(nsph):
        c = (rng.uniform(-L / 2 + 2, L / 2 - 2), rng.uniform(0.4, 7.0), rng.uniform(-W / 2 + 7, W / 2 - 7))
        parts.append(_sphere(c, rng.uniform(0.15, 0.5), sseg, sseg // 2))
    # foliage: clusters of small random leaves
    if foliage > 0:
        nleaf = int(num_tris * foliage)
        ncl = max(1, nleaf // 4000)
        centers = np.stack([rng.uniform(-L / 2 + 2, L / 2 - 2, ncl), rng.uniform(0.5, 14.0, ncl),
                            rng.uniform(-W / 2 + 6.5, W / 2 - 6.5, ncl)], axis=-1)
        which = rng.integers(0, ncl, nleaf)
        c = centers[which] + rng.normal(0, 0.9, size=(nleaf, 3))
        leaf = c[:, None, :] + rng.normal(0, 0.06, size=(nleaf, 3, 3))
        parts.append(leaf)
    have = sum(p.shape[0] for p in parts)
    if have > num_tris:
        # keep the coarse shell and thin the rest uniformly so every object survives
        V = np.concatenate(parts, axis=0)
        shell = 12 * 6 * 2 + 8 + 2 * 36 + 2 * 18
        keep = np.concatenate([np.arange(shell), shell + np.sort(rng.choice(have - shell, num_tris - shell, replace=False))])
        parts = [V[keep]]
    return _to_tris(parts, num_tris, rng, ([-L / 2 + 1, 0.1, -W / 2 + 1], [L / 2 - 1, 1.0, W / 2 - 1]))


def sponza262k() -> np.ndarray:
    """C2/C3 stand-in: 262 267 triangles (Crytek Sponza's count)."""
    return atrium(262267, SEED)


def sanmiguel7p8m() -> np.ndarray:
    """C5 stand-in: 7.8 M triangles, architecture + foliage."""
    return atrium(7_800_000, SEED + 5, foliage=0.45)


def hairball(num_tris: int = 2_000_000, seed: int = SEED + 4) -> np.ndarray:
    """C4 stand-in: curved strands of thin quads inside a unit ball."""
    rng = np.random.default_rng(seed)
    seg = 50
    nstrand = num_tris // (2 * seg)
    d = rng.normal(size=(nstrand, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    root = d * 0.15
    curl = rng.normal(0, 0.6, size=(nstrand, 3))
    s = np.linspace(0.0, 1.0, seg + 1)[None, :, None]
    P = root[:, None, :] + d[:, None, :] * 0.85 * s + curl[:, None, :] * (0.25 * s * s) * np.sin(3.0 * s + curl[:, None, :1])
    tang = np.gradient(P, axis=1)
    side = np.cross(tang, rng.normal(size=(nstrand, 1, 3)))
    side /= np.linalg.norm(side, axis=2, keepdims=True) + 1e-12
    width = 0.0015
    A, B = P - side * width, P + side * width
    a, b, c, dd = A[:, :-1], B[:, :-1], B[:, 1:], A[:, 1:]
    V = np.concatenate([np.stack([a, b, c], axis=2).reshape(-1, 3, 3), np.stack([a, c, dd], axis=2).reshape(-1, 3, 3)])
    tris = _to_tris([V], num_tris, rng, ([-0.5, -0.5, -0.5], [0.5, 0.5, 0.5]))
    return tris


def animate(tris: np.ndarray, phase: float, amplitude: float = 0.01) -> np.ndarray:
    """Dynamic-scene stand-in: every vertex sways by `amplitude` x scene diagonal (a travelling wave in x and z,
    shared by coincident vertices so the mesh stays watertight); triangle count and order are kept."""
    lo, hi = scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    k = 2 * np.pi * 3 / diag
    out = []
    for v in tri_vertices(tris):
        v = v.astype(np.float64)
        w = v.copy()
        w[:, 0] += amplitude * diag * np.sin(k * v[:, 1] + phase)
        w[:, 2] += amplitude * diag * np.cos(k * v[:, 0] + 0.7 * phase)
        out.append(w.astype(np.float32))
    return make_tris(*out)


def small_mixed(num_tris: int = 3000, seed: int = 7) -> np.ndarray:
    """Small random soup with a few huge and many tiny triangles (unit tests)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, size=(num_tris, 3))
    size = np.where(rng.uniform(size=num_tris) < 0.02, 0.8, 0.03)[:, None, None]
    V = c[:, None, :] + rng.normal(0, 1, size=(num_tris, 3, 3)) * size
    return make_tris(V[:, 0], V[:, 1], V[:, 2])


# ----------------------------------------------------------------------------- rays
def scene_bbox(tris: np.ndarray):
    v0, v1, v2 = tri_vertices(tris)
    lo = np.minimum(np.minimum(v0.min(0), v1.min(0)), v2.min(0))
    hi = np.maximum(np.maximum(v0.max(0), v1.max(0)), v2.max(0))
    return lo.astype(np.float32), hi.astype(np.float32)


def _normalize(v):
    v = np.asarray(v, dtype=np.float32)
    return v / np.float32(np.sqrt(np.dot(v, v)))


def primary_rays(eye, center, up, fov, width, height, clip) -> np.ndarray:
    """gen_camera + gen_rays of src/main.cpp:42-66 in float32 (unnormalised directions)."""
    eye = np.asarray(eye, np.float32)
    f = np.float32(np.tan(np.pi * fov / 360.0))
    d = _normalize(np.asarray(center, np.float32) - eye)
    right = _normalize(np.cross(d, np.asarray(up, np.float32))) * np.float32(f * np.float32(width / height))
    upv = _normalize(np.cross(right, d)) * f
    x = np.arange(width, dtype=np.int64)
    y = np.arange(height, dtype=np.int64)
    kx = ((2 * x).astype(np.float32) / np.float32(width) - np.float32(1))[None, :, None]
    ky = (np.float32(1) - (2 * y).astype(np.float32) / np.float32(height))[:, None, None]
    dirs = (d[None, None, :] + right[None, None, :] * kx) + upv[None, None, :] * ky
    rays = np.empty(width * height, dtype=RAY_DTYPE)
    rays["org"] = eye
    rays["dir"] = dirs.reshape(-1, 3).astype(np.float32)
    rays["tmin"] = 0.0
    rays["tmax"] = np.float32(clip)
    return rays


def default_view(tris: np.ndarray, width=1920, height=1080, along_long_axis=False) -> np.ndarray:
    """The reference's initial interactive view (src/main.cpp:572-596): eye at the
    scene centre looking down +z; clip = scene diagonal (src/main.cpp:538-542)."""
    lo, hi = scene_bbox(tris)
    ext = hi - lo
    center = 0.5 * (lo + hi)
    diag = float(np.sqrt(np.dot(ext, ext)))
    if along_long_axis:
        eye = center - np.array([0.42 * ext[0], -0.05 * ext[1], 0.03 * ext[2]], np.float32)
        target = eye + np.array([1.0, -0.1, 0.05], np.float32)
    else:
        eye = center
        target = eye + np.array([0.0, 0.0, 1.0], np.float32)
    return primary_rays(eye, target, (0, 1, 0), 60.0, width, height, diag)


def cornell_view(width=256, height=256) -> np.ndarray:
    return primary_rays((278, 273, -800), (278, 273, 0), (0, 1, 0), 60.0, width, height, 2000.0)


def random_rays(tris: np.ndarray, count: int, seed: int = SEED, tmax=np.finfo(np.float32).max) -> np.ndarray:
    """C3: origins uniform in the scene box, directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    lo, hi = scene_bbox(tris)
    rays = np.empty(count, dtype=RAY_DTYPE)
    rays["org"] = (lo + (hi - lo) * rng.random((count, 3), dtype=np.float32)).astype(np.float32)
    d = rng.normal(size=(count, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays["dir"] = d
    rays["tmin"] = 0.0
    rays["tmax"] = tmax
    return rays


def bounce_rays(tris: np.ndarray, rays: np.ndarray, hit_ids: np.ndarray, hit_t: np.ndarray, seed: int = SEED) -> np.ndarray:
    """C5 second wave: cosine-weighted bounce off the primary hit point; rays that
    missed are re-emitted unchanged."""
    rng = np.random.default_rng(seed + 17)
    n = rays.shape[0]
    out = rays.copy()
    ok = hit_ids >= 0
    idx = np.nonzero(ok)[0]
    tr = tris[hit_ids[idx]]
    nrm = np.stack([tr["nx"], tr["ny"], tr["nz"]], axis=-1).astype(np.float64)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-30
    d = rays["dir"][idx].astype(np.float64)
    flip = np.sum(nrm * d, axis=1) > 0
    nrm[flip] *= -1
    p = rays["org"][idx] + d * hit_t[idx, None]
    lo, hi = scene_bbox(tris)
    diag = float(np.linalg.norm(hi - lo))
    p = p + nrm * (1e-3 * diag)
    u1, u2 = rng.random(idx.size), rng.random(idx.size)
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = np.cross(nrm, a)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(nrm, t1)
    nd = t1 * (r * np.cos(phi))[:, None] + t2 * (r * np.sin(phi))[:, None] + nrm * np.sqrt(1 - u1)[:, None]
    out["org"][idx] = p.astype(np.float32)
    out["dir"][idx] = nd.astype(np.float32)
    out["tmin"][idx] = 0.0
    out["tmax"][idx] = np.float32(diag)
    return out


def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def bounce_rays_f32(tris: np.ndarray, rays: np.ndarray, hits: np.ndarray, offset: float, tmax: float, seed: int) -> np.ndarray:
    """Host restatement of hgb_generate_bounce_rays (include/hagrid_b200.h) in numpy float32, one rounding per
    operation like the device kernel: what a front end without the device stage would compute on the CPU."""
    f = np.float32
    out = rays.copy()
    ids = hits["id"]
    idx = np.nonzero((ids >= 0) & (ids < tris.shape[0]))[0]
    tr = tris[ids[idx]]
    nx, ny, nz = tr["nx"].astype(f), tr["ny"].astype(f), tr["nz"].astype(f)
    ln = np.sqrt((nx * nx + ny * ny) + nz * nz)
    keep = ln > 0
    idx, nx, ny, nz, ln = idx[keep], nx[keep], ny[keep], nz[keep], ln[keep]
    nx, ny, nz = nx / ln, ny / ln, nz / ln
    d, o, t = rays["dir"][idx], rays["org"][idx], hits["t"][idx].astype(f)
    flip = ((nx * d[:, 0] + ny * d[:, 1]) + nz * d[:, 2]) > 0
    nx, ny, nz = np.where(flip, -nx, nx), np.where(flip, -ny, ny), np.where(flip, -nz, nz)
    n3 = np.stack([nx, ny, nz], axis=1)
    org = (o + d * t[:, None]) + n3 * f(offset)
    base = _mix32(np.uint32(seed & 0xFFFFFFFF) ^ _mix32(idx.astype(np.uint32)))
    dx, dy, s = np.zeros(idx.size, f), np.zeros(idx.size, f), np.zeros(idx.size, f)
    todo = np.ones(idx.size, bool)
    for k in range(8):
        x = f(2) * ((_mix32(base + np.uint32(2 * k)) >> np.uint32(8)).astype(f) * f(2.0 ** -24)) - f(1)
        y = f(2) * ((_mix32(base + np.uint32(2 * k + 1)) >> np.uint32(8)).astype(f) * f(2.0 ** -24)) - f(1)
        q = x * x + y * y
        take = todo & (q < 1)
        dx[take], dy[take], s[take] = x[take], y[take], q[take]
        todo &= ~take
    dz = np.sqrt(f(1) - s)
    zero = np.zeros(idx.size, f)
    near_x = np.abs(nx) > f(0.9)
    t1 = np.where(near_x[:, None], np.stack([-nz, zero, nx], axis=1), np.stack([zero, nz, -ny], axis=1)).astype(f)
    tl = np.sqrt((t1[:, 0] * t1[:, 0] + t1[:, 1] * t1[:, 1]) + t1[:, 2] * t1[:, 2])
    t1 = t1 / tl[:, None]
    t2 = np.stack([ny * t1[:, 2] - nz * t1[:, 1], nz * t1[:, 0] - nx * t1[:, 2], nx * t1[:, 1] - ny * t1[:, 0]], axis=1)
    new_dir = (t1 * dx[:, None] + t2 * dy[:, None]) + n3 * dz[:, None]
    out["org"][idx] = org
    out["dir"][idx] = new_dir
    out["tmin"][idx] = 0.0
    out["tmax"][idx] = f(tmax)
    return out


# ----------------------------------------------------------------------------- files for the reference CLI
def write_obj(path, tris: np.ndarray):
    """`v`/`f` only, 9 significant digits so float32 values survive the text round trip."""
    v0, v1, v2 = tri_vertices(tris)
    V = np.stack([v0, v1, v2], axis=1).reshape(-1, 3)
    n = tris.shape[0]
    with open(path, "w") as f:
        np.savetxt(f, V, fmt="v %.9g %.9g %.9g")
        F = np.arange(1, 3 * n + 1).reshape(n, 3)
        np.savetxt(f, F, fmt="f %d %d %d")


def write_rays(path, rays: np.ndarray):
    """.rays: 6 little-endian float32 per ray, org then dir (src/main.cpp:277-300)."""
    np.concatenate([rays["org"], rays["dir"]], axis=1).astype("<f4").tofile(path)
