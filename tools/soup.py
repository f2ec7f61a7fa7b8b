"""Synthetic triangle soup of BASELINE.json config 5: triangle i draws from the
reference's own RNG (tools.cl:2-4) seeded with WangHash(i*17+1); centre uniform in
[-5,5] x [0,3] x [-5,5] m, each vertex = centre + (U-0.5)*size per axis."""
import numpy as np


def _wang(s):
    s = s.astype(np.uint32)
    s = (s ^ np.uint32(61)) ^ (s >> np.uint32(16))
    s = s * np.uint32(9)
    s = s ^ (s >> np.uint32(4))
    s = s * np.uint32(0x27d4eb2d)
    s = s ^ (s >> np.uint32(15))
    return s


def _next(s):
    s ^= s << np.uint32(13)
    s ^= s >> np.uint32(17)
    s ^= s << np.uint32(5)
    return s, (s.astype(np.float32) * np.float32(2.3283064365387e-10))


def make_soup(n, size=0.01):
    """Returns (n,16) float32 triangles in the reference's Tri layout."""
    with np.errstate(over="ignore"):
        s = _wang(np.arange(n, dtype=np.uint32) * np.uint32(17) + np.uint32(1))
        u = []
        for _ in range(12):
            s, f = _next(s)
            u.append(f)
    c = np.stack([u[0] * 10 - 5, u[1] * 3, u[2] * 10 - 5], axis=1).astype(np.float32)
    tris = np.zeros((n, 16), dtype=np.float32)
    for k in range(3):
        off = np.stack([u[3 + 3 * k], u[4 + 3 * k], u[5 + 3 * k]], axis=1).astype(np.float32)
        tris[:, 4 * k:4 * k + 3] = c + (off - np.float32(0.5)) * np.float32(size)
    return tris


def soup_route():
    """12 lamp positions on the line x = 0, z in [-4, 4]; 60 s each."""
    z = np.linspace(-4, 4, 12, dtype=np.float32)
    return np.stack([np.zeros(12, np.float32), z, np.full(12, 60, np.float32)], axis=1)
