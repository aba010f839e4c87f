"""Synthetic inputs of SURVEY §8(d) generated on the host with numpy (for the end-to-end path that starts from HOST
buffers, and for tests).  Not part of the reference; the device-side twin is `Mesh.kuhn`."""
from __future__ import annotations

import numpy as np

_PERM = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))
_ODD = (0, 1, 1, 0, 0, 1)

SIZES = {"S1": 70, "S10": 150, "S50": 256}  # vertices per edge of the cube


def kuhn_tet_mesh(nx, ny=None, nz=None, h=None, k_begin=0, k_end=None, index_dtype=np.uint32):
    """Kuhn-6 structured tet grid; planes [k_begin,k_end) with local vertex ids. -> (elem2vtx, vtx2xyz)"""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = 1.0 / (nx - 1) if h is None else h
    k_end = nz if k_end is None else k_end
    nzl = k_end - k_begin
    ii, jj, kk = np.arange(nx), np.arange(ny), np.arange(k_begin, k_end)
    xyz = np.empty((nzl, ny, nx, 3))
    xyz[..., 0] = (ii * h)[None, None, :]
    xyz[..., 1] = (jj * h)[None, :, None]
    xyz[..., 2] = (kk * h)[:, None, None]
    xyz = xyz.reshape(-1, 3)
    ci, cj, ck = np.arange(nx - 1, dtype=np.int64), np.arange(ny - 1, dtype=np.int64), np.arange(nzl - 1, dtype=np.int64)
    o = (ci[None, None, :] + nx * (cj[None, :, None] + ny * ck[:, None, None])).reshape(-1)
    stride = (1, nx, nx * ny)
    e2v = np.empty((o.size, 6, 4), dtype=index_dtype)
    c = o + stride[0] + stride[1] + stride[2]
    for t in range(6):
        a = o + stride[_PERM[t][0]]
        b = a + stride[_PERM[t][1]]
        if _ODD[t]:
            a, b = b, a
        e2v[:, t, 0], e2v[:, t, 1], e2v[:, t, 2], e2v[:, t, 3] = o, a, b, c
    return e2v.reshape(-1, 4), xyz


def grid_tri_mesh(n, disk=False, index_dtype=np.uint32):
    """n x n cells of [-1,1]^2 split into 2 n^2 triangles, optionally mapped onto the unit disk (the config-1 stand-in
    for the reference's Delaunay disk). -> (tri2vtx [2 n^2, 3], vtx2xy [(n+1)^2, 2])"""
    m = n + 1
    g = -1.0 + 2.0 * np.arange(m) / float(n)
    s, t = np.meshgrid(g, g)   # s varies fastest (i), t = rows (j)
    x, y = (s * np.sqrt(1.0 - 0.5 * t * t), t * np.sqrt(1.0 - 0.5 * s * s)) if disk else (s, t)
    xy = np.ascontiguousarray(np.stack([x.ravel(), y.ravel()], 1))
    i, j = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64))
    v00 = (i + m * j).ravel()
    v10, v01 = v00 + 1, v00 + m
    v11 = v01 + 1
    t2v = np.empty((v00.size, 2, 3), dtype=index_dtype)
    t2v[:, 0, 0], t2v[:, 0, 1], t2v[:, 0, 2] = v00, v10, v11
    t2v[:, 1, 0], t2v[:, 1, 1], t2v[:, 1, 2] = v00, v11, v01
    return t2v.reshape(-1, 3), xy


def splitmix_uniform(seed, first_idx, n, lo, hi):
    """same hash as dfb_vec_fill_splitmix (uniform in [lo,hi), indexed by first_idx + i)"""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.arange(first_idx, first_idx + n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def clamp_flags_z0(nx, ny, nzl, k_begin=0, ndof=3):
    """all dofs fixed on the z = 0 plane (config 2, cf. solid_linear_tri2.rs:124-126)"""
    flag = np.zeros((nx * ny * nzl, ndof), dtype=np.int32)
    if k_begin == 0:
        flag[: nx * ny, :] = 1
    return flag
