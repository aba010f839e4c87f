"""Vertex (block-row) partition into z-slabs for structured Kuhn grids: pure host arithmetic.

Rank r of P owns planes [k0, k1); its local mesh holds one extra ghost plane on each interior side. Local vertex
numbering is [owned | ghost-above | ghost-below] (owned first so that the local matrix rows are the owned vertices and the
ghost blocks sit in the tail of every vector).  `Mesh.kuhn(k_begin, k_end)` numbers plane-major, so a leading ghost plane
is rotated to the back with `Mesh.rotate_vertex_ids(shift)`.
The reference has no multi-device code; the contract is "distributed result == single-domain result".
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def split_planes(nz: int, nranks: int):
    """contiguous, balanced plane ranges; every rank gets at least 2 planes"""
    assert nz >= 2 * nranks, "need at least 2 planes per rank"
    base, rem = divmod(nz, nranks)
    bounds = [0]
    for r in range(nranks):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


@dataclass
class Slab:
    nx: int
    ny: int
    nz: int
    rank: int
    nranks: int
    k0: int
    k1: int
    kb: int            # first plane of the local mesh (k0 or k0-1)
    ke: int            # one past the last plane of the local mesh (k1 or k1+1)
    n_own: int
    n_ghost: int
    shift: int         # rotation that moves the ghost-below plane behind [owned | ghost-above]
    peers: list = field(default_factory=list)
    send_ptr: np.ndarray = None
    send_idx: np.ndarray = None
    recv_ptr: np.ndarray = None
    remote_off: np.ndarray = None   # per peer: offset (blocks) of this rank's packet in that peer's ghost tail
    int_begin: int = 0
    int_end: int = 0
    first_global_vtx: int = 0

    @property
    def nxy(self):
        return self.nx * self.ny

    def local_to_global(self):
        """global vertex id of every local vertex, in local numbering [owned | ghost-above | ghost-below]"""
        nxy = self.nxy
        own = np.arange(self.k0 * nxy, self.k1 * nxy, dtype=np.int64)
        parts = [own]
        if self.ke > self.k1:
            parts.append(np.arange(self.k1 * nxy, (self.k1 + 1) * nxy, dtype=np.int64))
        if self.kb < self.k0:
            parts.append(np.arange((self.k0 - 1) * nxy, self.k0 * nxy, dtype=np.int64))
        return np.concatenate(parts)

    def flags_clamp_z0(self, ndof=3):
        """config-2 Dirichlet flags (all dofs fixed on the global plane z = 0) for owned + ghost vertices"""
        g = self.local_to_global()
        flag = np.zeros((g.size, ndof), dtype=np.int32)
        flag[g < self.nxy] = 1
        return flag


def make_slab(nx, ny, nz, rank, nranks) -> Slab:
    b = split_planes(nz, nranks)
    k0, k1 = b[rank], b[rank + 1]
    has_below, has_above = rank > 0, rank < nranks - 1
    kb, ke = k0 - (1 if has_below else 0), k1 + (1 if has_above else 0)
    nxy = nx * ny
    n_own = (k1 - k0) * nxy
    n_ghost = (int(has_below) + int(has_above)) * nxy
    s = Slab(nx, ny, nz, rank, nranks, k0, k1, kb, ke, n_own, n_ghost, shift=(nxy if has_below else 0))
    peers, send_ptr, send_idx, recv_ptr, remote_off = [], [0], [], [0], []
    # ghost tail order is [above | below]; peers listed in the same order so recv ranges are contiguous
    if has_above:
        peers.append(rank + 1)
        send_idx.append(np.arange(n_own - nxy, n_own, dtype=np.uint32))  # my last owned plane
        send_ptr.append(send_ptr[-1] + nxy)
        recv_ptr.append(recv_ptr[-1] + nxy)
        remote_off.append(nxy if rank + 1 < nranks - 1 else 0)  # I am its ghost-BELOW plane: behind its ghost-above plane, if any
    if has_below:
        peers.append(rank - 1)
        send_idx.append(np.arange(0, nxy, dtype=np.uint32))  # my first owned plane
        send_ptr.append(send_ptr[-1] + nxy)
        recv_ptr.append(recv_ptr[-1] + nxy)
        remote_off.append(0)  # I am its ghost-ABOVE plane: first in its ghost tail
    s.peers = peers
    s.send_ptr = np.array(send_ptr, dtype=np.uint64)
    s.recv_ptr = np.array(recv_ptr, dtype=np.uint64)
    s.remote_off = np.array(remote_off, dtype=np.uint64)
    s.send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.uint32)
    # rows of the first / last owned plane reference ghosts; everything in between is interior
    s.int_begin = nxy if has_below else 0
    s.int_end = n_own - (nxy if has_above else 0)
    s.first_global_vtx = k0 * nxy
    return s


def local_mesh_host(s: Slab, h: float, index_dtype=np.uint32):
    """host twin of `Mesh.kuhn(k_begin=s.kb, k_end=s.ke)` followed by `rotate_vertex_ids(s.shift)`:
    (elem2vtx, vtx2xyz) of the rank's slab in local numbering [owned | ghost-above | ghost-below]"""
    from . import synthetic
    e2v, xyz = synthetic.kuhn_tet_mesh(s.nx, s.ny, s.nz, h, s.kb, s.ke, index_dtype=np.int64)
    nv = xyz.shape[0]
    if s.shift:
        e2v = np.where(e2v >= s.shift, e2v - s.shift, e2v + nv - s.shift)
        xyz = np.roll(xyz, -s.shift, axis=0)
    return e2v.astype(index_dtype), np.ascontiguousarray(xyz)
