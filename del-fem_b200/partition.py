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


# ------------------------------------------------------------------------------------------------ general meshes
# Any uniform mesh: vertices are owned in contiguous GLOBAL index ranges (so a locality-preserving numbering — `rcm_order` below —
# plays the role METIS/RCM plays upstream of other FEM codes).  Same contract as the slabs: every rank merges, in ascending element
# order, the elements that touch its owned vertices, so owned rows equal the global assembly; ghosts sit in the tail of every vector.
# Local numbering is [interior owned | boundary owned | ghosts by owner, then global id]: the rows that read ghost columns are the
# LAST owned rows, which is what `Matrix.set_halo(halo, int_begin=0, int_end=n_interior)` and the SpMV's tile order expect.
@dataclass
class Part:
    rank: int
    nranks: int
    bounds: np.ndarray          # [nranks+1] global vertex ranges
    n_own: int
    n_ghost: int
    local2global: np.ndarray    # [n_own + n_ghost]
    elem_ids: np.ndarray        # global ids of the local elements, ascending
    elem2vtx: np.ndarray        # [n_local_elem, n_node] local numbering
    peers: list = field(default_factory=list)
    send_ptr: np.ndarray = None
    send_idx: np.ndarray = None
    recv_ptr: np.ndarray = None
    remote_off: np.ndarray = None
    int_begin: int = 0
    int_end: int = 0

    def local_to_global(self):
        return self.local2global


def split_rows(num_vtx: int, nranks: int):
    base, rem = divmod(num_vtx, nranks)
    b = [0]
    for r in range(nranks):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return np.asarray(b, dtype=np.int64)


def partition_general(elem2vtx, num_vtx: int, nranks: int, rank: int, bounds=None, index_dtype=np.uint32) -> Part:
    e2v = np.asarray(elem2vtx, dtype=np.int64)
    assert e2v.ndim == 2
    bounds = split_rows(num_vtx, nranks) if bounds is None else np.asarray(bounds, dtype=np.int64)
    assert bounds.size == nranks + 1 and bounds[0] == 0 and bounds[-1] == num_vtx
    owner_of = lambda v: np.searchsorted(bounds, v, side="right") - 1  # noqa: E731
    eo = owner_of(e2v)
    b0, b1 = int(bounds[rank]), int(bounds[rank + 1])
    local = np.nonzero((eo == rank).any(axis=1))[0]
    le, leo = e2v[local], eo[local]
    ghosts = np.unique(le[leo != rank])                      # ascending global id == ascending (owner, id): ranges are contiguous
    mixed = (leo != rank).any(axis=1)
    bnd = np.unique(le[mixed][leo[mixed] == rank])            # owned vertices sharing an element with a ghost
    owned = np.arange(b0, b1, dtype=np.int64)
    is_bnd = np.zeros(b1 - b0, dtype=bool)
    is_bnd[bnd - b0] = True
    interior = owned[~is_bnd]
    l2g = np.concatenate([interior, bnd, ghosts])
    g2l = np.full(num_vtx, -1, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    part = Part(rank, nranks, bounds, b1 - b0, int(ghosts.size), l2g, local, g2l[le].astype(index_dtype))
    part.int_begin, part.int_end = 0, int(interior.size)
    g_owner = owner_of(ghosts)
    peers, send_ptr, send_idx, recv_ptr, remote_off = [], [0], [], [0], []
    for q in np.unique(g_owner):
        q = int(q)
        peers.append(q)
        recv_ptr.append(recv_ptr[-1] + int(np.count_nonzero(g_owner == q)))
        # what q holds as ghosts of mine: my vertices in elements that also hold a vertex of q (those elements are local to q)
        touch_q = (leo == q).any(axis=1)
        mine_q = np.unique(le[touch_q][leo[touch_q] == rank])
        send_idx.append(g2l[mine_q])
        send_ptr.append(send_ptr[-1] + int(mine_q.size))
        # my packet's offset in q's ghost tail = number of q's ghosts with a smaller global id than my range
        eq = (eo == q).any(axis=1)
        ghosts_q = np.unique(e2v[eq][eo[eq] != q])
        remote_off.append(int(np.searchsorted(ghosts_q, b0)))
    part.peers = peers
    part.send_ptr = np.asarray(send_ptr, dtype=np.uint64)
    part.recv_ptr = np.asarray(recv_ptr, dtype=np.uint64)
    part.remote_off = np.asarray(remote_off, dtype=np.uint64)
    part.send_idx = (np.concatenate(send_idx) if send_idx else np.zeros(0)).astype(np.uint32)
    return part


def rcm_order(elem2vtx, num_vtx: int):
    """reverse Cuthill-McKee numbering of the vertex graph (scipy): perm[new] = old. A band-limited numbering keeps the contiguous
    ownership ranges compact (few neighbours, small halos) — the role of the reordering step SURVEY §8f-3 mentions."""
    import scipy.sparse
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    e2v = np.asarray(elem2vtx, dtype=np.int64)
    nn = e2v.shape[1]
    rows = np.repeat(e2v, nn, axis=1).ravel()
    cols = np.tile(e2v, (1, nn)).ravel()
    a = scipy.sparse.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(num_vtx, num_vtx))
    return np.asarray(reverse_cuthill_mckee(a, symmetric_mode=True), dtype=np.int64)


def renumber_mesh(elem2vtx, vtx2xyz, perm):
    """apply perm[new] = old: returns (elem2vtx in new ids, vtx2xyz reordered, old2new)"""
    perm = np.asarray(perm, dtype=np.int64)
    old2new = np.empty_like(perm)
    old2new[perm] = np.arange(perm.size)
    return old2new[np.asarray(elem2vtx, dtype=np.int64)], np.ascontiguousarray(np.asarray(vtx2xyz)[perm]), old2new
