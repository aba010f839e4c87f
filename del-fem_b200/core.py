"""Handle classes over the C ABI (device-resident objects). Host arrays are numpy, borrowed per call."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DfbError, check, lib

ELEM = {"laplace_tri2": 0, "cot_laplace_tri3": 1, "solid_linear_tri2": 2, "laplace_tet": 3, "solid_linear_tet": 4,
        "neohookean_tet": 5, "spring3": 6, "stvk_tri3": 7, "bending_quad": 8, "mass_tet": 9}
ELEM_INFO = {0: (3, 2, 1), 1: (3, 3, 1), 2: (3, 2, 2), 3: (4, 3, 1), 4: (4, 3, 3), 5: (4, 3, 3), 6: (2, 3, 3), 7: (3, 3, 3), 8: (4, 3, 3), 9: (4, 3, 3)}
PRECOND = {"none": 0, "jacobi": 1, "block_jacobi": 2, "ilu0": 3}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Ctx:
    """one device + one stream (dfb_ctx). `Ctx.default()` returns a per-process context on LOCAL_RANK / device 0."""
    _default = None

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib().dfb_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        self.rank, self.nranks = 0, 1

    @classmethod
    def default(cls):
        if cls._default is None:
            import os
            cls._default = cls(int(os.environ.get("LOCAL_RANK", "0")))
        return cls._default

    def sync(self):
        check(lib().dfb_ctx_sync(self.h))

    def launch_count(self) -> int:
        v = C.c_uint64()
        check(lib().dfb_ctx_launch_count(self.h, C.byref(v)))
        return v.value

    def set_option(self, key: str, value: int):
        check(lib().dfb_ctx_set_option(self.h, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        check(lib().dfb_ctx_get_option(self.h, key.encode(), C.byref(v)))
        return v.value

    def timer_start(self):
        check(lib().dfb_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().dfb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(lib().dfb_flush_l2(self.h))

    def device_info(self):
        sm, l2, mem = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib().dfb_device_info(self.h, C.byref(sm), C.byref(l2), C.byref(mem)))
        return {"sm_count": sm.value, "l2_bytes": l2.value, "total_mem_bytes": mem.value}

    def host_register(self, a: np.ndarray):
        """page-lock a numpy buffer (so uploads from it are asynchronous DMA at full PCIe speed)"""
        check(lib().dfb_host_register(self.h, _ptr(a), a.nbytes))

    def host_unregister(self, a: np.ndarray):
        check(lib().dfb_host_unregister(self.h, _ptr(a)))

    def profile_spmv(self, reset=False):
        """(total_ms, count) of the SpMV launches bracketed with CUDA events while option profile_spmv = 1"""
        ms, n = C.c_double(), C.c_uint64()
        check(lib().dfb_profile_spmv(self.h, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    PHASES = ("elem", "gather", "grad", "pack", "tile", "bc", "precond", "unused")

    def phase_times(self, reset=False):
        """{phase: (total_ms, count)} of the library's own assembly / setup launches while option profile_phases = 1"""
        ms = np.zeros(8)
        cnt = np.zeros(8, dtype=np.uint64)
        check(lib().dfb_ctx_phase_times(self.h, _ptr(ms), _ptr(cnt), 1 if reset else 0))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PHASES)}

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        check(lib().dfb_ctx_comm_init(self.h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this rank's scalar mailbox"""
        buf = (C.c_uint8 * 64)()
        check(lib().dfb_ctx_peer_export(self.h, buf))
        return bytes(buf)

    def peer_import(self, handles):
        """handles: list of the nranks exported handles, in rank order"""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(lib().dfb_ctx_peer_import(self.h, buf, len(handles)))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib().dfb_comm_unique_id(buf))
        return bytes(buf)

    def close(self):
        if self.h:
            lib().dfb_ctx_destroy(self.h)
            self.h = None


class Vec:
    """device `[[f64; N]]` (slice_of_array.rs)"""

    def __init__(self, ctx: Ctx, n_blk: int, blk_dim: int, n_ghost: int = 0, data=None, _borrow=None):
        self.ctx, self.n_blk, self.blk_dim, self.n_ghost = ctx, int(n_blk), int(blk_dim), int(n_ghost)
        self._owned = _borrow is None
        if _borrow is not None:  # a handle owned by another object (linearsystem.Solver's vectors)
            self.h = C.c_void_p(_borrow)
            return
        h = C.c_void_p()
        check(lib().dfb_vec_create(ctx.h, self.n_blk, self.n_ghost, self.blk_dim, C.byref(h)))
        self.h = h
        if data is not None:
            self.upload(data)

    @classmethod
    def from_numpy(cls, ctx, a, n_ghost=0):
        a = np.asarray(a, dtype=np.float64)
        a2 = a.reshape(a.shape[0], -1)
        return cls(ctx, a2.shape[0], a2.shape[1], n_ghost, a2)

    def upload(self, a):
        a = _c(a, np.float64)
        if a.size != self.n_blk * self.blk_dim:
            raise DfbError(1, f"Vec.upload: size {a.size} != {self.n_blk}*{self.blk_dim}")
        check(lib().dfb_vec_upload(self.h, _ptr(a)))

    def download(self):
        out = np.empty((self.n_blk, self.blk_dim))
        check(lib().dfb_vec_download(self.h, _ptr(out)))
        return out

    def set_zero(self):
        check(lib().dfb_vec_set_zero(self.h))

    def copy_from(self, other):
        check(lib().dfb_vec_copy(self.h, other.h))

    def dot(self, other) -> float:
        v = C.c_double()
        check(lib().dfb_vec_dot(self.h, other.h, C.byref(v)))
        return v.value

    def add_scaled(self, alpha, p):
        check(lib().dfb_vec_add_scaled(self.h, float(alpha), p.h))

    def scale_and_add(self, beta, r):
        check(lib().dfb_vec_scale_and_add(self.h, float(beta), r.h))

    def scale(self, s):
        check(lib().dfb_vec_scale(self.h, float(s)))

    def set_fixed_dof_zero(self, blk2isfix):
        f = _c(blk2isfix, np.int32)
        if f.size != self.n_blk * self.blk_dim:
            raise DfbError(1, "Vec.set_fixed_dof_zero: flag size mismatch")
        check(lib().dfb_vec_set_fixed_dof_zero(self.h, _ptr(f)))

    def set_fixed_dof_zero_dev(self, flag: "BcFlag"):
        check(lib().dfb_vec_set_fixed_dof_zero_dev(self.h, flag.h))

    def fill_splitmix(self, seed, first_idx, lo, hi):
        check(lib().dfb_vec_fill_splitmix(self.h, int(seed), int(first_idx), float(lo), float(hi)))

    def device_ptr(self):
        return lib().dfb_vec_device_ptr(self.h)

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owned", True):
            lib().dfb_vec_destroy(self.h)
        self.h = None


class BcFlag:
    """device-resident `&[[i32; N]]` Dirichlet flags"""

    def __init__(self, ctx: Ctx, blk2isfix):
        f = _c(blk2isfix, np.int32)
        self.ctx = ctx
        h = C.c_void_p()
        check(lib().dfb_bcflag_create(ctx.h, _ptr(f), f.shape[0], f.shape[1], C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().dfb_bcflag_destroy(self.h)
            self.h = None


class Mesh:
    """device-resident uniform mesh (elem2vtx + vtx2xyz)"""

    def __init__(self, ctx: Ctx, elem2vtx=None, vtx2xyz=None, _handle=None):
        self.ctx = ctx
        if _handle is not None:
            self.h = _handle
        else:
            xyz = _c(vtx2xyz, np.float64)
            e2v = np.asarray(elem2vtx)
            h = C.c_void_p()
            if e2v.dtype == np.uint32:
                e2v = _c(e2v, np.uint32)
                check(lib().dfb_mesh_create_u32(ctx.h, _ptr(e2v), e2v.shape[0], e2v.shape[1], _ptr(xyz), xyz.shape[0],
                                                xyz.shape[1], C.byref(h)))
            else:
                e2v = _c(e2v, np.uint64)
                check(lib().dfb_mesh_create(ctx.h, _ptr(e2v), e2v.shape[0], e2v.shape[1], _ptr(xyz), xyz.shape[0],
                                            xyz.shape[1], C.byref(h)))
            self.h = h
        ne, nn, nv, nd = C.c_uint64(), C.c_int32(), C.c_uint64(), C.c_int32()
        check(lib().dfb_mesh_sizes(self.h, C.byref(ne), C.byref(nn), C.byref(nv), C.byref(nd)))
        self.num_elem, self.num_node, self.num_vtx, self.ndim = ne.value, nn.value, nv.value, nd.value

    @classmethod
    def kuhn(cls, ctx, nx, ny=None, nz=None, h=None, k_begin=0, k_end=None):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        h = 1.0 / (nx - 1) if h is None else h
        k_end = nz if k_end is None else k_end
        hd = C.c_void_p()
        check(lib().dfb_mesh_create_kuhn(ctx.h, nx, ny, nz, float(h), k_begin, k_end, C.byref(hd)))
        return cls(ctx, _handle=hd)

    def rotate_vertex_ids(self, shift):
        check(lib().dfb_mesh_rotate_vertex_ids(self.h, int(shift)))

    def update_xyz(self, vtx2xyz):
        xyz = _c(vtx2xyz, np.float64)
        if xyz.size != self.num_vtx * self.ndim:
            raise DfbError(1, "Mesh.update_xyz: size mismatch")
        check(lib().dfb_mesh_update_xyz(self.h, _ptr(xyz)))

    def download(self):
        e2v = np.empty((self.num_elem, self.num_node), dtype=np.uint64)
        xyz = np.empty((self.num_vtx, self.ndim))
        check(lib().dfb_mesh_download(self.h, _ptr(e2v), _ptr(xyz)))
        return e2v, xyz

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().dfb_mesh_destroy(self.h)
            self.h = None


class Pattern:
    """row2idx / idx2col on the device. convention 0 = csrdia, 1 = blkcsr"""

    def __init__(self, ctx: Ctx, row2idx=None, idx2col=None, convention=0, _handle=None, _owned=True):
        self.ctx = ctx
        self._owned = _owned
        if _handle is not None:
            self.h = _handle
        else:
            r = np.asarray(row2idx)
            h = C.c_void_p()
            if r.dtype == np.uint32:
                r = _c(r, np.uint32)
                c = _c(idx2col, np.uint32)
                check(lib().dfb_pattern_create_u32(ctx.h, _ptr(r), _ptr(c), r.size - 1, convention, C.byref(h)))
            else:
                r = _c(r, np.uint64)
                c = _c(idx2col, np.uint64)
                if r.size < 1 or (r.size and int(r[-1]) != c.size):
                    raise DfbError(1, "Pattern: row2idx[num_blk] != idx2col.len() (assert_eq!(num_idx, idx2col.len()))")
                check(lib().dfb_pattern_create(ctx.h, _ptr(r), _ptr(c), r.size - 1, convention, C.byref(h)))
            self.h = h
        nb, nnz, conv = C.c_uint64(), C.c_uint64(), C.c_int32()
        check(lib().dfb_pattern_sizes(self.h, C.byref(nb), C.byref(nnz), C.byref(conv)))
        self.num_blk, self.nnz, self.convention = nb.value, nnz.value, conv.value

    @classmethod
    def from_uniform_mesh(cls, ctx, mesh: Mesh, is_self: bool, num_rows=None):
        """del_msh_cpu::vtx2vtx::from_uniform_mesh on the device"""
        h = C.c_void_p()
        check(lib().dfb_pattern_from_uniform_mesh(ctx.h, mesh.h, 1 if is_self else 0,
                                                  mesh.num_vtx if num_rows is None else num_rows, C.byref(h)))
        return cls(ctx, _handle=h)

    def download(self):
        r = np.empty(self.num_blk + 1, dtype=np.uint64)
        c = np.empty(self.nnz, dtype=np.uint64)
        check(lib().dfb_pattern_download(self.h, _ptr(r), _ptr(c)))
        return r, c

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owned", True):
            lib().dfb_pattern_destroy(self.h)
        self.h = None


class Plan:
    """assembly plan: inverse of the element->nonzero map (contribution lists per nonzero, ascending element)"""

    def __init__(self, ctx: Ctx, pattern: Pattern, mesh: Mesh, flavour: int = 0):
        self.ctx, self.pattern, self.mesh = ctx, pattern, mesh
        h = C.c_void_p()
        check(lib().dfb_plan_create(ctx.h, pattern.h, mesh.h, flavour, C.byref(h)))
        self.h = h
        ns, nc = C.c_uint64(), C.c_uint64()
        check(lib().dfb_plan_sizes(self.h, C.byref(ns), C.byref(nc)))
        self.num_slots, self.num_contrib = ns.value, nc.value

    def elem2nz(self):
        n = self.mesh.num_node
        out = np.empty((self.mesh.num_elem, n, n), dtype=np.uint32)
        check(lib().dfb_plan_download_elem2nz(self.h, _ptr(out)))
        return out

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().dfb_plan_destroy(self.h)
            self.h = None


class Halo:
    def __init__(self, ctx: Ctx, peers, send_ptr, send_idx, recv_ptr):
        self.ctx = ctx
        peers = _c(peers, np.int32)
        send_ptr = _c(send_ptr, np.uint64)
        recv_ptr = _c(recv_ptr, np.uint64)
        send_idx = _c(send_idx, np.uint32)
        self.n_ghost = int(recv_ptr[-1]) if recv_ptr.size else 0
        h = C.c_void_p()
        check(lib().dfb_halo_create(ctx.h, peers.size, _ptr(peers), _ptr(send_ptr), _ptr(send_idx), _ptr(recv_ptr), C.byref(h)))
        self.h = h

    def exchange(self, v: Vec):
        check(lib().dfb_halo_exchange(self.h, v.h))

    def set_remote(self, remote_off):
        """remote_off[p]: where (in blocks) this rank's packet lands in peer p's ghost tail -> enables the NVLink peer push in (P)CG"""
        ro = _c(remote_off, np.uint64)
        check(lib().dfb_halo_set_remote(self.h, _ptr(ro)))

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            lib().dfb_halo_destroy(self.h)
            self.h = None


def device_malloc(ctx: Ctx, nbytes: int):
    p = C.c_void_p()
    check(lib().dfb_malloc(ctx.h, int(nbytes), C.byref(p)))
    return p


def device_free(ctx: Ctx, p):
    check(lib().dfb_free(ctx.h, p))


def memcpy_d2h(ctx: Ctx, host: np.ndarray, dev, nbytes=None):
    check(lib().dfb_memcpy_d2h(ctx.h, _ptr(host), dev, host.nbytes if nbytes is None else nbytes))


def memcpy_h2d(ctx: Ctx, dev, host: np.ndarray):
    host = np.ascontiguousarray(host)
    check(lib().dfb_memcpy_h2d(ctx.h, dev, _ptr(host), host.nbytes))
