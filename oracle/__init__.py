"""ctypes front-end of the CPU oracle (oracle/delfem_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  May be imported by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` — never by the product package `del-fem_b200/`.
Every wrapper names the reference function (file:line) it restates; see the C++ file for details.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

U64 = np.uint64
F64 = np.float64


def build(force: bool = False) -> str:
    """compile oracle/_build/liboracle.so (g++); building the checker is not using it."""
    src = os.path.join(_HERE, "delfem_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_tet_volume.restype = C.c_double
        _lib.orc_dot.restype = C.c_double
        _lib.orc_vtx2vtx_from_uniform_mesh.restype = C.c_uint64
        _lib.orc_conjugate_gradient.restype = C.c_int64
        _lib.orc_preconditioned_conjugate_gradient.restype = C.c_int64
        _lib.orc_pcg_jacobi3_mt.restype = C.c_int64
        _lib.orc_symbolic_iluk.restype = C.c_uint64
        _lib.orc_mult_square_matrices.restype = C.c_uint64
        _lib.orc_ilu_new.restype = C.c_void_p
        _lib.orc_ilu_row2val.restype = C.POINTER(C.c_double)
        _lib.orc_version.restype = C.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u64(a):
    return np.ascontiguousarray(a, dtype=U64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=F64)


class OraclePanic(RuntimeError):
    """the reference would have panicked (assert!/unwrap) on this input"""


def _chk(rc):
    if rc != 0:
        raise OraclePanic("reference contract violation (assert!/panic! in the Rust code)")


# ---------------------------------------------------------------- element kernels
def laplace_tri2_ddw(alpha, p0, p1, p2):
    """del-fem-cpu/src/laplace_tri2.rs:3-17 -> [3][3][1]"""
    out = np.zeros((3, 3, 1))
    lib().orc_laplace_tri2_ddw(C.c_double(alpha), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(out))
    return out


def cot_laplace_tri3(p0, p1, p2):
    """del_geo_core::tri3::emat_cotangent_laplacian (call site laplace_tri3.rs:18) -> [3][3][1]"""
    out = np.zeros((3, 3, 1))
    lib().orc_cot_laplace_tri3(_p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(out))
    return out


def emat_tri2(lam, mu, p0, p1, p2):
    """del-fem-cpu/src/solid_linear_tri2.rs:23-33 -> [3][3][4]"""
    out = np.zeros((3, 3, 4))
    lib().orc_emat_tri2(C.c_double(lam), C.c_double(mu), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(out))
    return out


def laplace_tet_ddw(alpha, p0, p1, p2, p3):
    """DERIVED -> [4][4][1]"""
    out = np.zeros((4, 4, 1))
    lib().orc_laplace_tet_ddw(C.c_double(alpha), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(_f64(p3)), _p(out))
    return out


def emat_tet(lam, mu, p0, p1, p2, p3):
    """DERIVED -> [4][4][9] column-major blocks"""
    out = np.zeros((4, 4, 9))
    lib().orc_emat_tet(C.c_double(lam), C.c_double(mu), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(_f64(p3)), _p(out))
    return out


def tet_volume(p0, p1, p2, p3):
    return lib().orc_tet_volume(_p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(_f64(p3)))


def wr_dwrdc_ddwrddc_neohookean(mu, lam, c):
    """DERIVED energy density, same return shape as solid_incompressive_hex.rs:3-53"""
    c = _f64(c).reshape(3, 3)
    w = C.c_double(0)
    dw = np.zeros(6)
    ddw = np.zeros((6, 6))
    lib().orc_wr_dwrdc_ddwrddc_neohookean(C.c_double(mu), C.c_double(lam), _p(c), C.byref(w), _p(dw), _p(ddw))
    return w.value, dw, ddw


def neohookean_tet_wdwddw(mu, lam, node2xyz, node2disp):
    """DERIVED: chain rule of solid_incompressive_hex.rs:86-148 with 4 nodes -> (w, dw[4][3], ddw[4][4][9] a*3+b)"""
    w = C.c_double(0)
    dw = np.zeros((4, 3))
    ddw = np.zeros((4, 4, 9))
    lib().orc_neohookean_tet_wdwddw(C.c_double(mu), C.c_double(lam), _p(_f64(node2xyz)), _p(_f64(node2disp)),
                                    C.byref(w), _p(dw), _p(ddw))
    return w.value, dw, ddw


def spring3_wdwddw(stiffness, node2xyz_def, edge_length_ini):
    """del-fem-cpu/src/spring3.rs:1-28"""
    w = C.c_double(0)
    dw = np.zeros((2, 3))
    ddw = np.zeros((2, 2, 9))
    lib().orc_spring3_wdwddw(C.c_double(stiffness), _p(_f64(node2xyz_def)), C.c_double(edge_length_ini),
                             C.byref(w), _p(dw), _p(ddw))
    return w.value, dw, ddw


def stvk_tri3_wdw(p_ini, p_def):
    """del-fem-cpu/src/stvk_tri3.rs:9-98 -> (w[3], dw[3][3][3] = dw[strain][node][dim])"""
    w = np.zeros(3)
    dw = np.zeros((3, 3, 3))
    lib().orc_stvk_tri3_wdw(_p(_f64(p_ini)), _p(_f64(p_def)), _p(w), _p(dw))
    return w, dw


def stvk_tri3_wdwddw(p0, p1, lam, myu):
    """del-fem-cpu/src/stvk_tri3.rs:152-330 -> (w, dw[3][3], ddw[3][3][9] with idim*3+jdim)"""
    w = C.c_double(0)
    dw = np.zeros((3, 3))
    ddw = np.zeros((3, 3, 9))
    lib().orc_stvk_tri3_wdwddw(_p(_f64(p0)), _p(_f64(p1)), C.c_double(lam), C.c_double(myu), C.byref(w), _p(dw), _p(ddw))
    return w.value, dw, ddw


def quadratic_bending_wdw(p_ini, p_def):
    """del-fem-cpu/src/quadratic_bending.rs:1-43 -> (w[3], dwdp[3][4][3], k[4])"""
    w = np.zeros(3)
    dwdp = np.zeros((3, 4, 3))
    k = np.zeros(4)
    lib().orc_quadratic_bending_wdw(_p(_f64(p_ini)), _p(_f64(p_def)), _p(w), _p(dwdp), _p(k))
    return w, dwdp, k


def quadratic_bending_wdwddw(stiffness, p_ini, p_def):
    """DERIVED energy stiffness/2 |w|^2 of quadratic_bending.rs:1-43 -> (w, dw[4][3], ddw[4][4][9] column-major)"""
    w = C.c_double(0)
    dw = np.zeros((4, 3))
    ddw = np.zeros((4, 4, 9))
    lib().orc_quadratic_bending_wdwddw(C.c_double(stiffness), _p(_f64(p_ini)), _p(_f64(p_def)), C.byref(w), _p(dw), _p(ddw))
    return w.value, dw, ddw


def emat_mass_tet(rho, p0, p1, p2, p3):
    """A9' DERIVED consistent tet mass rho V / 20 (1 + delta_ij) I_3 -> [4][4][9]"""
    e = np.zeros((4, 4, 9))
    lib().orc_emat_mass_tet(C.c_double(rho), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), _p(_f64(p3)), _p(e))
    return e


def mass_lumped_plate_bending(tri2vtx, vtx2xy, thick, rho):
    """del-fem-cpu/src/mitc_tri3.rs:304-332 -> vtx2mass [nv, 3]"""
    t2v = _u64(tri2vtx).reshape(-1, 3)
    xy = _f64(vtx2xy).reshape(-1, 2)
    out = np.zeros((xy.shape[0], 3))
    lib().orc_mass_lumped_plate_bending(_p(t2v), C.c_uint64(t2v.shape[0]), _p(xy), C.c_uint64(xy.shape[0]), C.c_double(thick),
                                        C.c_double(rho), _p(out))
    return out


def lumped_mass(elem2vtx, n_node, vtx2xyz, ndim, rho):
    """A9' DERIVED lumped mass (measure / n_node * rho per element node), loop shape of mitc_tri3.rs:321-331 -> [nv]"""
    e2v = _u64(elem2vtx).reshape(-1, n_node)
    xyz = _f64(vtx2xyz).reshape(-1, ndim)
    out = np.zeros(xyz.shape[0])
    _chk(lib().orc_lumped_mass(_p(e2v), C.c_uint64(e2v.shape[0]), C.c_uint64(n_node), _p(xyz), C.c_uint64(xyz.shape[0]),
                               C.c_uint64(ndim), C.c_double(rho), _p(out)))
    return out


# ---------------------------------------------------------------- meshes / pattern
def kuhn_tet_mesh(nx, ny=None, nz=None, h=None):
    """SURVEY B.6 synthetic mesh: (elem2vtx [ne,4] u64, vtx2xyz [nv,3])"""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = 1.0 / (nx - 1) if h is None else h
    ne = 6 * (nx - 1) * (ny - 1) * (nz - 1)
    e2v = np.zeros((ne, 4), dtype=U64)
    xyz = np.zeros((nx * ny * nz, 3))
    lib().orc_kuhn_tet_mesh(C.c_uint64(nx), C.c_uint64(ny), C.c_uint64(nz), C.c_double(h), _p(e2v), _p(xyz))
    return e2v, xyz


def grid_tri_mesh(n, disk=False):
    """(tri2vtx [2n^2,3] u64, vtx2xy [(n+1)^2,2])"""
    t2v = np.zeros((2 * n * n, 3), dtype=U64)
    xy = np.zeros(((n + 1) ** 2, 2))
    lib().orc_grid_tri_mesh(C.c_uint64(n), C.c_int(1 if disk else 0), _p(t2v), _p(xy))
    return t2v, xy


def splitmix_fill(seed, first_idx, n, lo, hi):
    out = np.zeros(n)
    lib().orc_splitmix_fill(C.c_uint64(seed), C.c_uint64(first_idx), C.c_uint64(n), C.c_double(lo), C.c_double(hi), _p(out))
    return out


def vtx2vtx_from_uniform_mesh(elem2vtx, num_node, num_vtx, is_self):
    """del_msh_cpu::vtx2vtx::from_uniform_mesh restated -> (row2idx, idx2col) u64"""
    e2v = _u64(elem2vtx).reshape(-1)
    ne = e2v.size // num_node
    r2i = np.zeros(num_vtx + 1, dtype=U64)
    nnz = lib().orc_vtx2vtx_from_uniform_mesh(_p(e2v), C.c_uint64(ne), C.c_uint64(num_node), C.c_uint64(num_vtx),
                                              C.c_int(int(is_self)), _p(r2i), None)
    i2c = np.zeros(nnz, dtype=U64)
    lib().orc_vtx2vtx_from_uniform_mesh(_p(e2v), C.c_uint64(ne), C.c_uint64(num_node), C.c_uint64(num_vtx),
                                        C.c_int(int(is_self)), _p(r2i), _p(i2c))
    return r2i, i2c


# ---------------------------------------------------------------- merge
def merge_csrdia(node2row, node2col, emat, row2idx, idx2col, row2val, idx2val, merge_buffer=None):
    """del-fem-cpu/src/merge.rs:3-48 (in-place on row2val / idx2val)"""
    emat = _f64(emat)
    n_node = emat.shape[0]
    blksize = emat.shape[2] if emat.ndim == 3 else 1
    nb = row2idx.size - 1
    buf = np.full(nb, np.iinfo(U64).max, dtype=U64) if merge_buffer is None else merge_buffer
    _chk(lib().orc_merge_csrdia(C.c_uint64(blksize), C.c_uint64(n_node), _p(_u64(node2row)), _p(_u64(node2col)), _p(emat),
                                _p(row2idx), C.c_uint64(nb), _p(idx2col), C.c_uint64(idx2col.size), _p(row2val),
                                _p(idx2val), _p(buf)))


def merge_blkcsr(node2row, node2col, emat, row2idx, idx2col, idx2val, merge_buffer=None):
    """del-fem-cpu/src/merge.rs:51-88"""
    emat = _f64(emat)
    n_node = emat.shape[0]
    blksize = emat.shape[2]
    nb = row2idx.size - 1
    buf = np.full(nb, np.iinfo(U64).max, dtype=U64) if merge_buffer is None else merge_buffer
    _chk(lib().orc_merge_blkcsr(C.c_uint64(blksize), C.c_uint64(n_node), _p(_u64(node2row)), _p(_u64(node2col)), _p(emat),
                                _p(row2idx), C.c_uint64(nb), _p(idx2col), C.c_uint64(idx2col.size), _p(idx2val), _p(buf)))


def merge_for_array_blk(emat, node2vtx, row2idx, idx2col, row2val, idx2val, merge_buffer=None, flavour=0):
    """del-fem-cpu/src/sparse_square.rs:180-214 (flavour 0) / del-fem-ls/src/sparse_square.rs:66-104 (flavour 1)"""
    emat = _f64(emat)
    n_node = emat.shape[0]
    nn = emat.shape[2]
    nb = row2idx.size - 1
    buf = np.full(nb, np.iinfo(U64).max, dtype=U64) if merge_buffer is None else merge_buffer
    _chk(lib().orc_merge_for_array_blk(C.c_uint64(nn), C.c_uint64(n_node), _p(_u64(node2vtx)), _p(emat), _p(row2idx),
                                       C.c_uint64(nb), _p(idx2col), _p(row2val), _p(idx2val), _p(buf), C.c_int(flavour)))


KIND = {"laplace_tri2": 0, "cot_laplace_tri3": 1, "solid_linear_tri2": 2, "laplace_tet": 3, "solid_linear_tet": 4, "mass_tet": 9}
KIND_INFO = {  # kind -> (n_node, ndim_x, blk_dim N)
    0: (3, 2, 1), 1: (3, 3, 1), 2: (3, 2, 2), 3: (4, 3, 1), 4: (4, 3, 3), 9: (4, 3, 3)}


def assemble(kind, convention, elem2vtx, vtx2xyz, p0, p1, row2idx, idx2col):
    """whole-mesh element loop + merge in ascending element order (solid_linear_tri2.rs:144-161 etc.)
    returns (row2val [nb,NN] (zeros for convention 1), idx2val [nnz,NN])"""
    k = KIND[kind] if isinstance(kind, str) else kind
    n_node, ndx, n = KIND_INFO[k]
    e2v = _u64(elem2vtx).reshape(-1, n_node)
    xyz = _f64(vtx2xyz).reshape(-1, ndx)
    nb = row2idx.size - 1
    row2val = np.zeros((nb, n * n))
    idx2val = np.zeros((idx2col.size, n * n))
    _chk(lib().orc_assemble(C.c_int(k), C.c_int(convention), _p(e2v), C.c_uint64(e2v.shape[0]), _p(xyz), C.c_double(p0),
                            C.c_double(p1), _p(row2idx), C.c_uint64(nb), _p(idx2col), C.c_uint64(idx2col.size),
                            _p(row2val), _p(idx2val)))
    return row2val, idx2val


def assemble_neohookean_tet(elem2vtx, vtx2xyz, vtx2disp, mu, lam, row2idx, idx2col):
    e2v = _u64(elem2vtx).reshape(-1, 4)
    nb = row2idx.size - 1
    row2val = np.zeros((nb, 9))
    idx2val = np.zeros((idx2col.size, 9))
    dw = np.zeros((nb, 3))
    w = C.c_double(0)
    _chk(lib().orc_assemble_neohookean_tet(_p(e2v), C.c_uint64(e2v.shape[0]), _p(_f64(vtx2xyz)), _p(_f64(vtx2disp)),
                                           C.c_double(mu), C.c_double(lam), _p(row2idx), C.c_uint64(nb), _p(idx2col),
                                           _p(row2val), _p(idx2val), _p(dw), C.byref(w)))
    return w.value, dw, row2val, idx2val


def assemble_spring3(edge2vtx, vtx2xyz, vtx2disp, stiffness, row2idx, idx2col):
    """edge loop of wdwddw_hair_system (rod3_darboux.rs:760-793) with 3x3 blocks -> (w, dw, row2val, idx2val)"""
    e2v = _u64(edge2vtx).reshape(-1, 2)
    nb = row2idx.size - 1
    row2val = np.zeros((nb, 9))
    idx2val = np.zeros((idx2col.size, 9))
    dw = np.zeros((nb, 3))
    w = C.c_double(0)
    _chk(lib().orc_assemble_spring3(_p(e2v), C.c_uint64(e2v.shape[0]), _p(_f64(vtx2xyz)), _p(_f64(vtx2disp)), C.c_double(stiffness),
                                    _p(row2idx), C.c_uint64(nb), _p(idx2col), _p(row2val), _p(idx2val), _p(dw), C.byref(w)))
    return w.value, dw, row2val, idx2val


def assemble_cloth(kind, elem2vtx, vtx2xyz, vtx2disp, p0, p1, row2idx, idx2col):
    """kind "stvk_tri3" (p0 = lambda, p1 = myu; stvk_tri3.rs:152) or "bending_quad" (p0 = stiffness; DERIVED from
    quadratic_bending.rs:1) whole-mesh loop -> (w, dw, row2val, idx2val)"""
    kid = {"stvk_tri3": 0, "bending_quad": 1}[kind]
    e2v = _u64(elem2vtx).reshape(-1, 3 if kid == 0 else 4)
    nb = row2idx.size - 1
    row2val = np.zeros((nb, 9))
    idx2val = np.zeros((idx2col.size, 9))
    dw = np.zeros((nb, 3))
    w = C.c_double(0)
    _chk(lib().orc_assemble_cloth(C.c_int(kid), _p(e2v), C.c_uint64(e2v.shape[0]), _p(_f64(vtx2xyz)), _p(_f64(vtx2disp)),
                                  C.c_double(p0), C.c_double(p1), _p(row2idx), C.c_uint64(nb), _p(idx2col), _p(row2val),
                                  _p(idx2val), _p(dw), C.byref(w)))
    return w.value, dw, row2val, idx2val


def elem2nz(convention, flavour, elem2vtx, n_node, row2idx, idx2col):
    """element -> nonzero slot map implied by the merge routines (diagonal slot = nnz + row for convention 0)"""
    e2v = _u64(elem2vtx).reshape(-1, n_node)
    out = np.zeros((e2v.shape[0], n_node, n_node), dtype=np.uint32)
    _chk(lib().orc_elem2nz(C.c_int(convention), C.c_int(flavour), _p(e2v), C.c_uint64(e2v.shape[0]), C.c_uint64(n_node),
                           _p(row2idx), C.c_uint64(row2idx.size - 1), _p(idx2col), C.c_uint64(idx2col.size), _p(out)))
    return out


# ---------------------------------------------------------------- BC / SpMV / solvers
def set_fixed_bc(val_dia, bc_flag, row2val, idx2val, row2idx, idx2col):
    """del-fem-cpu/src/sparse_square.rs:3-55 (in place)"""
    flag = np.ascontiguousarray(bc_flag, dtype=np.int32)
    n = flag.shape[1]
    lib().orc_set_fixed_bc(C.c_uint64(n), C.c_double(val_dia), _p(flag), _p(row2val), _p(idx2val), _p(row2idx),
                           C.c_uint64(flag.shape[0]), _p(idx2col), C.c_uint64(idx2col.size))


def set_fix_dof_to_rhs_vector(rhs, flag):
    """del-fem-cpu/src/sparse_square.rs:57-70 (in place)"""
    flag = np.ascontiguousarray(flag, dtype=np.int32)
    lib().orc_set_fix_dof_to_rhs_vector(C.c_uint64(flag.shape[1]), _p(rhs), _p(flag), C.c_uint64(flag.shape[0]))


def mult_vec(y, beta, alpha, x, row2idx, idx2col, idx2val, row2val):
    """del-fem-cpu/src/sparse_square.rs:419-447 (in place on y [nb,N])"""
    n = y.shape[1] if y.ndim == 2 else 1
    _chk(lib().orc_mult_vec(C.c_uint64(n), _p(y), C.c_double(beta), C.c_double(alpha), _p(_f64(x)), _p(row2idx),
                            C.c_uint64(row2idx.size - 1), _p(idx2col), _p(idx2val), _p(row2val)))


def ls_mult_mat(y, beta, alpha, x, row2idx, idx2col, idx2val, row2val):
    """del-fem-ls/src/sparse_square.rs:134-164"""
    nrow = row2idx.size - 1
    ndim = y.size // nrow
    lib().orc_ls_mult_mat(C.c_uint64(ndim), _p(y), C.c_double(beta), C.c_double(alpha), _p(_f64(x)), _p(row2idx),
                          C.c_uint64(nrow), _p(idx2col), _p(idx2val), _p(row2val))


def dot(a, b):
    """slice_of_array::dot (slice_of_array.rs:8-16): strict left fold"""
    a = _f64(a)
    n = a.shape[1] if a.ndim == 2 else 1
    return lib().orc_dot(C.c_uint64(n), _p(a), _p(_f64(b)), C.c_uint64(a.shape[0]))


EPS_CPU = float(np.finfo(np.float64).eps)  # T::epsilon()   sparse_square.rs:235
EPS_LS = float(np.float32(1.0e-20))        # 1.0e-20_f32.as_()  solver_sparse.rs:40


def conjugate_gradient(r, u, ap, p, tol, max_it, row2idx, idx2col, idx2val, row2val, flavour="cpu"):
    """del-fem-cpu/src/sparse_square.rs:218-261 (flavour 'cpu') / del-fem-ls/src/solver_sparse.rs:5-70 ('ls').
    in place on r,u,ap,p; returns the convergence history (ratios, k=0 excluded)"""
    n = r.shape[1] if r.ndim == 2 else 1
    hist = np.zeros(max_it + 1)
    nh = lib().orc_conjugate_gradient(C.c_uint64(n), _p(r), _p(u), _p(ap), _p(p), C.c_double(tol), C.c_uint64(max_it),
                                      _p(row2idx), C.c_uint64(row2idx.size - 1), _p(idx2col), _p(idx2val), _p(row2val),
                                      C.c_double(EPS_CPU if flavour == "cpu" else EPS_LS),
                                      C.c_int(0 if flavour == "cpu" else 1), _p(hist))
    if nh < 0:
        raise OraclePanic("assert!(pap >= 0)")
    return hist[:nh].copy()


class Ilu:
    """del-fem-cpu/src/sparse_ilu.rs Preconditioner. mode: 'ilu0' | 'full' | 'block_jacobi' | 'jacobi' | ('iluk', lev_fill)"""
    MODES = {"ilu0": 0, "full": 1, "block_jacobi": 2, "jacobi": 3}

    def __init__(self, mode, n, row2idx, idx2col):
        self.n = n
        code = 4 + int(mode[1]) if isinstance(mode, tuple) else self.MODES[mode]
        self.h = C.c_void_p(lib().orc_ilu_new(C.c_int(code), C.c_uint64(n), _p(row2idx),
                                              C.c_uint64(row2idx.size - 1), _p(idx2col)))
        self.num_blk = row2idx.size - 1

    def copy_value(self, row2idx, idx2col, idx2val, row2val):
        _chk(lib().orc_ilu_copy_value(self.h, _p(row2idx), _p(idx2col), _p(idx2val), _p(row2val)))

    def decompose(self):
        _chk(lib().orc_ilu_decompose(self.h))

    def solve(self, vec):
        lib().orc_ilu_solve(_p(vec), self.h)

    def row2val(self):
        ptr = lib().orc_ilu_row2val(self.h)
        return np.ctypeslib.as_array(ptr, shape=(self.num_blk, self.n * self.n)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ilu_free(self.h)
            self.h = None


def symbolic_iluk(row2idx, idx2col, lev_fill):
    """del-fem-cpu/src/sparse_ilu.rs:221-312 -> (row2idx, idx2col, row2idx_dia) of the level-of-fill pattern"""
    n = row2idx.size - 1
    r = np.zeros(n + 1, dtype=U64)
    d = np.zeros(n, dtype=U64)
    cap = max(int(idx2col.size) * 4, 16)
    while True:
        c = np.zeros(cap, dtype=U64)
        need = int(lib().orc_symbolic_iluk(_p(row2idx), _p(idx2col), C.c_uint64(n), C.c_uint64(lev_fill), _p(r), _p(c), C.c_uint64(cap), _p(d)))
        if need <= cap:
            return r, c[:need].copy(), d
        cap = need


def preconditioned_conjugate_gradient(r, x, pr, p, tol, max_it, row2idx, idx2col, idx2val, row2val, ilu, flavour="cpu"):
    """del-fem-cpu/src/sparse_square.rs:264-347 generalised to N. history = absolute norms incl. k=0 (+1 trailing)"""
    n = r.shape[1] if r.ndim == 2 else 1
    hist = np.zeros(max_it + 3)
    nh = lib().orc_preconditioned_conjugate_gradient(
        C.c_uint64(n), _p(r), _p(x), _p(pr), _p(p), C.c_double(tol), C.c_uint64(max_it), _p(row2idx),
        C.c_uint64(row2idx.size - 1), _p(idx2col), _p(idx2val), _p(row2val), ilu.h,
        C.c_double(EPS_CPU if flavour == "cpu" else EPS_LS), _p(hist))
    return hist[:nh].copy()


def pcg_jacobi3_mt(nthreads, r, x, pr, p, tol, max_it, row2idx, idx2col, idx2val, row2val, eps_sq=0.0):
    """BASELINE-TIMING AID: row-parallel (std::thread) Jacobi-PCG, 3x3 blocks, A12 history semantics"""
    hist = np.zeros(max_it + 2)
    nh = lib().orc_pcg_jacobi3_mt(C.c_int(nthreads), _p(r), _p(x), _p(pr), _p(p), C.c_double(tol), C.c_uint64(max_it), _p(row2idx),
                                  C.c_uint64(row2idx.size - 1), _p(idx2col), _p(idx2val), _p(row2val), C.c_double(eps_sq), _p(hist))
    return hist[:nh]


def mult_square_matrices(r0, c0, v0, d0, r1, c1, v1, d1):
    """del-fem-ls/src/sparse_matrix_multiplication.rs:109-188: C = M0 * M1 (scalar csrdia) -> (row2idx, idx2col, idx2val, row2val)"""
    n = r0.size - 1
    v0, d0, v1, d1 = (_f64(a).reshape(-1) for a in (v0, d0, v1, d1))
    r = np.zeros(n + 1, dtype=U64)
    nnz = int(lib().orc_mult_square_matrices(C.c_uint64(n), _p(r0), _p(c0), _p(v0), _p(d0), _p(r1), _p(c1), _p(v1), _p(d1), _p(r),
                                             None, None, None))
    c = np.zeros(nnz, dtype=U64)
    iv = np.zeros(nnz)
    rv = np.zeros(n)
    lib().orc_mult_square_matrices(C.c_uint64(n), _p(r0), _p(c0), _p(v0), _p(d0), _p(r1), _p(c1), _p(v1), _p(d1), _p(r), _p(c), _p(iv),
                                   _p(rv))
    return r, c, iv, rv
