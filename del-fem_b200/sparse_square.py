"""Mirror of del-fem-cpu/src/sparse_square.rs (and del-fem-ls/src/{sparse_square,solver_sparse}.rs) on the device.

Same names, argument order and error behaviour as the reference; vectors are `core.Vec` (device) instead of slices.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import DfbError, check, lib
from .core import ELEM, Ctx, Mesh, Pattern, Plan, Vec, _c, _ptr

EPS_CPU = float(np.finfo(np.float64).eps)  # T::epsilon()            sparse_square.rs:235
EPS_LS = float(np.float32(1.0e-20))        # 1.0e-20_f32.as_()       solver_sparse.rs:40


class Matrix:
    """sparse_square::Matrix<[f64; NN]>  (del-fem-cpu/src/sparse_square.rs:75-214)

    `row2idx`/`idx2col` live in `self.pattern`; `idx2val`/`row2val` are device buffers (download() to read)."""

    def __init__(self, ctx: Ctx, pattern: Pattern, blk_dim: int, _borrow=None):
        self.ctx, self.pattern, self.blk_dim = ctx, pattern, int(blk_dim)
        self.num_blk = pattern.num_blk
        self.halo = None
        self._owned = _borrow is None
        if _borrow is not None:  # a handle owned by another object (linearsystem.Solver.sparse)
            self.h = C.c_void_p(_borrow)
            return
        h = C.c_void_p()
        check(lib().dfb_bsr_create(ctx.h, pattern.h, self.blk_dim, C.byref(h)))
        self.h = h

    @classmethod
    def from_vtx2vtx(cls, vtx2idx, idx2vtx, blk_dim, ctx: Ctx = None, convention=0):
        """Matrix::from_vtx2vtx (:117)"""
        ctx = ctx or Ctx.default()
        return cls(ctx, Pattern(ctx, vtx2idx, idx2vtx, convention), blk_dim)

    def set_zero(self):
        """Matrix::set_zero (:174)"""
        check(lib().dfb_bsr_set_zero(self.h))

    def upload(self, row2val=None, idx2val=None):
        nn = self.blk_dim ** 2
        r = None if row2val is None else _c(row2val, np.float64)
        v = None if idx2val is None else _c(idx2val, np.float64)
        if r is not None and r.size != self.num_blk * nn:
            raise DfbError(1, "Matrix.upload: row2val size mismatch")
        if v is not None and v.size != self.pattern.nnz * nn:
            raise DfbError(1, "Matrix.upload: idx2val size mismatch (assert_eq!(idx2val.len(), idx2col.len()))")
        check(lib().dfb_bsr_upload(self.h, _ptr(r), _ptr(v)))

    def download(self):
        """-> (row2val [num_blk, NN] or None, idx2val [nnz, NN])"""
        nn = self.blk_dim ** 2
        v = np.empty((self.pattern.nnz, nn))
        r = np.empty((self.num_blk, nn)) if self.pattern.convention == 0 else None
        check(lib().dfb_bsr_download(self.h, _ptr(r), _ptr(v)))
        return r, v

    def set_fixed_dof(self, val_dia, blk2isfix):
        """Matrix::set_fixed_dof / set_fixed_bc (:129 / :3-55)"""
        f = _c(blk2isfix, np.int32)
        n_all = self.num_blk + (self.halo.n_ghost if self.halo is not None else 0)
        if f.size != n_all * self.blk_dim:
            raise DfbError(1, "set_fixed_dof: assert_eq!(bc_flag.len(), row2val.len())")
        check(lib().dfb_bsr_set_fixed_dof(self.h, float(val_dia), _ptr(f)))

    def set_fixed_dof_dev(self, val_dia, flag):
        """set_fixed_dof with device-resident flags (core.BcFlag): no host traffic"""
        check(lib().dfb_bsr_set_fixed_dof_dev(self.h, float(val_dia), flag.h))

    def mult_vec(self, y_vec: Vec, beta, alpha, x_vec: Vec):
        """Matrix::mult_vec: y <- alpha*A*x + beta*y (:143-171)"""
        check(lib().dfb_bsr_mult_vec(self.h, y_vec.h, float(beta), float(alpha), x_vec.h))

    def mult_square_matrices(self, other: "Matrix") -> "Matrix":
        """del-fem-ls mult_square_matrices (sparse_matrix_multiplication.rs:109-188): self * other for scalar matrices of the
        separate-diagonal convention; returns a new Matrix on a new Pattern (columns in the reference's discovery order)"""
        hp, hm = C.c_void_p(), C.c_void_p()
        check(lib().dfb_bsr_mult_square_matrices(self.h, other.h, C.byref(hp), C.byref(hm)))
        out = Matrix.__new__(Matrix)
        out.ctx, out.pattern, out.blk_dim = self.ctx, Pattern(self.ctx, _handle=hp), 1
        out.num_blk, out.h, out.halo = out.pattern.num_blk, hm, None
        return out

    def mult_mat(self, y_mat: Vec, beta, alpha, x_mat: Vec):
        """del-fem-ls mult_mat (del-fem-ls/src/sparse_square.rs:134-164): scalar matrix times num_dim interleaved components"""
        check(lib().dfb_bsr_mult_mat(self.h, y_mat.h, float(beta), float(alpha), x_mat.h))

    def merge_for_array_blk(self, emat, node2vtx, col2idx=None, flavour=0):
        """Matrix::merge_for_array_blk (:180-214). `col2idx` is accepted for signature parity and ignored."""
        e = _c(emat, np.float64)
        n = _c(node2vtx, np.uint64)
        if e.size != n.size * n.size * self.blk_dim ** 2:
            raise DfbError(1, "merge_for_array_blk: emat shape mismatch")
        check(lib().dfb_bsr_merge_one(self.h, _ptr(e), _ptr(n), n.size, flavour))

    def pack(self):
        """build the SpMV mirror of the current values now (otherwise the first SpMV after a value change does it)"""
        check(lib().dfb_bsr_pack(self.h))

    def add_to_diagonal(self, scale, v: Vec):
        check(lib().dfb_bsr_add_to_diagonal(self.h, float(scale), v.h))

    def add_scaled(self, alpha, other: "Matrix"):
        """values(self) += alpha * values(other); both matrices must have been created on the same Pattern"""
        check(lib().dfb_bsr_add_scaled(self.h, float(alpha), other.h))

    def set_halo(self, halo, int_begin, int_end):
        check(lib().dfb_bsr_set_halo(self.h, halo.h if halo is not None else None, int_begin, int_end))
        self.halo = halo

    # ---- batched assembly (replaces the reference's per-element loop)
    def assemble(self, plan: Plan, kind, mesh: Mesh, p0=1.0, p1=0.0):
        k = ELEM[kind] if isinstance(kind, str) else kind
        check(lib().dfb_assemble(plan.h, k, mesh.h, float(p0), float(p1), self.h))

    def assemble_from_emat(self, plan: Plan, emat_dev):
        check(lib().dfb_assemble_from_emat(plan.h, emat_dev, self.h))

    def as_ref(self):
        return self

    def as_ref_mut(self):
        return self

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owned", True):
            lib().dfb_bsr_destroy(self.h)
        self.h = None


def set_fix_dof_to_rhs_vector(blk2rhs: Vec, blk2isfix):
    """sparse_square::set_fix_dof_to_rhs_vector (:57-70)"""
    blk2rhs.set_fixed_dof_zero(blk2isfix)


_HIST_FIRST_GUESS = 4096


def _history_from_sq(ctx, n_hist, pcg):
    """history longer than the first-guess buffer: rebuild it from the library's host copy of ||r_k||^2 (k = 0..iterations)"""
    n = C.c_uint64()
    check(lib().dfb_ctx_last_history_sq(ctx.h, None, 0, C.byref(n)))
    sq = np.zeros(n.value)
    check(lib().dfb_ctx_last_history_sq(ctx.h, _ptr(sq), sq.size, C.byref(n)))
    if pcg:
        h = np.sqrt(sq)
        return np.concatenate([h, h[-1:]])[:n_hist] if n_hist == sq.size + 1 else h[:n_hist]
    return np.sqrt(sq[1:] * (1.0 / sq[0]))[:n_hist]


def conjugate_gradient(r_vec: Vec, u_vec: Vec, ap_vec: Vec, p_vec: Vec, conv_ratio_tol, max_iteration, mat: Matrix,
                       flavour="cpu"):
    """sparse_square::conjugate_gradient (:218-261; flavour 'ls' = del-fem-ls/src/solver_sparse.rs:5-70).
    Returns the convergence history (||r_k||/||r_0||, k>=1) as a numpy array."""
    hist = np.zeros(min(int(max_iteration) + 2, _HIST_FIRST_GUESS))
    n = C.c_uint64()
    check(lib().dfb_cg(mat.h, r_vec.h, u_vec.h, ap_vec.h, p_vec.h, float(conv_ratio_tol), int(max_iteration),
                       EPS_CPU if flavour == "cpu" else EPS_LS, 0 if flavour == "cpu" else 1, _ptr(hist), hist.size,
                       C.byref(n)))
    if n.value > hist.size:
        return _history_from_sq(mat.ctx, n.value, False)
    return hist[:n.value].copy()


def preconditioned_conjugate_gradient(r_vec: Vec, x_vec: Vec, pr_vec: Vec, p_vec: Vec, conv_ratio_tol, max_nitr,
                                      mat: Matrix, ilu, flavour="cpu"):
    """sparse_square::preconditioned_conjugate_gradient (:264-347) generalised to N with a (block-)Jacobi
    `sparse_ilu.Preconditioner`. History = absolute ||r_k|| incl. k=0 (+1 trailing entry when max_nitr is exhausted)."""
    hist = np.zeros(min(int(max_nitr) + 3, _HIST_FIRST_GUESS))
    n = C.c_uint64()
    check(lib().dfb_pcg(mat.h, ilu.h, r_vec.h, x_vec.h, pr_vec.h, p_vec.h, float(conv_ratio_tol), int(max_nitr),
                        EPS_CPU if flavour == "cpu" else EPS_LS, _ptr(hist), hist.size, C.byref(n)))
    if n.value > hist.size:
        return _history_from_sq(mat.ctx, n.value, True)
    return hist[:n.value].copy()
