"""Mirror of del-fem-ls/src/linearsystem.rs:8-112 (`Solver`) over the C ABI's `dfb_solver_*` entry points, one method per
reference method, generalised from the reference's scalar `T` to blk_dim 1..4.

The fields are the reference's: `sparse`, `ilu`, `r_vec`, `u_vec`, `ap_vec`, `p_vec` (device objects BORROWED from the solver
handle: merge / apply BCs on `solver.sparse` and `solver.r_vec` between `begin_merge` and `end_merge` as reference code does),
`conv`, `conv_ratio_tol`, `max_num_iteration`, `merge_buffer` (kept for signature parity; the device merge needs no scratch).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import sparse_ilu, sparse_square
from ._lib import check, lib
from .core import PRECOND, Ctx, Pattern, Vec, _c, _ptr


class Solver:
    def __init__(self, ctx: Ctx = None, blk_dim: int = 1, _handle=None):
        """Solver::new (:33): conv_ratio_tol = 1e-5 (f32), max_num_iteration = 100, ILU(0) set up by `initialize` (:53)"""
        self.ctx = ctx or Ctx.default()
        self.blk_dim = int(blk_dim)
        if _handle is None:
            h = C.c_void_p()
            check(lib().dfb_solver_new(self.ctx.h, self.blk_dim, C.byref(h)))
            _handle = h
        self.h = _handle
        self.merge_buffer = None
        self.sparse = self.ilu = self.r_vec = self.u_vec = self.ap_vec = self.p_vec = None
        self.conv = np.zeros(0)
        self._bind()

    # ---- scalar fields live in the handle
    @property
    def conv_ratio_tol(self):
        t, n = C.c_double(), C.c_uint64()
        check(lib().dfb_solver_get_params(self.h, C.byref(t), C.byref(n)))
        return t.value

    @conv_ratio_tol.setter
    def conv_ratio_tol(self, v):
        check(lib().dfb_solver_set_params(self.h, float(v), self.max_num_iteration))

    @property
    def max_num_iteration(self):
        t, n = C.c_double(), C.c_uint64()
        check(lib().dfb_solver_get_params(self.h, C.byref(t), C.byref(n)))
        return n.value

    @max_num_iteration.setter
    def max_num_iteration(self, v):
        check(lib().dfb_solver_set_params(self.h, self.conv_ratio_tol, int(v)))

    def _bind(self):
        """wrap the solver's borrowed handles (None before `initialize`)"""
        L = lib()
        hs = L.dfb_solver_sparse(self.h)
        if not hs:
            return
        pat = Pattern(self.ctx, _handle=C.c_void_p(L.dfb_bsr_pattern(hs)), _owned=False)
        self.sparse = sparse_square.Matrix(self.ctx, pat, self.blk_dim, _borrow=hs)
        n = pat.num_blk
        self.r_vec = Vec(self.ctx, n, self.blk_dim, _borrow=L.dfb_solver_r_vec(self.h))
        self.u_vec = Vec(self.ctx, n, self.blk_dim, _borrow=L.dfb_solver_u_vec(self.h))
        self.ap_vec = Vec(self.ctx, n, self.blk_dim, _borrow=L.dfb_solver_ap_vec(self.h))
        self.p_vec = Vec(self.ctx, n, self.blk_dim, _borrow=L.dfb_solver_p_vec(self.h))
        self.ilu = sparse_ilu.Preconditioner("ilu0", _borrow=L.dfb_solver_ilu(self.h), _ctx=self.ctx)

    def set_preconditioner(self, kind="ilu0", lev_fill=0):
        """the alternatives the reference keeps commented out at :50,:54-56 (`initialize_iluk`), plus the derived Jacobi kinds"""
        check(lib().dfb_solver_set_preconditioner(self.h, PRECOND[kind] if isinstance(kind, str) else int(kind), int(lev_fill)))
        self._bind()

    def initialize(self, colind, rowptr):
        """Solver::initialize (:48). NB the reference's parameter names are swapped relative to their meaning:
        the FIRST argument is the row-pointer array, the second the column indices."""
        a, b = _c(colind, np.uint64), _c(rowptr, np.uint64)
        check(lib().dfb_solver_initialize(self.h, _ptr(a), a.size, _ptr(b), b.size))
        self._bind()

    def begin_merge(self):
        """Solver::begin_merge (:60)"""
        check(lib().dfb_solver_begin_merge(self.h))

    def end_merge(self):
        """Solver::end_merge (:67): copy_value + decompose"""
        check(lib().dfb_solver_end_merge(self.h))

    def _fetch_conv(self):
        n = C.c_uint64()
        check(lib().dfb_solver_conv(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        check(lib().dfb_solver_conv(self.h, _ptr(out), out.size, C.byref(n)))
        self.conv = out

    def solve_cg(self):
        """Solver::solve_cg (:73)"""
        check(lib().dfb_solver_solve_cg(self.h))
        self._fetch_conv()

    def solve_pcg(self):
        """Solver::solve_pcg (:85)"""
        check(lib().dfb_solver_solve_pcg(self.h))
        self._fetch_conv()

    def clone(self):
        """Solver::clone (:96-111): deep copy; `max_num_iteration` is reset to 100 as the reference does (:109)"""
        h = C.c_void_p()
        check(lib().dfb_solver_clone(self.h, C.byref(h)))
        other = Solver(self.ctx, self.blk_dim, _handle=h)
        other._fetch_conv()
        return other

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.sparse = self.ilu = self.r_vec = self.u_vec = self.ap_vec = self.p_vec = None
            lib().dfb_solver_destroy(self.h)
        self.h = None
