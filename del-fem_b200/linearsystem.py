"""Mirror of del-fem-ls/src/linearsystem.rs:8-112 (`Solver` façade) over device objects, generalised to block dims."""
from __future__ import annotations

import numpy as np

from . import sparse_ilu, sparse_square
from .core import Ctx, Pattern, Vec


class Solver:
    def __init__(self, ctx: Ctx = None, blk_dim: int = 1, precond="jacobi"):
        """Solver::new (:33): tol 1e-5, 100 iterations. precond: "jacobi" (fast GPU path, default), "block_jacobi", or "ilu0"
        (what the reference's Solver::initialize sets up, :53)"""
        self.ctx = ctx or Ctx.default()
        self.blk_dim = blk_dim
        self.sparse = None
        self.ilu = sparse_ilu.Preconditioner(precond)
        self.merge_buffer = None  # kept for signature parity; the device merge needs no scratch
        self.r_vec = self.u_vec = self.ap_vec = self.p_vec = None
        self.conv = np.zeros(0)
        self.conv_ratio_tol = float(np.float32(1.0e-5))
        self.max_num_iteration = 100

    def initialize(self, colind, rowptr):
        """Solver::initialize (:48). NB the reference's parameter names are swapped relative to their meaning:
        the FIRST argument is the row-pointer array, the second the column indices."""
        pat = Pattern(self.ctx, colind, rowptr, 0)
        self.sparse = sparse_square.Matrix(self.ctx, pat, self.blk_dim)
        nblk = pat.num_blk
        self.r_vec = Vec(self.ctx, nblk, self.blk_dim)
        self.u_vec = Vec(self.ctx, nblk, self.blk_dim)
        self.ap_vec = Vec(self.ctx, nblk, self.blk_dim)
        self.p_vec = Vec(self.ctx, nblk, self.blk_dim)
        self.ilu.initialize(self.sparse)

    def begin_merge(self):
        """Solver::begin_merge (:60)"""
        self.sparse.set_zero()
        self.r_vec.set_zero()

    def end_merge(self):
        """Solver::end_merge (:67)"""
        sparse_ilu.copy_value(self.ilu, self.sparse)
        sparse_ilu.decompose(self.ilu)

    def solve_cg(self):
        """Solver::solve_cg (:73)"""
        self.conv = sparse_square.conjugate_gradient(self.r_vec, self.u_vec, self.ap_vec, self.p_vec, self.conv_ratio_tol,
                                                     self.max_num_iteration, self.sparse, flavour="ls")

    def solve_pcg(self):
        """Solver::solve_pcg (:85)"""
        self.conv = sparse_square.preconditioned_conjugate_gradient(self.r_vec, self.u_vec, self.ap_vec, self.p_vec,
                                                                    self.conv_ratio_tol, self.max_num_iteration,
                                                                    self.sparse, self.ilu, flavour="ls")
