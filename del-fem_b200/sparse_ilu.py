"""Mirror of del-fem-cpu/src/sparse_ilu.rs: ILU(0) (`initialize_ilu0`), ILU(k) (`initialize_iluk`, level of fill), the same code
path with an empty off-diagonal pattern (= block-Jacobi) and point Jacobi (DERIVED)."""
from __future__ import annotations

import ctypes as C

from ._lib import check, lib
from .core import PRECOND, Ctx, Vec
from .sparse_square import Matrix


class Preconditioner:
    def __init__(self, kind="block_jacobi", _borrow=None, _ctx=None):
        self.kind = PRECOND[kind] if isinstance(kind, str) else kind
        self.h = None
        self.ctx = None
        self._owned = _borrow is None
        if _borrow is not None:  # a handle owned by another object (linearsystem.Solver.ilu)
            self.h, self.ctx = C.c_void_p(_borrow), _ctx

    def initialize(self, a: Matrix):
        """initialize_ilu0-like symbolic phase (:28): allocates the inverse-diagonal storage"""
        self.ctx = a.ctx
        h = C.c_void_p()
        check(lib().dfb_precond_create(a.ctx.h, self.kind, a.h, C.byref(h)))
        self.h = h

    def initialize_ilu0(self, a: Matrix):
        """Preconditioner::initialize_ilu0 (:28-56)"""
        self.kind = PRECOND["ilu0"]
        self.initialize(a)

    def initialize_iluk(self, a: Matrix, lev_fill: int):
        """Preconditioner::initialize_iluk (:58-64): level-of-fill pattern from symbolic_iluk (:221-312)"""
        self.kind = PRECOND["ilu0"]
        self.ctx = a.ctx
        h = C.c_void_p()
        check(lib().dfb_precond_create_iluk(a.ctx.h, a.h, int(lev_fill), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None) and getattr(self, "_owned", True):
            lib().dfb_precond_destroy(self.h)
        self.h = None


def copy_value(ilu: Preconditioner, a: Matrix):
    """copy_value (:87-125) — fused with decompose on the device (the copy is never materialised)"""
    ilu._src = a


def decompose(ilu: Preconditioner):
    """decompose (:127-181): stores D_i^-1 (raises DfbError(SINGULAR_DIAG) where the reference unwraps None)"""
    check(lib().dfb_precond_update(ilu.h, ilu._src.h))


def solve_preconditioning_vec(vec: Vec, ilu: Preconditioner):
    """solve_preconditioning_vec (:183-219), in place"""
    check(lib().dfb_precond_apply(ilu.h, vec.h))
