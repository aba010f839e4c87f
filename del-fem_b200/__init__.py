"""del_fem_b200 — host-side mirror (Python/ctypes) of the reference interface for the hot path
element kernels -> merge -> BC -> (P)CG, over libdelfem_b200.so (hand-written sm_100a CUDA, fp64).

Module names follow the reference crates: `sparse_square`, `sparse_ilu`, `linearsystem` (del-fem-cpu / del-fem-ls),
`elements` (laplace_tri2, solid_linear_tri2, ... and the derived tet kernels).  There is no CPU fallback: importing
succeeds without the library, but the first call raises ImportError / DfbError if it (or a GPU) is missing.
"""
from . import core, elements, linearsystem, sparse_ilu, sparse_square  # noqa: F401
from ._lib import LIB_PATH, STATUS, DfbError, build, header_symbols, lib  # noqa: F401
from .core import ELEM, PRECOND, BcFlag, Ctx, Halo, Mesh, Pattern, Plan, Vec  # noqa: F401
from .sparse_square import Matrix, conjugate_gradient, preconditioned_conjugate_gradient  # noqa: F401

__version__ = "0.1.0"
