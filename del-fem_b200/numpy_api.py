"""Flat-numpy-array surface with the names and argument order of `del-fem-numpy` (PyO3 module + Python wrappers), evaluated on
the device.  Like the reference functions these ADD the assembled values into the caller's arrays in place.

  del-fem-numpy/src/laplacian.rs:28   merge_hessian_mesh_laplacian_on_trimesh3   (csrdia: row2val + idx2val)
  del-fem-numpy/src/laplacian.rs:71   merge_laplace_to_bsr_for_meshtri2          (blkcsr: diagonal inside the pattern)
  del-fem-numpy/src/laplacian.rs:106  merge_laplace_to_bsr_for_meshtri3
  del-fem-numpy/src/linear_solid.rs:17 merge_linear_solid_to_bsr_for_meshtri2    (lambda = myu = 1 hard-wired upstream)
  del-fem-numpy/src/block_sparse.rs:22,60,78  block_sparse_apply_bc, block_sparse_set_fixed_bc_to_rhs_vector, conjugate_gradient
  del_fem_numpy/LaplacianMatrix.py:4  from_uniform_mesh -> scipy.sparse.csr_matrix
The reference uses f32 arrays for most of these; here any float dtype is accepted, arithmetic is fp64 on the device, results are
cast back to the array's dtype.
"""
from __future__ import annotations

import numpy as np

from .core import Ctx, Mesh, Pattern, Plan, Vec
from .sparse_square import Matrix, conjugate_gradient as _cg


def _assemble(kind, elem2vtx, vtx2xyz, row2idx, idx2col, convention, p0, p1, blk_dim, ctx):
    ctx = ctx or Ctx.default()
    mesh = Mesh(ctx, np.ascontiguousarray(elem2vtx, dtype=np.uint64), np.ascontiguousarray(vtx2xyz, dtype=np.float64))
    pat = Pattern(ctx, np.ascontiguousarray(row2idx, dtype=np.uint64), np.ascontiguousarray(idx2col, dtype=np.uint64), convention)
    plan = Plan(ctx, pat, mesh, 1)  # merge::csrdia / merge::blkcsr test the diagonal on vertex ids (merge.rs:31)
    mat = Matrix(ctx, pat, blk_dim)
    mat.assemble(plan, kind, mesh, p0, p1)
    return mat.download()


def merge_hessian_mesh_laplacian_on_trimesh3(tri2vtx, vtx2xyz, row2idx, idx2col, row2val, idx2val, ctx=None):
    rv, iv = _assemble("cot_laplace_tri3", tri2vtx, vtx2xyz, row2idx, idx2col, 0, 0.0, 0.0, 1, ctx)
    row2val += rv[:, 0].astype(row2val.dtype)
    idx2val += iv[:, 0].astype(idx2val.dtype)


def merge_laplace_to_bsr_for_meshtri2(tri2vtx, vtx2xy, row2idx, idx2col, idx2val, ctx=None):
    _, iv = _assemble("laplace_tri2", tri2vtx, vtx2xy, row2idx, idx2col, 1, 1.0, 0.0, 1, ctx)
    idx2val += iv[:, 0].astype(idx2val.dtype)


def merge_laplace_to_bsr_for_meshtri3(tri2vtx, vtx2xyz, row2idx, idx2col, idx2val, ctx=None):
    _, iv = _assemble("cot_laplace_tri3", tri2vtx, vtx2xyz, row2idx, idx2col, 1, 0.0, 0.0, 1, ctx)
    idx2val += iv[:, 0].astype(idx2val.dtype)


def merge_linear_solid_to_bsr_for_meshtri2(tri2vtx, vtx2xy, row2idx, idx2col, idx2val, ctx=None):
    """idx2val has shape [nnz, 2, 2]; like the reference it receives the flat column-major block (merge.rs:78-80)"""
    assert idx2val.shape == (len(idx2col), 2, 2)
    _, iv = _assemble("solid_linear_tri2", tri2vtx, vtx2xy, row2idx, idx2col, 1, 1.0, 1.0, 2, ctx)
    idx2val += iv.reshape(-1, 2, 2).astype(idx2val.dtype)


def block_sparse_apply_bc(val_dia, blk2isfix, row2val, idx2val, row2idx, idx2col, ctx=None):
    """set_fixed_bc on flat [num_blk, N*N] / [num_idx, N*N] arrays (any N in 1..4; upstream implements only N = 4)"""
    ctx = ctx or Ctx.default()
    n = blk2isfix.shape[1]
    assert row2val.shape == (blk2isfix.shape[0], n * n) and idx2val.shape == (len(idx2col), n * n)
    mat = Matrix.from_vtx2vtx(np.ascontiguousarray(row2idx, dtype=np.uint64), np.ascontiguousarray(idx2col, dtype=np.uint64), n, ctx=ctx)
    mat.upload(row2val.astype(np.float64), idx2val.astype(np.float64))
    mat.set_fixed_dof(val_dia, blk2isfix)
    rv, iv = mat.download()
    row2val[...] = rv.astype(row2val.dtype)
    idx2val[...] = iv.astype(idx2val.dtype)


def block_sparse_set_fixed_bc_to_rhs_vector(blk2isfix, rhs):
    rhs[np.asarray(blk2isfix) != 0] = 0


def conjugate_gradient(r_vec, u_vec, p_vec, ap_vec, row2idx, idx2col, idx2val, row2val, conv_ratio_tol=1.0e-5, max_iteration=1000,
                       ctx=None):
    """block_sparse.rs:78-120 (tol 1e-5, 1000 iterations upstream); arrays [num_blk, N] are updated in place, returns the history"""
    ctx = ctx or Ctx.default()
    n = r_vec.shape[1]
    mat = Matrix.from_vtx2vtx(np.ascontiguousarray(row2idx, dtype=np.uint64), np.ascontiguousarray(idx2col, dtype=np.uint64), n, ctx=ctx)
    mat.upload(np.asarray(row2val, dtype=np.float64), np.asarray(idx2val, dtype=np.float64))
    r, u, ap, p = (Vec.from_numpy(ctx, np.asarray(a, dtype=np.float64)) for a in (r_vec, u_vec, ap_vec, p_vec))
    hist = _cg(r, u, ap, p, conv_ratio_tol, max_iteration, mat)
    for dst, src in ((r_vec, r), (u_vec, u), (ap_vec, ap), (p_vec, p)):
        dst[...] = src.download().astype(dst.dtype)
    return hist.astype(r_vec.dtype)


class LaplacianMatrix:
    """del_fem_numpy/LaplacianMatrix.py:4-21"""

    @staticmethod
    def from_uniform_mesh(elem2vtx, vtx2xyz, ctx=None):
        import scipy.sparse
        ctx = ctx or Ctx.default()
        xyz = np.ascontiguousarray(vtx2xyz, dtype=np.float64)
        mesh = Mesh(ctx, np.ascontiguousarray(elem2vtx, dtype=np.uint64), xyz)
        pat = Pattern.from_uniform_mesh(ctx, mesh, True)  # TriMesh.vtx2vtx(elem2vtx, nv, True)
        row2idx, idx2col = pat.download()
        idx2val = np.zeros(idx2col.shape, dtype=np.float64)
        if xyz.shape[1] == 2:
            merge_laplace_to_bsr_for_meshtri2(elem2vtx, xyz, row2idx, idx2col, idx2val, ctx)
        else:
            merge_laplace_to_bsr_for_meshtri3(elem2vtx, xyz, row2idx, idx2col, idx2val, ctx)
        return scipy.sparse.csr_matrix((idx2val, idx2col.astype(np.int64), row2idx.astype(np.int64)))


class MassMatrix:
    """del_fem_numpy/MassMatrix.py:3-10 — diagonal (lumped) mass: del_msh_numpy TriMesh.vtx2area (area / 3 per corner) repeated ndim
    times, as a scipy dia_matrix.  Evaluated by dfb_lumped_mass (rho = 1); also accepts tet meshes (volume / 4 per corner)."""

    @staticmethod
    def from_uniform_mesh(elem2vtx, vtx2xyz, ndim=1, ctx=None):
        import scipy.sparse
        from . import elements
        ctx = ctx or Ctx.default()
        xyz = np.ascontiguousarray(vtx2xyz, dtype=np.float64)
        mesh = Mesh(ctx, np.ascontiguousarray(elem2vtx, dtype=np.uint64), xyz)
        pat = Pattern.from_uniform_mesh(ctx, mesh, False)
        plan = Plan(ctx, pat, mesh, 0)
        out = Vec(ctx, xyz.shape[0], 1)
        elements.lumped_mass(plan, mesh, 1.0, out)
        vtx2area = out.download().ravel()
        n = xyz.shape[0] * ndim
        return scipy.sparse.dia_matrix((vtx2area.repeat(ndim).reshape(1, -1), np.array([0])), shape=(n, n))


class LinearSolid:
    """del_fem_numpy/LinearSolid.py:3-12 (upstream imports a name that does not exist, `merge_linear_solid_to_csr_for_meshtri2`;
    the function that exists is `merge_linear_solid_to_bsr_for_meshtri2`, linear_solid.rs:17 — that is the one used here)"""

    @staticmethod
    def stiffness_matrix_from_uniform_mesh(tri2vtx, vtx2xy, ctx=None):
        import scipy.sparse
        ctx = ctx or Ctx.default()
        xy = np.ascontiguousarray(vtx2xy, dtype=np.float64)
        mesh = Mesh(ctx, np.ascontiguousarray(tri2vtx, dtype=np.uint64), xy)
        row2idx, idx2col = Pattern.from_uniform_mesh(ctx, mesh, True).download()
        idx2val = np.zeros((idx2col.shape[0], 2, 2), dtype=np.float64)
        merge_linear_solid_to_bsr_for_meshtri2(tri2vtx, xy, row2idx, idx2col, idx2val, ctx)
        # blocks are column-major (a + 2 b, solid_linear_tri2.rs:11-19); scipy's bsr blocks are row-major [a][b]
        blocks = idx2val.reshape(-1, 2, 2).transpose(0, 2, 1)
        return scipy.sparse.bsr_matrix((blocks, idx2col.astype(np.int64), row2idx.astype(np.int64)))
