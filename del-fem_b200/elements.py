"""Element kernels with the reference's call signatures (one element, host arrays in / nested array out), evaluated on the
device by the same code the batched assembly uses, plus the batched entry points."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .core import ELEM, ELEM_INFO, Ctx, Mesh, Vec, _c, _ptr


def _one(kind, pts, p0, p1, ctx=None):
    ctx = ctx or Ctx.default()
    n_node, ndim, n = ELEM_INFO[kind]
    xyz = _c(np.stack([np.asarray(p, dtype=np.float64) for p in pts]), np.float64)
    assert xyz.shape == (n_node, ndim)
    out = np.empty((n_node, n_node, n * n))
    check(lib().dfb_emat_one(ctx.h, kind, _ptr(xyz), float(p0), float(p1), _ptr(out)))
    return out


def laplace_tri2_ddw_(alpha, p0, p1, p2, ctx=None):
    """laplace_tri2::ddw_ (del-fem-cpu/src/laplace_tri2.rs:3) -> [3][3][1]"""
    return _one(0, (p0, p1, p2), alpha, 0.0, ctx)


def tri3_emat_cotangent_laplacian(p0, p1, p2, ctx=None):
    """del_geo_core::tri3::emat_cotangent_laplacian (call site laplace_tri3.rs:18) -> [3][3][1]"""
    return _one(1, (p0, p1, p2), 0.0, 0.0, ctx)


def solid_linear_tri2_emat_tri2(lam, myu, p0, p1, p2, ctx=None):
    """solid_linear_tri2::emat_tri2 (del-fem-cpu/src/solid_linear_tri2.rs:23) -> [3][3][4]"""
    return _one(2, (p0, p1, p2), lam, myu, ctx)


def laplace_tet_ddw_(alpha, p0, p1, p2, p3, ctx=None):
    """DERIVED tet Laplacian -> [4][4][1]"""
    return _one(3, (p0, p1, p2, p3), alpha, 0.0, ctx)


def solid_linear_tet_emat_tet(lam, myu, p0, p1, p2, p3, ctx=None):
    """DERIVED tet linear solid -> [4][4][9]"""
    return _one(4, (p0, p1, p2, p3), lam, myu, ctx)


def emat_batch(ctx: Ctx, kind, mesh: Mesh, p0=1.0, p1=0.0, disp: Vec = None, want_evec=False, want_w=False):
    """batched element kernels; returns host arrays (emat [ne,n,n,NN], evec [ne,n,N] | None, w [ne] | None)"""
    from .core import device_free, device_malloc, memcpy_d2h
    k = ELEM[kind] if isinstance(kind, str) else kind
    n_node, _, n = ELEM_INFO[k]
    ne = mesh.num_elem
    emat = np.empty((ne, n_node, n_node, n * n))
    d_e = device_malloc(ctx, emat.nbytes)
    evec = np.empty((ne, n_node, n)) if want_evec else None
    w = np.empty(ne) if want_w else None
    d_v = device_malloc(ctx, evec.nbytes) if want_evec else None
    d_w = device_malloc(ctx, w.nbytes) if want_w else None
    try:
        check(lib().dfb_emat_batch(ctx.h, k, mesh.h, float(p0), float(p1), disp.h if disp is not None else None, d_e, d_v, d_w))
        memcpy_d2h(ctx, emat, d_e)
        if want_evec:
            memcpy_d2h(ctx, evec, d_v)
        if want_w:
            memcpy_d2h(ctx, w, d_w)
    finally:
        device_free(ctx, d_e)
        if d_v:
            device_free(ctx, d_v)
        if d_w:
            device_free(ctx, d_w)
    return emat, evec, w


def assemble_neohookean_tet(plan, mesh: Mesh, disp: Vec, myu, lam, mat, dw: Vec = None):
    """Hessian -> mat, gradient -> dw, returns the energy"""
    w = C.c_double()
    check(lib().dfb_assemble_neohookean_tet(plan.h, mesh.h, disp.h, float(myu), float(lam), mat.h,
                                            dw.h if dw is not None else None, C.byref(w)))
    return w.value


def assemble_energy(plan, kind, mesh: Mesh, disp: Vec, p0, p1, mat, dw: Vec = None):
    """generic energy-model assembly: Hessian -> mat, gradient -> dw, returns the energy
    (neohookean_tet | spring3 | stvk_tri3 | bending_quad)"""
    k = ELEM[kind] if isinstance(kind, str) else kind
    w = C.c_double()
    check(lib().dfb_assemble_energy(plan.h, k, mesh.h, disp.h, float(p0), float(p1), mat.h, dw.h if dw is not None else None, C.byref(w)))
    return w.value


def lumped_mass(plan, mesh: Mesh, rho, out: Vec):
    """out[v][:] = sum over elements at v of measure / n_node * rho (DERIVED A9'; loop shape of mitc_tri3.rs:304-332)"""
    check(lib().dfb_lumped_mass(plan.h, mesh.h, float(rho), out.h))
    return out
