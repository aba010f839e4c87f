"""shared problem builders for the parity tests (oracle side and device side on identical inputs)"""
import numpy as np


def flags_z0(xyz, ndof=3):
    f = np.zeros((xyz.shape[0], ndof), dtype=np.int32)
    f[xyz[:, 2] == 0.0] = 1
    return f


def oracle_elasticity_problem(orc, n, lam=1.0, mu=1.0, nz=None):
    """config-2 recipe on an n^3 (or n x n x nz) Kuhn cube, entirely on the oracle. returns a dict"""
    e2v, xyz = orc.kuhn_tet_mesh(n, n, nz or n, 1.0 / (n - 1))
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv, iv = orc.assemble("solid_linear_tet", 0, e2v, xyz, lam, mu, r2i, i2c)
    rv0, iv0 = rv.copy(), iv.copy()
    flag = flags_z0(xyz)
    u0 = orc.splitmix_fill(0, 0, 3 * nv, -0.1, 0.1).reshape(nv, 3)
    u0[flag != 0] = 0.0
    rhs = np.zeros((nv, 3))
    orc.mult_vec(rhs, 0.0, -1.0, u0, r2i, i2c, iv, rv)
    orc.set_fix_dof_to_rhs_vector(rhs, flag)
    orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
    return dict(e2v=e2v, xyz=xyz, nv=nv, r2i=r2i, i2c=i2c, rv0=rv0, iv0=iv0, rv=rv, iv=iv, flag=flag, u0=u0, rhs=rhs)


def device_elasticity_problem(dfb, ctx, prob, lam=1.0, mu=1.0, pattern_on_device=False):
    """same recipe through the C ABI, starting from the oracle's host mesh. returns dict of device objects"""
    mesh = dfb.Mesh(ctx, prob["e2v"], prob["xyz"])
    if pattern_on_device:
        pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
    else:
        pat = dfb.Pattern(ctx, prob["r2i"], prob["i2c"], 0)
    plan = dfb.Plan(ctx, pat, mesh, 0)
    mat = dfb.Matrix(ctx, pat, 3)
    mat.assemble(plan, "solid_linear_tet", mesh, lam, mu)
    nv = prob["nv"]
    u0 = dfb.Vec(ctx, nv, 3)
    u0.fill_splitmix(0, 0, -0.1, 0.1)
    u0.set_fixed_dof_zero(prob["flag"])
    rhs = dfb.Vec(ctx, nv, 3)
    mat.mult_vec(rhs, 0.0, -1.0, u0)
    rhs.set_fixed_dof_zero(prob["flag"])
    return dict(mesh=mesh, pat=pat, plan=plan, mat=mat, u0=u0, rhs=rhs)
