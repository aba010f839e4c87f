"""wall-clock breakdown of the end-to-end step (host buffers -> solution) at S10"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap
dfb = _bootstrap.load()
from del_fem_b200 import sparse_ilu, synthetic
n = synthetic.SIZES.get(sys.argv[1], None) or int(sys.argv[1])
ctx = dfb.Ctx(0)
h = 1.0 / (n - 1)
e2v, xyz = synthetic.kuhn_tet_mesh(n, n, n, h, index_dtype=np.uint64)
m0 = dfb.Mesh.kuhn(ctx, n)
r2i, i2c = dfb.Pattern.from_uniform_mesh(ctx, m0, False).download()
del m0
flag = synthetic.clamp_flags_z0(n, n, n)
u0h = synthetic.splitmix_uniform(0, 0, 3 * n ** 3, -0.1, 0.1).reshape(-1, 3); u0h[flag != 0] = 0
xout = np.empty((n ** 3, 3))
for a in (e2v, xyz, r2i, i2c, flag, u0h, xout):
    ctx.host_register(a)
def T(label, t0, acc):
    ctx.sync(); t = time.perf_counter(); acc.append((label, (t - t0) * 1e3)); return t
for rep in range(3):
    acc = []; t = time.perf_counter(); t_start = t
    mesh = dfb.Mesh(ctx, e2v, xyz); t = T("mesh (H2D usize conn + xyz)", t, acc)
    pat = dfb.Pattern(ctx, r2i, i2c, 0); t = T("pattern (H2D usize)", t, acc)
    plan = dfb.Plan(ctx, pat, mesh, 0); t = T("plan build", t, acc)
    mat = dfb.Matrix(ctx, pat, 3); t = T("matrix alloc+zero", t, acc)
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0); t = T("assemble", t, acc)
    nv = n ** 3
    u0 = dfb.Vec(ctx, nv, 3, 0, u0h); rhs = dfb.Vec(ctx, nv, 3); t = T("u0 H2D + vec alloc", t, acc)
    mat.mult_vec(rhs, 0.0, -1.0, u0); t = T("rhs = -A u0 (incl. pack)", t, acc)
    rhs.set_fixed_dof_zero(flag); mat.set_fixed_dof(1.0, flag); t = T("BC (H2D flags x2)", t, acc)
    pc = sparse_ilu.Preconditioner("jacobi"); pc.initialize(mat); sparse_ilu.copy_value(pc, mat); sparse_ilu.decompose(pc); t = T("jacobi", t, acc)
    x, q, p = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3); t = T("work vecs", t, acc)
    hist = dfb.preconditioned_conjugate_gradient(rhs, x, q, p, 1e-8, 20000, mat, pc); t = T(f"pcg ({len(hist)-1} its)", t, acc)
    dfb.lib().dfb_vec_download(x.h, xout.ctypes.data); t = T("D2H solution", t, acc)
    del mesh, pat, plan, mat, u0, rhs, pc, x, q, p; t = T("free", t, acc)
    print(f"rep {rep}: total {(t - t_start) * 1e3:.1f} ms :: " + " | ".join(f"{l} {v:.1f}" for l, v in acc), flush=True)
