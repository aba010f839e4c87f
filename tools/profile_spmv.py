"""short program for ncu: builds the S10 (or given) problem on the device and launches a few SpMVs / PCG iterations"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap

dfb = _bootstrap.load()
from del_fem_b200 import sparse_ilu, synthetic

n = synthetic.SIZES.get(sys.argv[1], None) or int(sys.argv[1])
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = dfb.Ctx(0)
ctx.set_option("spmv_variant", variant)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
mesh = dfb.Mesh.kuhn(ctx, n)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
plan = dfb.Plan(ctx, pat, mesh, 0)
mat = dfb.Matrix(ctx, pat, 3)
mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
nv = mesh.num_vtx
flag = dfb.BcFlag(ctx, synthetic.clamp_flags_z0(n, n, n))
u0, rhs, x, q, p = (dfb.Vec(ctx, nv, 3) for _ in range(5))
u0.fill_splitmix(0, 0, -0.1, 0.1)
u0.set_fixed_dof_zero_dev(flag)
mat.mult_vec(rhs, 0.0, -1.0, u0)
rhs.set_fixed_dof_zero_dev(flag)
mat.set_fixed_dof_dev(1.0, flag)
pc = sparse_ilu.Preconditioner("jacobi")
pc.initialize(mat)
sparse_ilu.copy_value(pc, mat)
sparse_ilu.decompose(pc)
hist = dfb.preconditioned_conjugate_gradient(rhs, x, q, p, 1e-8, iters, mat, pc)
ctx.sync()
print("ok", len(hist), hist[-1] / hist[0], "launches", ctx.launch_count())
