"""time-to-solution of the config-2 problem with the preconditioners (jacobi, block_jacobi, ilu0)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap
dfb = _bootstrap.load()
from del_fem_b200 import sparse_ilu, synthetic
n = synthetic.SIZES.get(sys.argv[1], None) or int(sys.argv[1])  # usage: precond_compare.py S1 jacobi,block_jacobi,ilu0,ilu0_levels
ctx = dfb.Ctx(0)
mesh = dfb.Mesh.kuhn(ctx, n)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
plan = dfb.Plan(ctx, pat, mesh, 0)
mat = dfb.Matrix(ctx, pat, 3)
mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
nv = mesh.num_vtx
flag = dfb.BcFlag(ctx, synthetic.clamp_flags_z0(n, n, n))
u0, rhs, r, x, q, p = (dfb.Vec(ctx, nv, 3) for _ in range(6))
u0.fill_splitmix(0, 0, -0.1, 0.1); u0.set_fixed_dof_zero_dev(flag)
mat.mult_vec(rhs, 0.0, -1.0, u0); rhs.set_fixed_dof_zero_dev(flag); mat.set_fixed_dof_dev(1.0, flag)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
for kind in sys.argv[2].split(","):
    # "ilu0_levels" = ILU(0) with one launch per level (the round-1 schedule), "ilu0" = persistent sweeps (one launch per sweep, grid barrier per level)
    ctx.set_option("ilu_level_launches", 1 if kind.endswith("_levels") else 0)
    pc = sparse_ilu.Preconditioner(kind.replace("_levels", ""))
    ctx.timer_start(); pc.initialize(mat); t_sym = ctx.timer_stop()
    sparse_ilu.copy_value(pc, mat)
    ctx.timer_start(); sparse_ilu.decompose(pc); t_num = ctx.timer_stop()
    r.copy_from(rhs)
    ctx.timer_start()
    hist = dfb.preconditioned_conjugate_gradient(r, x, q, p, 1e-8, 20000, mat, pc)
    t = ctx.timer_stop()
    print(f"{kind:13s} symbolic {t_sym:8.1f} ms  numeric {t_num:8.2f} ms  PCG {len(hist)-1:5d} its  {t:9.1f} ms  ({t/(len(hist)-1):.3f} ms/it)  rel {hist[-1]/hist[0]:.2e}", flush=True)
