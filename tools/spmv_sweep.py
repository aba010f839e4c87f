"""times the SpMV variants on one resident problem: isolated (back-to-back launches) and inside a PCG loop"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap

dfb = _bootstrap.load()
from del_fem_b200 import sparse_ilu, synthetic

n = synthetic.SIZES.get(sys.argv[1], None) or int(sys.argv[1])
variants = [int(v) for v in sys.argv[2].split(",")]
loop_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
ctx = dfb.Ctx(0)
mesh = dfb.Mesh.kuhn(ctx, n)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
plan = dfb.Plan(ctx, pat, mesh, 0)
mat = dfb.Matrix(ctx, pat, 3)
mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
nv = mesh.num_vtx
flag = dfb.BcFlag(ctx, synthetic.clamp_flags_z0(n, n, n))
u0, rhs, r, x, q, p = (dfb.Vec(ctx, nv, 3) for _ in range(6))
u0.fill_splitmix(0, 0, -0.1, 0.1)
u0.set_fixed_dof_zero_dev(flag)
mat.mult_vec(rhs, 0.0, -1.0, u0)
rhs.set_fixed_dof_zero_dev(flag)
mat.set_fixed_dof_dev(1.0, flag)
pc = sparse_ilu.Preconditioner("jacobi")
pc.initialize(mat)
sparse_ilu.copy_value(pc, mat)
sparse_ilu.decompose(pc)
bytes_ = 76 * pat.nnz + 124 * nv
ref = None
for v in variants:
    ctx.set_option("spmv_variant", v)
    mat.mult_vec(q, 0.0, 1.0, u0)
    y = q.download()
    if ref is None:
        ref = y
    err = np.abs(y - ref).max() / np.abs(ref).max()
    best = 1e9
    for rep in range(3):
        ctx.timer_start()
        for _ in range(20):
            mat.mult_vec(q, 0.0, 1.0, u0)
        best = min(best, ctx.timer_stop() / 20)
    ctx.set_option("profile_spmv", 1)
    ctx.profile_spmv(reset=True)
    r.copy_from(rhs)
    ctx.timer_start()
    hist = dfb.preconditioned_conjugate_gradient(r, x, q, p, 1e-30, loop_iters, mat, pc)
    tot = ctx.timer_stop()
    ms, cnt = ctx.profile_spmv(reset=True)
    ctx.set_option("profile_spmv", 0)
    r.copy_from(rhs)
    ctx.timer_start()
    hist = dfb.preconditioned_conjugate_gradient(r, x, q, p, 1e-30, loop_iters, mat, pc)
    tot_np = ctx.timer_stop()
    print(f"variant {v}: iso {best:.4f} ms {bytes_ / best / 1e6:.0f} GB/s | in-loop {ms / max(cnt, 1):.4f} ms {bytes_ / (ms / max(cnt, 1)) / 1e6:.0f} GB/s "
          f"(n={cnt}) | iter {tot / loop_iters:.4f} ms (unprofiled {tot_np / loop_iters:.4f}) | maxrel vs first {err:.1e}", flush=True)
