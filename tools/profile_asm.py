"""short program for ncu: a few fused assemblies on the S10 (or given) problem"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap

dfb = _bootstrap.load()
from del_fem_b200 import synthetic

n = synthetic.SIZES.get(sys.argv[1], None) or int(sys.argv[1])
ctx = dfb.Ctx(0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
mesh = dfb.Mesh.kuhn(ctx, n)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
plan = dfb.Plan(ctx, pat, mesh, 0)
mat = dfb.Matrix(ctx, pat, 3)
for _ in range(3):
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
ctx.sync()
best = 1e30
for _ in range(5):
    ctx.timer_start()
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
    best = min(best, ctx.timer_stop())
print(f"ok: fused assembly {best:.3f} ms ({' '.join(sys.argv[2:])})")
