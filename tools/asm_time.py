"""times the fused assembly variants on one resident problem and checks they produce identical bits"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap

dfb = _bootstrap.load()
from del_fem_b200 import synthetic

arg = sys.argv[1] if len(sys.argv) > 1 else "S10"
n = synthetic.SIZES.get(arg, None) or int(arg)
ctx = dfb.Ctx(0)
mesh = dfb.Mesh.kuhn(ctx, n)
ctx.timer_start()
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
t_pat = ctx.timer_stop()
ctx.timer_start()
plan = dfb.Plan(ctx, pat, mesh, 0)
t_plan = ctx.timer_stop()
print(f"pattern {t_pat:.2f} ms, plan {t_plan:.2f} ms, {mesh.num_elem} tets, {plan.num_slots} slots")
mat = dfb.Matrix(ctx, pat, 3)
ref = None
bytes_alg = 16 * mesh.num_elem + 24 * mesh.num_vtx + 72 * plan.num_slots + 64 * mesh.num_elem + 4 * (plan.num_slots + 1)
for variant in (4, 0):
    ctx.set_option("assemble_variant", variant)
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
    best = 1e9
    for _ in range(5):
        ctx.timer_start()
        mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
        best = min(best, ctx.timer_stop())
    rv, iv = mat.download()
    h = (float(np.abs(rv).sum()), float(np.abs(iv).sum()))
    same = ref is None or (np.array_equal(rv, ref[0]) and np.array_equal(iv, ref[1]))
    if ref is None:
        ref = (rv, iv)
    print(f"assemble_variant {variant}: {best:.3f} ms  {mesh.num_elem / best / 1e6:.2f} G tets/s  {bytes_alg / best / 1e6:.0f} GB/s algorithmic  identical_bits={same} checksum={h}")

# two-phase path (element matrices in HBM -> gather): slot-centric vs row-tile
from del_fem_b200 import elements  # noqa: E402
if n <= 110:
    d = dfb.Vec(ctx, mesh.num_vtx, 3)
    d.fill_splitmix(0, 0, -1e-3, 1e-3)
    g = dfb.Vec(ctx, mesh.num_vtx, 3)
    ref = None
    for variant in (4, 0):
        ctx.set_option("assemble_variant", variant)
        ctx.set_option("assemble_tile_neo", 1)
        elements.assemble_energy(plan, "neohookean_tet", mesh, d, 1.0, 10.0, mat, g)
        best = 1e9
        for _ in range(3):
            ctx.timer_start()
            elements.assemble_energy(plan, "neohookean_tet", mesh, d, 1.0, 10.0, mat, g)
            best = min(best, ctx.timer_stop())
        rv, iv = mat.download()
        same = ref is None or (np.array_equal(rv, ref[0]) and np.array_equal(iv, ref[1]))
        ref = ref if ref is not None else (rv, iv)
        print(f"neo-Hookean energy assembly, variant {variant}: {best:.3f} ms identical_bits={same}")
