"""times the fused row-tile assembly for a sweep of tile sizes / CTA sizes against the slot-centric gather (variant 4) on one
resident problem, checking identical bits; also the spring3 cloth assembly. usage: asm_sweep.py [S10|<n>] [cloth_n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _bootstrap

dfb = _bootstrap.load()
from del_fem_b200 import elements, synthetic
from del_fem_b200.cloth import edges_of_grid_cloth

arg = sys.argv[1] if len(sys.argv) > 1 else "S10"
n = synthetic.SIZES.get(arg, None) or int(arg)
ctx = dfb.Ctx(0)
mesh = dfb.Mesh.kuhn(ctx, n)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
mat = dfb.Matrix(ctx, pat, 3)
nslots = pat.nnz + pat.num_blk
bytes_alg = 16 * mesh.num_elem + 24 * mesh.num_vtx + 72 * nslots + 64 * mesh.num_elem + 4 * (nslots + 1)


def time_assemble(plan, reps=5):
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
    best = 1e9
    for _ in range(reps):
        ctx.timer_start()
        mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
        best = min(best, ctx.timer_stop())
    return best


ctx.set_option("assemble_variant", 4)
plan = dfb.Plan(ctx, pat, mesh, 0)
t = time_assemble(plan)
ref = mat.download()
print(f"variant 4 (node records + slot gather): {t:.3f} ms  {bytes_alg / t / 1e6:.0f} GB/s algorithmic")
del plan
ctx.set_option("assemble_variant", 0)
for rows, threads in ((8, 0), (16, 0), (16, 128), (16, 384), (24, 0)):
    ctx.set_option("assemble_tile_rows", rows)
    ctx.set_option("assemble_tile_threads", threads)
    plan = dfb.Plan(ctx, pat, mesh, 0)
    t = time_assemble(plan, 3)
    rv, iv = mat.download()
    same = np.array_equal(rv, ref[0]) and np.array_equal(iv, ref[1])
    print(f"tile rows {rows} threads {threads or 'auto'}: {t:.3f} ms  {bytes_alg / t / 1e6:.0f} GB/s algorithmic  identical_bits={same}", flush=True)
    del plan
ctx.set_option("assemble_tile_rows", 0)
ctx.set_option("assemble_tile_threads", 0)
del mat, pat, mesh

# spring3 cloth
cn = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
edges = edges_of_grid_cloth(cn, cn)
g = np.arange(cn) / (cn - 1)
gx, gy = np.meshgrid(g, g)
rest = np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel(), np.zeros(cn * cn)], 1))
mesh = dfb.Mesh(ctx, edges, rest)
pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
mat = dfb.Matrix(ctx, pat, 3)
d = dfb.Vec(ctx, mesh.num_vtx, 3)
d.fill_splitmix(0, 0, -1e-5, 1e-5)
gv = dfb.Vec(ctx, mesh.num_vtx, 3)
ref = None
for variant, rows in ((4, 0), (0, 8), (0, 16), (0, 32)):
    ctx.set_option("assemble_variant", variant)
    ctx.set_option("assemble_tile_rows", rows)
    plan = dfb.Plan(ctx, pat, mesh, 0)
    w = elements.assemble_energy(plan, "spring3", mesh, d, 1.0, 0.0, mat, gv)
    best = 1e9
    for _ in range(5):
        ctx.timer_start()
        w = elements.assemble_energy(plan, "spring3", mesh, d, 1.0, 0.0, mat, gv)
        best = min(best, ctx.timer_stop())
    rv, iv = mat.download()
    gh = gv.download()
    same = ref is None or (np.array_equal(rv, ref[0]) and np.array_equal(iv, ref[1]) and np.array_equal(gh, ref[2]))
    if ref is None:
        ref = (rv, iv, gh, w)
    print(f"spring3 {cn}^2 variant {variant} rows {rows or 'auto'}: {best:.3f} ms  identical_bits={same}  |w - w_ref|/w = {abs(w - ref[3]) / abs(ref[3]):.2e}")
    del plan
