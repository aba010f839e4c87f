"""Implicit mass-spring cloth step (BASELINE config 3), device-resident across steps.

Pattern = `RodSimulator::update_dynamic` (del-fem-cpu/src/rod3_darboux.rs:1174-1263): x_tmp = x + dt v (free dofs) ->
assemble W, dW, ddW at x_tmp -> inertia on the diagonal (`row2val[i][0,4,8] += m/dt^2`, :1219-1226) -> zero fixed dofs of dW,
`set_fixed_dof(1.0)` -> CG -> x_tmp -= u -> v = (x_tmp - x)/dt, x = x_tmp.  Springs: `spring3` (3x3 blocks) over the mesh edges.
Gravity enters as a constant force f (dW - f is the out-of-balance force).

`StVKCloth` is the same step with the St.Venant-Kirchhoff membrane (`stvk_tri3::wdwddw_`, stvk_tri3.rs:152-330) on the triangles
plus the quadratic-bending hinge energy (DERIVED from `quadratic_bending::wdw`, quadratic_bending.rs:1-43) on the interior edges.
Both families are merged into one matrix on the hinge-quad pattern (a superset of the triangle pattern).
"""
from __future__ import annotations

import numpy as np

from . import elements, sparse_square
from .core import BcFlag, Ctx, Mesh, Pattern, Plan, Vec


def edges_of_grid_cloth(nx, ny, index_dtype=np.uint64):
    """unique edges of the (nx x ny)-vertex grid split along the (0,0)-(1,1) diagonal: horizontal, vertical, diagonal"""
    idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    hz = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)
    vt = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)
    dg = np.stack([idx[:-1, :-1].ravel(), idx[1:, 1:].ravel()], 1)
    return np.concatenate([hz, vt, dg]).astype(index_dtype)


def tris_of_grid_cloth(nx, ny, index_dtype=np.uint64):
    """the two triangles per cell of the (nx x ny)-vertex grid, split along the (0,0)-(1,1) diagonal, counter-clockwise"""
    idx = np.arange(nx * ny, dtype=np.int64).reshape(ny, nx)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    return np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1).reshape(-1, 3).astype(index_dtype)


def hinges_of_tri_mesh(tri2vtx, index_dtype=np.uint64):
    """4-node hinge elements (v0, v1, v2, v3) of a manifold triangle mesh: one per interior edge v2-v3, with triangles
    (v0, v2, v3) and (v1, v3, v2) as `quadratic_bending::wdw` numbers them (quadratic_bending.rs:7-8). Order: by edge key."""
    t = np.asarray(tri2vtx, dtype=np.int64).reshape(-1, 3)
    nv = int(t.max()) + 1 if t.size else 0
    opp = t.ravel()                                   # half-edge k of triangle: from t[k+1] to t[k+2], opposite t[k]
    frm = t[:, [1, 2, 0]].ravel()
    to = t[:, [2, 0, 1]].ravel()
    key = np.minimum(frm, to) * nv + np.maximum(frm, to)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.nonzero(ks[1:] == ks[:-1])[0]          # positions where two half-edges share an edge
    h1, h2 = order[first], order[first + 1]
    return np.stack([opp[h1], opp[h2], frm[h1], to[h1]], 1).astype(index_dtype)


class MassSpringCloth:
    def __init__(self, ctx: Ctx, edge2vtx, vtx2xyz_rest, blk2isfix, stiffness, mass_per_vtx, gravity=(0.0, 0.0, -9.8),
                 pattern: Pattern = None):
        self.ctx = ctx
        self.mesh = Mesh(ctx, edge2vtx, vtx2xyz_rest)
        self.pat = pattern or Pattern.from_uniform_mesh(ctx, self.mesh, False)
        self.plan = Plan(ctx, self.pat, self.mesh, 0)
        self.mat = sparse_square.Matrix(ctx, self.pat, 3)
        self.flag = BcFlag(ctx, blk2isfix)
        nv = self.mesh.num_vtx
        self.k = float(stiffness)
        m = np.broadcast_to(np.asarray(mass_per_vtx, dtype=np.float64).reshape(-1, 1), (nv, 3)).copy()
        self.mass = Vec.from_numpy(ctx, m)
        self.f = Vec.from_numpy(ctx, m * np.asarray(gravity, dtype=np.float64)[None, :])
        self.disp = Vec(ctx, nv, 3)      # x - rest
        self.disp_tmp = Vec(ctx, nv, 3)
        self.vel = Vec(ctx, nv, 3)
        self.g, self.u, self.ap, self.p = (Vec(ctx, nv, 3) for _ in range(4))
        self.log = []

    def step(self, dt, conv_ratio=1e-5, max_it=1000, timing=None):
        """one implicit-Euler step; returns (energy at the predictor, CG iterations). `timing` (dict) accumulates device ms per
        phase (CUDA events on the ctx stream, one synchronisation per phase): assemble / diag_bc / pack / cg / update"""
        ctx = self.ctx

        def lap(key):
            if timing is not None:
                timing[key] = timing.get(key, 0.0) + ctx.timer_stop()
                ctx.timer_start()

        if timing is not None:
            ctx.timer_start()
        # x_tmp = x + dt v on free dofs (fixed dofs keep v = 0)
        self.vel.set_fixed_dof_zero_dev(self.flag)
        self.disp_tmp.copy_from(self.disp)
        self.disp_tmp.add_scaled(dt, self.vel)
        lap("update")
        w = elements.assemble_energy(self.plan, "spring3", self.mesh, self.disp_tmp, self.k, 0.0, self.mat, self.g)
        lap("assemble")
        self.g.add_scaled(-1.0, self.f)                               # dW - f
        self.mat.add_to_diagonal(1.0 / (dt * dt), self.mass)          # inertia (:1219-1226)
        self.g.set_fixed_dof_zero_dev(self.flag)                      # set_fix_dof_to_rhs_vector (:1228)
        self.mat.set_fixed_dof_dev(1.0, self.flag)                    # set_fixed_dof (:1229)
        lap("diag_bc")
        if timing is not None:
            self.mat.pack()
            lap("pack")
        hist = sparse_square.conjugate_gradient(self.g, self.u, self.ap, self.p, conv_ratio, max_it, self.mat)  # (:1235)
        lap("cg")
        self.disp_tmp.add_scaled(-1.0, self.u)                        # update_solution (x_tmp -= u)
        # v = (x_tmp - x) / dt ; x = x_tmp
        self.vel.copy_from(self.disp_tmp)
        self.vel.add_scaled(-1.0, self.disp)
        self.vel.scale(1.0 / dt)
        self.disp.copy_from(self.disp_tmp)
        lap("update")
        self.log.append((w, len(hist)))
        return w, len(hist)


class StVKCloth:
    """implicit-Euler cloth with StVK membrane + quadratic bending; same step as MassSpringCloth (rod3_darboux.rs:1174-1263)"""

    def __init__(self, ctx: Ctx, tri2vtx, vtx2xyz_rest, blk2isfix, lam, myu, bend_stiffness, mass_per_vtx,
                 gravity=(0.0, 0.0, -9.8)):
        self.ctx = ctx
        self.tri = Mesh(ctx, tri2vtx, vtx2xyz_rest)
        self.hinge = Mesh(ctx, hinges_of_tri_mesh(tri2vtx), vtx2xyz_rest)
        self.pat = Pattern.from_uniform_mesh(ctx, self.hinge, False)
        self.plan_tri = Plan(ctx, self.pat, self.tri, 0)
        self.plan_hinge = Plan(ctx, self.pat, self.hinge, 0)
        self.mat = sparse_square.Matrix(ctx, self.pat, 3)
        self.mat_bend = sparse_square.Matrix(ctx, self.pat, 3)
        self.flag = BcFlag(ctx, blk2isfix)
        nv = self.tri.num_vtx
        self.lam, self.myu, self.kb = float(lam), float(myu), float(bend_stiffness)
        m = np.broadcast_to(np.asarray(mass_per_vtx, dtype=np.float64).reshape(-1, 1), (nv, 3)).copy()
        self.mass = Vec.from_numpy(ctx, m)
        self.f = Vec.from_numpy(ctx, m * np.asarray(gravity, dtype=np.float64)[None, :])
        self.disp = Vec(ctx, nv, 3)
        self.disp_tmp = Vec(ctx, nv, 3)
        self.vel = Vec(ctx, nv, 3)
        self.g, self.g_bend, self.u, self.ap, self.p = (Vec(ctx, nv, 3) for _ in range(5))
        self.log = []

    def assemble(self, disp: Vec):
        """W, dW -> self.g, ddW -> self.mat at the given displacement (membrane + bending)"""
        w = elements.assemble_energy(self.plan_tri, "stvk_tri3", self.tri, disp, self.lam, self.myu, self.mat, self.g)
        w += elements.assemble_energy(self.plan_hinge, "bending_quad", self.hinge, disp, self.kb, 0.0, self.mat_bend, self.g_bend)
        self.mat.add_scaled(1.0, self.mat_bend)
        self.g.add_scaled(1.0, self.g_bend)
        return w

    def step(self, dt, conv_ratio=1e-5, max_it=1000):
        self.vel.set_fixed_dof_zero_dev(self.flag)
        self.disp_tmp.copy_from(self.disp)
        self.disp_tmp.add_scaled(dt, self.vel)
        w = self.assemble(self.disp_tmp)
        self.g.add_scaled(-1.0, self.f)
        self.mat.add_to_diagonal(1.0 / (dt * dt), self.mass)
        self.g.set_fixed_dof_zero_dev(self.flag)
        self.mat.set_fixed_dof_dev(1.0, self.flag)
        hist = sparse_square.conjugate_gradient(self.g, self.u, self.ap, self.p, conv_ratio, max_it, self.mat)
        self.disp_tmp.add_scaled(-1.0, self.u)
        self.vel.copy_from(self.disp_tmp)
        self.vel.add_scaled(-1.0, self.disp)
        self.vel.scale(1.0 / dt)
        self.disp.copy_from(self.disp_tmp)
        self.log.append((w, len(hist)))
        return w, len(hist)
