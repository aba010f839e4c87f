"""Implicit mass-spring cloth step (BASELINE config 3), device-resident across steps.

Pattern = `RodSimulator::update_dynamic` (del-fem-cpu/src/rod3_darboux.rs:1174-1263): x_tmp = x + dt v (free dofs) ->
assemble W, dW, ddW at x_tmp -> inertia on the diagonal (`row2val[i][0,4,8] += m/dt^2`, :1219-1226) -> zero fixed dofs of dW,
`set_fixed_dof(1.0)` -> CG -> x_tmp -= u -> v = (x_tmp - x)/dt, x = x_tmp.  Springs: `spring3` (3x3 blocks) over the mesh edges.
Gravity enters as a constant force f (dW - f is the out-of-balance force).
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

    def step(self, dt, conv_ratio=1e-5, max_it=1000):
        """one implicit-Euler step; returns (energy at the predictor, CG iterations)"""
        # x_tmp = x + dt v on free dofs (fixed dofs keep v = 0)
        self.vel.set_fixed_dof_zero_dev(self.flag)
        self.disp_tmp.copy_from(self.disp)
        self.disp_tmp.add_scaled(dt, self.vel)
        w = elements.assemble_energy(self.plan, "spring3", self.mesh, self.disp_tmp, self.k, 0.0, self.mat, self.g)
        self.g.add_scaled(-1.0, self.f)                               # dW - f
        self.mat.add_to_diagonal(1.0 / (dt * dt), self.mass)          # inertia (:1219-1226)
        self.g.set_fixed_dof_zero_dev(self.flag)                      # set_fix_dof_to_rhs_vector (:1228)
        self.mat.set_fixed_dof_dev(1.0, self.flag)                    # set_fixed_dof (:1229)
        hist = sparse_square.conjugate_gradient(self.g, self.u, self.ap, self.p, conv_ratio, max_it, self.mat)  # (:1235)
        self.disp_tmp.add_scaled(-1.0, self.u)                        # update_solution (x_tmp -= u)
        # v = (x_tmp - x) / dt ; x = x_tmp
        self.vel.copy_from(self.disp_tmp)
        self.vel.add_scaled(-1.0, self.disp)
        self.vel.scale(1.0 / dt)
        self.disp.copy_from(self.disp_tmp)
        self.log.append((w, len(hist)))
        return w, len(hist)
