"""Newton driver for the derived neo-Hookean tet model (BASELINE config 4), device-resident across iterations.

Pattern = the reference's only Newton loop, `RodSimulator::update_static` (del-fem-cpu/src/rod3_darboux.rs:1113-1172):
assemble W, dW, ddW -> zero the fixed dofs of the rhs (`set_fix_dof_to_rhs_vector`) -> `set_fixed_dof(1.0)` -> CG/PCG -> update.
Here: minimise  Pi(u) = W(u) - f.u ;  K du = dW - f ;  u <- u - du.
"""
from __future__ import annotations

import numpy as np

from . import elements, sparse_ilu, sparse_square
from .core import BcFlag, Ctx, Mesh, Pattern, Plan, Vec


class NeoHookeanTetNewton:
    def __init__(self, ctx: Ctx, mesh: Mesh, blk2isfix, myu, lam, precond="jacobi", pattern: Pattern = None):
        self.ctx, self.mesh, self.myu, self.lam = ctx, mesh, float(myu), float(lam)
        self.pat = pattern or Pattern.from_uniform_mesh(ctx, mesh, False)
        self.plan = Plan(ctx, self.pat, mesh, 0)
        self.mat = sparse_square.Matrix(ctx, self.pat, 3)
        self.flag = BcFlag(ctx, blk2isfix)
        nv = mesh.num_vtx
        self.u = Vec(ctx, nv, 3)
        self.f = Vec(ctx, nv, 3)
        self.g = Vec(ctx, nv, 3)
        self.du, self.q, self.p = Vec(ctx, nv, 3), Vec(ctx, nv, 3), Vec(ctx, nv, 3)
        self.pc = sparse_ilu.Preconditioner(precond)
        self.pc.initialize(self.mat)
        self.log = []

    def set_load(self, f_host):
        self.f.upload(f_host)

    def step(self, load_scale=1.0, pcg_tol=1e-8, pcg_max=5000, timing=None):
        """one Newton iteration; returns (energy W, ||dW - s f|| over free dofs, PCG iterations). `timing` (dict) accumulates
        device ms per phase (CUDA events on the ctx stream): assemble / rhs_bc / precond / pack / pcg / update"""
        ctx = self.ctx

        def lap(key):
            if timing is not None:
                timing[key] = timing.get(key, 0.0) + ctx.timer_stop()
                ctx.timer_start()

        if timing is not None:
            ctx.timer_start()
        w = elements.assemble_neohookean_tet(self.plan, self.mesh, self.u, self.myu, self.lam, self.mat, self.g)
        lap("assemble")
        self.g.add_scaled(-float(load_scale), self.f)          # g = dW - s f
        self.g.set_fixed_dof_zero_dev(self.flag)               # set_fix_dof_to_rhs_vector
        gnorm = float(np.sqrt(self.g.dot(self.g)))
        self.mat.set_fixed_dof_dev(1.0, self.flag)              # set_fixed_dof::<3>(1.0, ...)
        lap("rhs_bc")
        sparse_ilu.copy_value(self.pc, self.mat)
        sparse_ilu.decompose(self.pc)
        lap("precond")
        if timing is not None:
            self.mat.pack()
            lap("pack")
        hist = sparse_square.preconditioned_conjugate_gradient(self.g, self.du, self.q, self.p, pcg_tol, pcg_max, self.mat, self.pc)
        lap("pcg")
        self.u.add_scaled(-1.0, self.du)                        # u <- u - du
        lap("update")
        self.log.append((w, gnorm, len(hist) - 1))
        return w, gnorm, len(hist) - 1

    def solve(self, load_steps=1, tol_ratio=1e-8, max_newton=30):
        """gravity-ramp Newton: for s in ramp: iterate until ||g|| / ||g_first|| < tol_ratio"""
        for k in range(1, load_steps + 1):
            s = k / load_steps
            g0 = None
            for _ in range(max_newton):
                _, gn, _ = self.step(s)
                if g0 is None:
                    g0 = max(gn, 1e-300)
                if gn / g0 < tol_ratio:
                    break
        return self.u.download()
