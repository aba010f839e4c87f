"""Device-resident state of BASELINE config 2/5 (3-D linear-elastic Kuhn tet cube, z=0 clamped, rhs = -K u0) for one rank,
and one pass of the hot path over it. Shared by bench.py and the multi-GPU tests."""
from __future__ import annotations


class Problem:
    """device-resident state of one rank: mesh slab, pattern, plan, matrix, flags, work vectors"""

    def __init__(self, dfb, ctx, nx, ny, nz, rank, nranks, precond, lam=1.0, mu=1.0):
        from del_fem_b200 import partition, sparse_ilu
        self.dfb, self.ctx = dfb, ctx
        self.lam, self.mu = lam, mu
        self.h = 1.0 / (nx - 1)
        self.slab = s = partition.make_slab(nx, ny, nz, rank, nranks)
        self.mesh = dfb.Mesh.kuhn(ctx, nx, ny, nz, self.h, s.kb, s.ke)
        self.mesh.rotate_vertex_ids(s.shift)
        self.pat = dfb.Pattern.from_uniform_mesh(ctx, self.mesh, False, s.n_own)
        self.plan = dfb.Plan(ctx, self.pat, self.mesh, 0)
        self.mat = dfb.Matrix(ctx, self.pat, 3)
        self.halo = None
        if nranks > 1:
            self.halo = dfb.Halo(ctx, s.peers, s.send_ptr, s.send_idx, s.recv_ptr)
            self.halo.set_remote(s.remote_off)  # enables the NVLink peer push when the ctx has mapped peers
            self.mat.set_halo(self.halo, s.int_begin, s.int_end)
        self.flag_host = s.flags_clamp_z0()
        self.flag = dfb.BcFlag(ctx, self.flag_host)
        ng = s.n_ghost
        self.u0 = dfb.Vec(ctx, s.n_own, 3, ng)
        self.rhs = dfb.Vec(ctx, s.n_own, 3, ng)
        self.x = dfb.Vec(ctx, s.n_own, 3, ng)
        self.q = dfb.Vec(ctx, s.n_own, 3, ng)
        self.p = dfb.Vec(ctx, s.n_own, 3, ng)
        self.pc = sparse_ilu.Preconditioner(precond)
        self.pc.initialize(self.mat)
        self.n_dof_local = 3 * s.n_own
        self.hist = None

    def step(self, tol=1e-8, max_it=20000):
        """one pass of the hot path, everything already resident in HBM. returns (assembly_ms, total_ms, iterations)"""
        from del_fem_b200 import sparse_ilu
        dfb, ctx, s = self.dfb, self.ctx, self.slab
        ctx.timer_start()
        self.mat.assemble(self.plan, "solid_linear_tet", self.mesh, self.lam, self.mu)
        asm_ms = ctx.timer_stop()
        ctx.timer_start()
        self.u0.fill_splitmix(0, 3 * s.first_global_vtx, -0.1, 0.1)
        self.u0.set_fixed_dof_zero_dev(self.flag)
        self.mat.mult_vec(self.rhs, 0.0, -1.0, self.u0)
        self.rhs.set_fixed_dof_zero_dev(self.flag)
        self.mat.set_fixed_dof_dev(1.0, self.flag)
        sparse_ilu.copy_value(self.pc, self.mat)
        sparse_ilu.decompose(self.pc)
        self.hist = dfb.preconditioned_conjugate_gradient(self.rhs, self.x, self.q, self.p, tol, max_it, self.mat, self.pc)
        rest_ms = ctx.timer_stop()
        return asm_ms, asm_ms + rest_ms, len(self.hist) - 1


    def true_rel_residual(self):
        """||b - A x|| / ||b|| of the last solve, recomputed from scratch OUTSIDE any timed region: b and the constrained matrix are
        rebuilt exactly as step() builds them (the solver overwrote rhs with its recursive residual), then one (distributed)
        SpMV; both norms are all-reduced over the ranks"""
        import math
        s = self.slab
        self.mat.assemble(self.plan, "solid_linear_tet", self.mesh, self.lam, self.mu)
        self.u0.fill_splitmix(0, 3 * s.first_global_vtx, -0.1, 0.1)
        self.u0.set_fixed_dof_zero_dev(self.flag)
        self.mat.mult_vec(self.rhs, 0.0, -1.0, self.u0)
        self.rhs.set_fixed_dof_zero_dev(self.flag)
        self.mat.set_fixed_dof_dev(1.0, self.flag)
        self.mat.mult_vec(self.q, 0.0, 1.0, self.x)
        self.q.add_scaled(-1.0, self.rhs)
        return math.sqrt(self.q.dot(self.q)) / math.sqrt(self.rhs.dot(self.rhs))


class ClothBench:
    """BASELINE config 3: n x n-vertex cloth on the unit square (n = 2048: 4 194 304 vertices, 12 574 721 spring3 edges,
    k = 1), lumped mass rho = 1, two corners pinned, gravity, implicit steps of reassembly + CG to 1e-5 (loop shape of
    rod3_darboux.rs:1174-1263). dt = 1e-3 (not the survey's 1e-2, which leaves the SPD regime on this mesh: tests/test_gpu_widen.py)."""

    def __init__(self, dfb, ctx, n=2048, dt=1e-3):
        import numpy as np
        from del_fem_b200.cloth import MassSpringCloth, edges_of_grid_cloth
        self.n, self.dt = n, dt
        edges = edges_of_grid_cloth(n, n)
        g = np.arange(n) / (n - 1)
        gx, gy = np.meshgrid(g, g)
        rest = np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel(), np.zeros(n * n)], 1))
        flag = np.zeros((n * n, 3), dtype=np.int32)
        flag[0] = 1
        flag[n - 1] = 1
        self.cloth = MassSpringCloth(ctx, edges, rest, flag, 1.0, 1.0 / (n * n), gravity=(0.0, 0.0, -9.8))
        self.n_vtx, self.n_edge, self.nnz = n * n, edges.shape[0], self.cloth.pat.nnz

    def step(self, timing=None):
        return self.cloth.step(self.dt, 1e-5, 3000, timing=timing)


class BeamBench:
    """BASELINE config 4: neo-Hookean beam of 256 x 16 x 16 cells (74 273 vertices, 393 216 tets), mu = 1, lambda = 10,
    x-min face clamped, gravity load ramp in `load_steps` steps; Newton per load step: reassemble W, dW, ddW + BC + Jacobi-PCG
    (1e-8) until ||dW - s f|| has dropped by 1e-4 (fp64 rounding of the internal forces bounds the ratio near 2e-5 at this load)."""

    def __init__(self, dfb, ctx, cells=(256, 16, 16), load=1.0e-5, precond="jacobi"):
        import numpy as np
        from del_fem_b200 import elements as el
        from del_fem_b200.newton import NeoHookeanTetNewton
        nx, ny, nz = cells[0] + 1, cells[1] + 1, cells[2] + 1
        self.mesh = dfb.Mesh.kuhn(ctx, nx, ny, nz, 1.0 / cells[1])
        nv = self.mesh.num_vtx
        flag = np.zeros((nv, 3), dtype=np.int32)
        flag[np.arange(nv) % nx == 0] = 1
        self.nt = NeoHookeanTetNewton(ctx, self.mesh, flag, 1.0, 10.0, precond=precond)
        mass = dfb.Vec(ctx, nv, 3)
        el.lumped_mass(self.nt.plan, self.mesh, 1.0, mass)
        f = np.zeros((nv, 3))
        f[:, 2] = -load * mass.download()[:, 0]
        self.nt.set_load(f)
        self.n_vtx, self.n_tet, self.nnz = nv, self.mesh.num_elem, self.nt.pat.nnz

    def reset(self):
        self.nt.u.set_zero()
        self.nt.log = []

    def ramp(self, load_steps=2, drop=1e-4, max_newton=12, timing=None):
        """the whole gravity ramp from u = 0; returns (newton iterations, total PCG iterations, final ||g|| / first ||g||)"""
        self.reset()
        n_newton = n_pcg = 0
        ratio = 1.0
        for k in range(1, load_steps + 1):
            g0 = None
            for _ in range(max_newton):
                _, gn, its = self.nt.step(k / load_steps, pcg_tol=1e-8, pcg_max=50000, timing=timing)
                n_newton += 1
                n_pcg += its
                g0 = gn if g0 is None else g0
                ratio = gn / g0
                if gn < drop * g0:
                    break
        return n_newton, n_pcg, ratio


class GeneralProblem:
    """one rank of a vertex-partitioned solve on ANY uniform mesh (partition.partition_general): local mesh in the numbering
    [interior owned | boundary owned | ghosts], matrix rows = owned vertices, halo + peer push attached when nranks > 1.
    `flags` / `rhs` / `u` are given and returned in GLOBAL vertex order on every rank (host-side convenience for tests and examples)."""

    def __init__(self, dfb, ctx, elem2vtx, vtx2xyz, rank, nranks, kind="solid_linear_tet", p0=1.0, p1=1.0, precond="jacobi"):
        import numpy as np
        from del_fem_b200 import partition, sparse_ilu
        from del_fem_b200.core import ELEM, ELEM_INFO
        self.dfb, self.ctx, self.kind, self.p0, self.p1 = dfb, ctx, kind, p0, p1
        self.blk_dim = ELEM_INFO[ELEM[kind]][2]
        xyz = np.ascontiguousarray(vtx2xyz, dtype=np.float64)
        self.part = s = partition.partition_general(elem2vtx, xyz.shape[0], nranks, rank)
        self.mesh = dfb.Mesh(ctx, np.ascontiguousarray(s.elem2vtx, dtype=np.uint32), np.ascontiguousarray(xyz[s.local2global]))
        self.pat = dfb.Pattern.from_uniform_mesh(ctx, self.mesh, False, s.n_own)
        self.plan = dfb.Plan(ctx, self.pat, self.mesh, 0)
        self.mat = dfb.Matrix(ctx, self.pat, self.blk_dim)
        self.halo = None
        if nranks > 1:
            self.halo = dfb.Halo(ctx, s.peers, s.send_ptr, s.send_idx, s.recv_ptr)
            self.halo.set_remote(s.remote_off)
            self.mat.set_halo(self.halo, s.int_begin, s.int_end)
        self.pc = sparse_ilu.Preconditioner(precond)
        self.pc.initialize(self.mat)
        self.hist = None

    def solve(self, flags_global, rhs_global, tol=1e-8, max_it=20000):
        """assemble -> Dirichlet rows/columns -> Jacobi-PCG; returns this rank's owned solution and their global vertex ids"""
        import numpy as np
        from del_fem_b200 import sparse_ilu
        dfb, ctx, s, n = self.dfb, self.ctx, self.part, self.blk_dim
        l2g = s.local2global
        flag = dfb.BcFlag(ctx, np.ascontiguousarray(np.asarray(flags_global, dtype=np.int32).reshape(-1, n)[l2g]))
        self.mat.assemble(self.plan, self.kind, self.mesh, self.p0, self.p1)
        rhs = dfb.Vec(ctx, s.n_own, n, s.n_ghost,
                      np.ascontiguousarray(np.asarray(rhs_global, dtype=np.float64).reshape(-1, n)[l2g[: s.n_own]]))
        rhs.set_fixed_dof_zero_dev(flag)
        self.mat.set_fixed_dof_dev(1.0, flag)
        sparse_ilu.copy_value(self.pc, self.mat)
        sparse_ilu.decompose(self.pc)
        x, q, p = (dfb.Vec(ctx, s.n_own, n, s.n_ghost) for _ in range(3))
        self.hist = dfb.preconditioned_conjugate_gradient(rhs, x, q, p, tol, max_it, self.mat, self.pc)
        return x.download(), l2g[: s.n_own]
