"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): pattern + index maps bit-exact; assembled fp64 values <= 1e-12 relative (we get
bit-exact for the linear kernels because the gather reproduces the serial accumulation order and the element formulas
are compiled without FMA contraction); SpMV <= 1e-13 relative; PCG relative residual <= 1e-8 at an iteration count
within +-1 of the oracle's (dot products are deterministic tree sums on the device, left folds in the reference).
"""
import numpy as np
import pytest

from conftest import rel_err
from helpers import device_elasticity_problem, flags_z0, oracle_elasticity_problem

pytestmark = pytest.mark.gpu

VARIANTS = [0, 4, 20]


# ------------------------------------------------------------------------------------------------ mesh / pattern / plan
@pytest.mark.parametrize("shape", [(5, 5, 5), (7, 4, 9), (2, 2, 2), (13, 3, 6)])
def test_kuhn_mesh_device_matches_oracle(dfb, ctx, orc, shape):
    nx, ny, nz = shape
    h = 1.0 / (nx - 1)
    e2v, xyz = orc.kuhn_tet_mesh(nx, ny, nz, h)
    mesh = dfb.Mesh.kuhn(ctx, nx, ny, nz, h)
    e2v_d, xyz_d = mesh.download()
    assert np.array_equal(e2v_d, e2v)
    assert np.array_equal(xyz_d, xyz)
    # host numpy generator used by the end-to-end path
    from del_fem_b200 import synthetic
    e2v_n, xyz_n = synthetic.kuhn_tet_mesh(nx, ny, nz, h)
    assert np.array_equal(e2v_n.astype(np.uint64), e2v)
    assert np.array_equal(xyz_n, xyz)


@pytest.mark.parametrize("is_self", [False, True])
@pytest.mark.parametrize("n", [3, 9])
def test_pattern_from_uniform_mesh_bit_exact(dfb, ctx, orc, n, is_self):
    e2v, xyz = orc.kuhn_tet_mesh(n)
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, xyz.shape[0], is_self)
    mesh = dfb.Mesh(ctx, e2v, xyz)
    pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, is_self)
    r, c = pat.download()
    assert np.array_equal(r, r2i)
    assert np.array_equal(c, i2c)
    assert pat.convention == (1 if is_self else 0)


def test_pattern_tri_mesh_bit_exact(dfb, ctx, orc):
    t2v, xy = orc.grid_tri_mesh(17, disk=True)
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, xy.shape[0], False)
    mesh = dfb.Mesh(ctx, t2v, xy)
    r, c = dfb.Pattern.from_uniform_mesh(ctx, mesh, False).download()
    assert np.array_equal(r, r2i) and np.array_equal(c, i2c)


@pytest.mark.parametrize("flavour", [0, 1])
@pytest.mark.parametrize("convention", [0, 1])
def test_elem2nz_map_bit_exact(dfb, ctx, orc, convention, flavour):
    e2v, xyz = orc.kuhn_tet_mesh(6, 5, 4, 0.25)
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, xyz.shape[0], convention == 1)
    want = orc.elem2nz(convention, flavour, e2v, 4, r2i, i2c)
    mesh = dfb.Mesh(ctx, e2v, xyz)
    pat = dfb.Pattern(ctx, r2i, i2c, convention)
    plan = dfb.Plan(ctx, pat, mesh, flavour)
    got = plan.elem2nz()
    assert np.array_equal(got, want)
    assert plan.num_contrib == e2v.shape[0] * 16
    assert plan.num_slots == i2c.size + (xyz.shape[0] if convention == 0 else 0)


def test_plan_pattern_miss_is_an_error(dfb, ctx, orc):
    e2v, xyz = orc.kuhn_tet_mesh(4)
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, xyz.shape[0], False)
    # drop the last column of row 5 from the pattern
    k = int(r2i[6]) - 1
    i2c_bad = np.delete(i2c, k)
    r2i_bad = r2i.copy()
    r2i_bad[6:] -= 1
    mesh = dfb.Mesh(ctx, e2v, xyz)
    pat = dfb.Pattern(ctx, r2i_bad, i2c_bad, 0)
    with pytest.raises(dfb.DfbError) as ei:
        dfb.Plan(ctx, pat, mesh, 0)
    assert ei.value.status == dfb.STATUS["PATTERN_MISS"]


# ------------------------------------------------------------------------------------------------ element kernels
def test_single_element_kernels_match_oracle(dfb, ctx, orc):
    from del_fem_b200 import elements as el
    rng = np.random.default_rng(7)
    for _ in range(5):
        p2 = rng.standard_normal((3, 2))
        a = el.laplace_tri2_ddw_(1.7, *p2, ctx=ctx)
        assert np.array_equal(a, orc.laplace_tri2_ddw(1.7, *p2))
        a = el.solid_linear_tri2_emat_tri2(1.3, 0.9, *p2, ctx=ctx)
        assert np.array_equal(a, orc.emat_tri2(1.3, 0.9, *p2))
        p3 = rng.standard_normal((3, 3))
        a = el.tri3_emat_cotangent_laplacian(*p3, ctx=ctx)
        assert rel_err(a, orc.cot_laplace_tri3(*p3)) < 1e-14
        p4 = rng.standard_normal((4, 3))
        if orc.tet_volume(*p4) < 0:
            p4[[1, 2]] = p4[[2, 1]]
        a = el.laplace_tet_ddw_(0.7, *p4, ctx=ctx)
        assert np.array_equal(a, orc.laplace_tet_ddw(0.7, *p4))
        a = el.solid_linear_tet_emat_tet(1.1, 2.3, *p4, ctx=ctx)
        assert np.array_equal(a, orc.emat_tet(1.1, 2.3, *p4))


KINDS = [("laplace_tri2", 1.5, 0.0), ("cot_laplace_tri3", 0.0, 0.0), ("solid_linear_tri2", 1.2, 0.8),
         ("laplace_tet", 0.9, 0.0), ("solid_linear_tet", 1.0, 1.0)]


def _mesh_for(orc, kind, rng):
    if "tet" in kind:
        e2v, xyz = orc.kuhn_tet_mesh(6, 5, 7, 0.2)
        xyz = xyz + 0.03 * rng.standard_normal(xyz.shape)
        return e2v, xyz, 4
    t2v, xy = orc.grid_tri_mesh(9, disk=True)
    xy = xy + 0.01 * rng.standard_normal(xy.shape)
    if kind == "cot_laplace_tri3":
        xy = np.concatenate([xy, 0.3 * np.sin(3 * xy[:, :1])], axis=1)
    return t2v, np.ascontiguousarray(xy), 3


@pytest.mark.parametrize("asm_variant", [0, 1, 4])
@pytest.mark.parametrize("convention", [0, 1])
@pytest.mark.parametrize("kind,p0,p1", KINDS)
def test_fused_assembly_matches_oracle(dfb, ctx, orc, kind, p0, p1, convention, asm_variant):
    """asm_variant 0: fused row-tile assembly (default; writes the SpMV's packed records directly; falls back to 4 where it does not
    apply: convention 1); 4: node records + slot-centric gather; 1: single kernel, geometry recomputed per contribution"""
    ctx.set_option("assemble_variant", asm_variant)
    rng = np.random.default_rng(3)
    e2v, xyz, n_node = _mesh_for(orc, kind, rng)
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, n_node, nv, convention == 1)
    rv, iv = orc.assemble(kind, convention, e2v, xyz, p0, p1, r2i, i2c)
    mesh = dfb.Mesh(ctx, e2v, xyz)
    pat = dfb.Pattern(ctx, r2i, i2c, convention)
    # csrdia (merge.rs:31) tests the diagonal on vertex ids, merge_for_array_blk on node ids: same for valid meshes
    plan = dfb.Plan(ctx, pat, mesh, 0)
    n = dfb.core.ELEM_INFO[dfb.ELEM[kind]][2]
    mat = dfb.Matrix(ctx, pat, n)
    mat.assemble(plan, kind, mesh, p0, p1)
    rv_d, iv_d = mat.download()
    if kind == "cot_laplace_tri3":  # sqrt/div heavy third-party formula: 1e-12 contract
        assert rel_err(iv_d, iv) < 1e-13
        if convention == 0:
            assert rel_err(rv_d, rv) < 1e-13
    else:
        assert np.array_equal(iv_d, iv), f"max rel {rel_err(iv_d, iv)}"
        if convention == 0:
            assert np.array_equal(rv_d, rv)
    # two-phase path (element matrices -> gather) gives the same matrix
    from del_fem_b200 import core, elements as el
    emat, _, _ = el.emat_batch(ctx, kind, mesh, p0, p1)
    d_e = core.device_malloc(ctx, emat.nbytes)
    core.memcpy_h2d(ctx, d_e, emat)
    mat2 = dfb.Matrix(ctx, pat, n)
    mat2.assemble_from_emat(plan, d_e)
    rv2, iv2 = mat2.download()
    core.device_free(ctx, d_e)
    assert np.array_equal(iv2, iv_d)
    if convention == 0:
        assert np.array_equal(rv2, rv_d)
    ctx.set_option("assemble_variant", 0)


def test_assembly_is_run_to_run_deterministic(dfb, ctx, orc):
    prob = oracle_elasticity_problem(orc, 9)
    mesh = dfb.Mesh(ctx, prob["e2v"], prob["xyz"])
    pat = dfb.Pattern(ctx, prob["r2i"], prob["i2c"], 0)
    outs = []
    for _ in range(3):
        plan = dfb.Plan(ctx, pat, mesh, 0)
        mat = dfb.Matrix(ctx, pat, 3)
        mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
        outs.append(mat.download())
    for rv, iv in outs[1:]:
        assert np.array_equal(rv, outs[0][0]) and np.array_equal(iv, outs[0][1])


def test_merge_one_matches_oracle(dfb, ctx, orc):
    e2v, xyz = orc.kuhn_tet_mesh(3)
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv = np.zeros((nv, 9))
    iv = np.zeros((i2c.size, 9))
    mat = dfb.Matrix.from_vtx2vtx(r2i, i2c, 3, ctx=ctx)
    mat.set_zero()
    for e in range(0, e2v.shape[0], 5):
        emat = orc.emat_tet(1.0, 1.0, *xyz[e2v[e]])
        orc.merge_for_array_blk(emat, e2v[e], r2i, i2c, rv, iv)
        mat.merge_for_array_blk(emat, e2v[e], None)
    rv_d, iv_d = mat.download()
    assert np.array_equal(rv_d, rv) and np.array_equal(iv_d, iv)
    with pytest.raises(dfb.DfbError) as ei:  # vertices 0 and nv-1 are not neighbours: assert_ne!(idx0, usize::MAX)
        mat.merge_for_array_blk(np.ones((2, 2, 9)), np.array([0, nv - 1]), None)
    assert ei.value.status == dfb.STATUS["PATTERN_MISS"]


# ------------------------------------------------------------------------------------------------ BC / SpMV / vectors
@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_set_fixed_bc_matches_oracle(dfb, ctx, orc, n):
    rng = np.random.default_rng(11)
    e2v, xyz = orc.kuhn_tet_mesh(5, 4, 3, 0.3)
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv = rng.standard_normal((nv, n * n))
    iv = rng.standard_normal((i2c.size, n * n))
    flag = (rng.random((nv, n)) < 0.3).astype(np.int32)
    mat = dfb.Matrix.from_vtx2vtx(r2i, i2c, n, ctx=ctx)
    mat.upload(rv, iv)
    mat.set_fixed_dof(2.5, flag)
    orc.set_fixed_bc(2.5, flag, rv, iv, r2i, i2c)
    rv_d, iv_d = mat.download()
    assert np.array_equal(rv_d, rv) and np.array_equal(iv_d, iv)
    rhs = rng.standard_normal((nv, n))
    v = dfb.Vec.from_numpy(ctx, rhs)
    dfb.sparse_square.set_fix_dof_to_rhs_vector(v, flag)
    orc.set_fix_dof_to_rhs_vector(rhs, flag)
    assert np.array_equal(v.download(), rhs)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_mult_vec_matches_oracle(dfb, ctx, orc, n, variant):
    rng = np.random.default_rng(5)
    e2v, xyz = orc.kuhn_tet_mesh(9, 7, 8, 0.1)
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv = rng.standard_normal((nv, n * n))
    iv = rng.standard_normal((i2c.size, n * n))
    x = rng.standard_normal((nv, n))
    y0 = rng.standard_normal((nv, n))
    ctx.set_option("spmv_variant", variant)
    try:
        mat = dfb.Matrix.from_vtx2vtx(r2i, i2c, n, ctx=ctx)
        mat.upload(rv, iv)
        xd = dfb.Vec.from_numpy(ctx, x)
        for alpha, beta in [(1.0, 0.0), (-0.7, 1.3), (2.0, 1.0)]:
            y = y0.copy()
            orc.mult_vec(y, beta, alpha, x, r2i, i2c, iv, rv)
            yd = dfb.Vec.from_numpy(ctx, y0)
            mat.mult_vec(yd, beta, alpha, xd)
            assert rel_err(yd.download(), y) < 1e-13
    finally:
        ctx.set_option("spmv_variant", 20)


@pytest.mark.parametrize("variant", VARIANTS)
def test_mult_vec_blkcsr_and_ragged_rows(dfb, ctx, orc, variant):
    """convention 1 (diag inside the pattern), empty rows, one very long row (exceeds the staged tile), tiny matrices"""
    rng = np.random.default_rng(9)
    import scipy.sparse as sp
    ctx.set_option("spmv_variant", variant)
    try:
        for nb, dens in [(1, 1.0), (33, 0.2), (700, 0.02)]:
            A = sp.random(nb, nb, density=dens, random_state=3, format="lil")
            if nb > 40:
                A[7, :] = 1.0  # dense row: 700 blocks > tile capacity
                A[11, :] = 0.0  # empty row
            A = A.tocsr()
            A.sort_indices()
            r2i = A.indptr.astype(np.uint64)
            i2c = A.indices.astype(np.uint64)
            iv = rng.standard_normal((i2c.size, 9))
            x = rng.standard_normal((nb, 3))
            pat = dfb.Pattern(ctx, r2i, i2c, 1)
            mat = dfb.Matrix(ctx, pat, 3)
            mat.upload(None, iv)
            blocks = iv.reshape(-1, 3, 3).transpose(0, 2, 1)
            B = sp.bsr_matrix((blocks, A.indices, A.indptr), shape=(3 * nb, 3 * nb))
            y = dfb.Vec(ctx, nb, 3)
            mat.mult_vec(y, 0.0, 1.0, dfb.Vec.from_numpy(ctx, x))
            want = (B @ x.reshape(-1)).reshape(nb, 3)
            assert rel_err(y.download(), want) < 1e-13
    finally:
        ctx.set_option("spmv_variant", 20)


def test_ls_smoke_pattern(dfb, ctx, orc):
    """del-fem-ls/src/sparse_square.rs:166-184: 4-row pattern with explicit diagonal slots, merge + mult_vec"""
    colind = np.array([0, 2, 5, 8, 10], dtype=np.uint64)
    rowptr = np.array([0, 1, 0, 1, 2, 1, 2, 3, 2, 3], dtype=np.uint64)
    sparse = dfb.Matrix.from_vtx2vtx(colind, rowptr, 1, ctx=ctx)   # Matrix::new + symbolic_initialization (:168-171)
    sparse.set_zero()
    emat = np.array([1.0, 0.0, 0.0, 1.0]).reshape(2, 2, 1)
    sparse.merge_for_array_blk(emat, np.array([0, 1]), None, flavour=1)
    rv, iv = sparse.download()
    rv_o, iv_o = np.zeros((4, 1)), np.zeros((10, 1))
    orc.merge_for_array_blk(emat, np.array([0, 1]), colind, rowptr, rv_o, iv_o, flavour=1)
    assert np.array_equal(rv, rv_o) and np.array_equal(iv, iv_o)
    lhs = dfb.Vec(ctx, 4, 1)
    sparse.mult_vec(lhs, 1.0, 1.0, dfb.Vec(ctx, 4, 1))
    assert np.array_equal(lhs.download(), np.zeros((4, 1)))


def test_vector_ops(dfb, ctx, orc):
    rng = np.random.default_rng(2)
    for nb, n in [(1, 1), (1001, 3), (4096, 2), (77777, 3)]:
        a, b = rng.standard_normal((nb, n)), rng.standard_normal((nb, n))
        va, vb = dfb.Vec.from_numpy(ctx, a), dfb.Vec.from_numpy(ctx, b)
        d = va.dot(vb)
        assert abs(d - orc.dot(a, b)) <= 1e-13 * np.sqrt(nb * n) * max(1.0, abs(d))
        assert va.dot(vb) == d  # deterministic
        va.add_scaled(0.37, vb)
        assert rel_err(va.download(), a + 0.37 * b) < 1e-13
        vb.scale_and_add(-1.25, va)
        assert rel_err(vb.download(), (a + 0.37 * b) - 1.25 * b) < 1e-13
        from del_fem_b200 import synthetic
        va.fill_splitmix(3, 11, -0.1, 0.1)
        want = orc.splitmix_fill(3, 11, nb * n, -0.1, 0.1).reshape(nb, n)
        assert np.array_equal(va.download(), want)
        assert np.array_equal(synthetic.splitmix_uniform(3, 11, nb * n, -0.1, 0.1).reshape(nb, n), want)


# ------------------------------------------------------------------------------------------------ solvers
@pytest.mark.parametrize("kind", ["jacobi", "block_jacobi"])
def test_preconditioner_matches_oracle(dfb, ctx, orc, kind):
    prob = oracle_elasticity_problem(orc, 6)
    dev = device_elasticity_problem(dfb, ctx, prob)
    dev["mat"].set_fixed_dof(1.0, prob["flag"])
    ilu_o = orc.Ilu(kind, 3, prob["r2i"], prob["i2c"])
    ilu_o.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    ilu_o.decompose()
    v = np.random.default_rng(1).standard_normal((prob["nv"], 3))
    want = v.copy()
    ilu_o.solve(want)
    from del_fem_b200 import sparse_ilu
    pc = sparse_ilu.Preconditioner(kind)
    pc.initialize(dev["mat"])
    sparse_ilu.copy_value(pc, dev["mat"])
    sparse_ilu.decompose(pc)
    vd = dfb.Vec.from_numpy(ctx, v)
    sparse_ilu.solve_preconditioning_vec(vd, pc)
    assert rel_err(vd.download(), want) < 1e-13


def test_singular_diagonal_is_an_error(dfb, ctx, orc):
    from del_fem_b200 import sparse_ilu
    r2i = np.array([0, 1, 2], dtype=np.uint64)
    i2c = np.array([1, 0], dtype=np.uint64)
    mat = dfb.Matrix.from_vtx2vtx(r2i, i2c, 3, ctx=ctx)
    mat.upload(np.zeros((2, 9)), np.ones((2, 9)))
    pc = sparse_ilu.Preconditioner("block_jacobi")
    pc.initialize(mat)
    sparse_ilu.copy_value(pc, mat)
    with pytest.raises(dfb.DfbError) as ei:
        sparse_ilu.decompose(pc)
    assert ei.value.status == dfb.STATUS["SINGULAR_DIAG"]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n_mesh", [7, 12])
def test_cg_matches_oracle(dfb, ctx, orc, n_mesh, variant):
    prob = oracle_elasticity_problem(orc, n_mesh)
    ctx.set_option("spmv_variant", variant)
    try:
        dev = device_elasticity_problem(dfb, ctx, prob)
        mat = dev["mat"]
        mat.set_fixed_dof(1.0, prob["flag"])
        nv = prob["nv"]
        r = prob["rhs"].copy()
        u, ap, p = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
        hist_o = orc.conjugate_gradient(r, u, ap, p, 1e-8, 2000, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
        rd = dfb.Vec(ctx, nv, 3)
        rd.copy_from(dev["rhs"])
        ud, apd, pd = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
        hist_d = dfb.conjugate_gradient(rd, ud, apd, pd, 1e-8, 2000, mat)
        assert abs(len(hist_d) - len(hist_o)) <= 1, (len(hist_d), len(hist_o))
        m = min(len(hist_d), len(hist_o)) - 1
        assert np.allclose(hist_d[: m // 2], hist_o[: m // 2], rtol=1e-6)
        assert hist_d[-1] < 1e-8
        x = ud.download()
        assert rel_err(x, u) < 1e-6
        # true residual recomputed on the oracle side
        ax = np.zeros_like(x)
        orc.mult_vec(ax, 0.0, 1.0, x, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
        assert np.linalg.norm(prob["rhs"] - ax) / np.linalg.norm(prob["rhs"]) < 2e-8
        # r is overwritten with the final (recursive) residual like the reference
        assert rel_err(rd.download(), prob["rhs"] - ax) < 1e-4
    finally:
        ctx.set_option("spmv_variant", 20)


@pytest.mark.parametrize("kind", ["jacobi", "block_jacobi"])
@pytest.mark.parametrize("n_mesh", [7, 12])
def test_pcg_matches_oracle(dfb, ctx, orc, n_mesh, kind):
    from del_fem_b200 import sparse_ilu
    prob = oracle_elasticity_problem(orc, n_mesh)
    dev = device_elasticity_problem(dfb, ctx, prob, pattern_on_device=True)
    mat = dev["mat"]
    mat.set_fixed_dof(1.0, prob["flag"])
    nv = prob["nv"]
    ilu_o = orc.Ilu(kind, 3, prob["r2i"], prob["i2c"])
    ilu_o.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    ilu_o.decompose()
    r = prob["rhs"].copy()
    x, pr, p = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist_o = orc.preconditioned_conjugate_gradient(r, x, pr, p, 1e-8, 2000, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"], ilu_o)
    pc = sparse_ilu.Preconditioner(kind)
    pc.initialize(mat)
    sparse_ilu.copy_value(pc, mat)
    sparse_ilu.decompose(pc)
    rd = dfb.Vec(ctx, nv, 3)
    rd.copy_from(dev["rhs"])
    xd, prd, pd = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
    hist_d = dfb.preconditioned_conjugate_gradient(rd, xd, prd, pd, 1e-8, 2000, mat, pc)
    assert abs(len(hist_d) - len(hist_o)) <= 1, (len(hist_d), len(hist_o))
    assert abs(hist_d[0] - hist_o[0]) <= 1e-13 * hist_o[0]  # k = 0 entry is the absolute ||r_0||
    m = min(len(hist_d), len(hist_o))
    assert np.allclose(hist_d[: m // 2], hist_o[: m // 2], rtol=1e-6)
    xs = xd.download()
    assert rel_err(xs, x) < 1e-6
    assert rel_err(xs, -prob["u0"]) < 1e-6  # rhs = -K u0 => solution = -u0 on free dofs, 0 on fixed
    ax = np.zeros_like(xs)
    orc.mult_vec(ax, 0.0, 1.0, xs, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    assert np.linalg.norm(prob["rhs"] - ax) / np.linalg.norm(prob["rhs"]) < 2e-8
    # bit-reproducible from run to run
    rd.copy_from(dev["rhs"])
    hist_d2 = dfb.preconditioned_conjugate_gradient(rd, xd, prd, pd, 1e-8, 2000, mat, pc)
    assert np.array_equal(hist_d, hist_d2)
    assert np.array_equal(xd.download(), xs)


def test_degenerate_sizes(dfb, ctx, orc):
    """the smallest inputs the path accepts: ONE tet (4 vertices, a single partial tile everywhere), a 2 x 2 x 2-vertex cube, and
    a matrix with zero rows — pattern, fused assembly, BC, SpMV and PCG against the oracle"""
    from del_fem_b200 import sparse_ilu
    for shape in [(2, 2, 2), None]:
        if shape is None:   # one tet
            e2v = np.array([[0, 1, 2, 3]], dtype=np.uint64)
            xyz = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
        else:
            e2v, xyz = orc.kuhn_tet_mesh(*shape)
        nv = xyz.shape[0]
        r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
        rv, iv = orc.assemble("solid_linear_tet", 0, e2v, xyz, 1.0, 1.0, r2i, i2c)
        mesh = dfb.Mesh(ctx, e2v, xyz)
        pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
        r2i_d, i2c_d = pat.download()
        assert np.array_equal(r2i_d, r2i) and np.array_equal(i2c_d, i2c)
        plan = dfb.Plan(ctx, pat, mesh, 0)
        mat = dfb.Matrix(ctx, pat, 3)
        mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
        rv_d, iv_d = mat.download()
        assert np.array_equal(rv_d, rv) and np.array_equal(iv_d, iv)
        flag = flags_z0(xyz)
        x = np.random.default_rng(0).standard_normal((nv, 3))
        y = np.zeros((nv, 3))
        orc.mult_vec(y, 0.0, 1.0, x, r2i, i2c, iv, rv)
        xd, yd = dfb.Vec.from_numpy(ctx, x), dfb.Vec(ctx, nv, 3)
        mat.mult_vec(yd, 0.0, 1.0, xd)
        assert rel_err(yd.download(), y) < 1e-13
        orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
        mat.set_fixed_dof(1.0, flag)
        rv_d, iv_d = mat.download()
        assert np.array_equal(rv_d, rv) and np.array_equal(iv_d, iv)
        b = np.zeros((nv, 3))
        b[flag == 0] = 1e-2
        pc = sparse_ilu.Preconditioner("jacobi")
        pc.initialize(mat)
        sparse_ilu.copy_value(pc, mat)
        sparse_ilu.decompose(pc)
        rd = dfb.Vec.from_numpy(ctx, b)
        ud, qd, pd = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
        h = dfb.preconditioned_conjugate_gradient(rd, ud, qd, pd, 1e-10, 100, mat, pc)
        ilu_o = orc.Ilu("jacobi", 3, r2i, i2c)
        ilu_o.copy_value(r2i, i2c, iv, rv)
        ilu_o.decompose()
        ro = b.copy()
        xo, pro, po = np.zeros_like(ro), np.zeros_like(ro), np.zeros_like(ro)
        ho = orc.preconditioned_conjugate_gradient(ro, xo, pro, po, 1e-10, 100, r2i, i2c, iv, rv, ilu_o)
        assert abs(len(h) - len(ho)) <= 1
        assert rel_err(ud.download(), xo) < 1e-8
    # zero rows: every entry point is a no-op
    m0 = dfb.Matrix.from_vtx2vtx(np.array([0], dtype=np.uint64), np.array([], dtype=np.uint64), 3, ctx=ctx)
    v0, w0 = dfb.Vec(ctx, 0, 3), dfb.Vec(ctx, 0, 3)
    m0.mult_vec(w0, 0.0, 1.0, v0)
    assert w0.download().shape == (0, 3)


def test_solver_history_semantics(dfb, ctx, orc):
    """A11/A12 quirks: CG early-out returns an empty history; PCG history has k=0 and a trailing entry when max_it
    is exhausted; CG hits max_it with exactly max_it entries."""
    from del_fem_b200 import sparse_ilu
    prob = oracle_elasticity_problem(orc, 6)
    dev = device_elasticity_problem(dfb, ctx, prob)
    mat = dev["mat"]
    mat.set_fixed_dof(1.0, prob["flag"])
    nv = prob["nv"]
    z = dfb.Vec(ctx, nv, 3)
    u, ap, p = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
    assert len(dfb.conjugate_gradient(z, u, ap, p, 1e-5, 100, mat)) == 0
    pc = sparse_ilu.Preconditioner("jacobi")
    pc.initialize(mat)
    sparse_ilu.copy_value(pc, mat)
    sparse_ilu.decompose(pc)
    h = dfb.preconditioned_conjugate_gradient(z, u, ap, p, 1e-5, 100, mat, pc)
    assert len(h) == 1 and h[0] == 0.0
    for max_it in [1, 5, 16, 17, 33]:
        r = dfb.Vec(ctx, nv, 3)
        r.copy_from(dev["rhs"])
        h = dfb.preconditioned_conjugate_gradient(r, u, ap, p, 1e-30, max_it, mat, pc)
        ro = prob["rhs"].copy()
        xo, pro, po = np.zeros_like(ro), np.zeros_like(ro), np.zeros_like(ro)
        ilu_o = orc.Ilu("jacobi", 3, prob["r2i"], prob["i2c"])
        ilu_o.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
        ilu_o.decompose()
        ho = orc.preconditioned_conjugate_gradient(ro, xo, pro, po, 1e-30, max_it, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"], ilu_o)
        assert len(h) == len(ho) == max_it + 2
        assert h[-1] == h[-2]
        assert np.allclose(h, ho, rtol=1e-8)
        r.copy_from(dev["rhs"])
        hc = dfb.conjugate_gradient(r, u, ap, p, 1e-30, max_it, mat)
        assert len(hc) == max_it


def test_energy_identity_and_nullspace_on_device(dfb, ctx, orc):
    """the reference's own check solid_linear_tri2.rs:164-176 (0.5 u^T A u == sum of element energies) and the
    rigid-translation null space, evaluated with the device SpMV on a device-assembled matrix"""
    prob = oracle_elasticity_problem(orc, 10)
    dev = device_elasticity_problem(dfb, ctx, prob)
    nv = prob["nv"]
    t = dfb.Vec.from_numpy(ctx, np.tile([1.0, -2.0, 0.5], (nv, 1)))
    y = dfb.Vec(ctx, nv, 3)
    dev["mat"].mult_vec(y, 0.0, 1.0, t)
    assert np.abs(y.download()).max() < 1e-12
    u = prob["u0"]
    ud = dfb.Vec.from_numpy(ctx, u)
    dev["mat"].mult_vec(y, 0.0, 1.0, ud)
    pap = 0.5 * ud.dot(y)
    eng = 0.0
    for e in range(prob["e2v"].shape[0]):
        nvx = prob["e2v"][e]
        emat = orc.emat_tet(1.0, 1.0, *prob["xyz"][nvx]).reshape(4, 4, 3, 3).transpose(0, 1, 3, 2)  # [i][j][a][b]
        ue = u[nvx.astype(np.int64)]
        eng += 0.5 * np.einsum("ia,ijab,jb->", ue, emat, ue)
    assert abs(pap - eng) < 1e-10


# ------------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("size", ["S1"])
def test_full_size_properties(dfb, ctx, orc, size):
    """BASELINE config 2 at full size (S1 = 70^3 vertices, 1.03 M DoF) through size-independent properties:
    counts, translation null-space, symmetry (x.Ay == y.Ax), SpMV variants agree, PCG reaches 1e-8 with a true
    residual check and solution == -u0, all on the device."""
    from del_fem_b200 import sparse_ilu, synthetic
    n = synthetic.SIZES[size]
    mesh = dfb.Mesh.kuhn(ctx, n)
    assert mesh.num_elem == 6 * (n - 1) ** 3
    pat = dfb.Pattern.from_uniform_mesh(ctx, mesh, False)
    assert pat.nnz == 2 * (3 * n * n * (n - 1) + 3 * n * (n - 1) ** 2 + (n - 1) ** 3)
    plan = dfb.Plan(ctx, pat, mesh, 0)
    assert plan.num_contrib == 16 * mesh.num_elem
    mat = dfb.Matrix(ctx, pat, 3)
    mat.assemble(plan, "solid_linear_tet", mesh, 1.0, 1.0)
    nv = mesh.num_vtx
    t = dfb.Vec.from_numpy(ctx, np.tile([1.0, 1.0, 1.0], (nv, 1)))
    y = dfb.Vec(ctx, nv, 3)
    mat.mult_vec(y, 0.0, 1.0, t)
    assert np.abs(y.download()).max() < 1e-10
    a, b, ya, yb = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
    a.fill_splitmix(1, 0, -1, 1)
    b.fill_splitmix(2, 0, -1, 1)
    results = []
    for variant in VARIANTS:
        ctx.set_option("spmv_variant", variant)
        mat.mult_vec(ya, 0.0, 1.0, a)
        mat.mult_vec(yb, 0.0, 1.0, b)
        results.append(ya.download())
        s1, s2 = b.dot(ya), a.dot(yb)
        assert abs(s1 - s2) <= 1e-11 * abs(s1)
    ctx.set_option("spmv_variant", 20)
    assert all(rel_err(r, results[0]) < 1e-14 for r in results[1:])
    # config-2 solve
    flag = synthetic.clamp_flags_z0(n, n, n)
    u0 = dfb.Vec(ctx, nv, 3)
    u0.fill_splitmix(0, 0, -0.1, 0.1)
    u0.set_fixed_dof_zero(flag)
    rhs = dfb.Vec(ctx, nv, 3)
    mat.mult_vec(rhs, 0.0, -1.0, u0)
    rhs.set_fixed_dof_zero(flag)
    mat.set_fixed_dof(1.0, flag)
    pc = sparse_ilu.Preconditioner("jacobi")
    pc.initialize(mat)
    sparse_ilu.copy_value(pc, mat)
    sparse_ilu.decompose(pc)
    r = dfb.Vec(ctx, nv, 3)
    r.copy_from(rhs)
    x, pr, p = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
    hist = dfb.preconditioned_conjugate_gradient(r, x, pr, p, 1e-8, 20000, mat, pc)
    assert hist[-1] / hist[0] < 1e-8
    mat.mult_vec(pr, 0.0, 1.0, x)  # true residual b - A x
    pr.scale_and_add(-1.0, rhs)
    assert np.sqrt(pr.dot(pr)) / hist[0] < 2e-8
    x.add_scaled(1.0, u0)
    assert np.sqrt(x.dot(x)) / np.sqrt(u0.dot(u0)) < 1e-4  # forward error ~ cond(A) * 1e-8


def test_config2_full_size_matches_oracle(dfb, ctx, orc):
    """BASELINE config 2 AT ITS REAL SIZE (S1 = 70^3 vertices, 1 029 000 DoF) against the oracle, not only through properties:
    pattern and assembled values bit-exact vs `orc.assemble` (the serial element loop + merge_for_array_blk, sparse_square.rs:180-214),
    constrained matrix and rhs bit-exact, Jacobi-PCG iteration count equal (+-1) to the oracle's SERIAL PCG (the loop of
    sparse_square.rs:264-347, generalised to 3x3 blocks), which is also the count bench.py's CPU arms use for S1, and the solution
    within 1e-6.  The oracle side costs ~45 s of one host core (861 iterations x ~50 ms)."""
    import bench
    from helpers import device_elasticity_problem, oracle_elasticity_problem
    from del_fem_b200 import sparse_ilu
    n = 70
    prob = oracle_elasticity_problem(orc, n)
    dev = device_elasticity_problem(dfb, ctx, prob, pattern_on_device=True)
    r2i, i2c = dev["pat"].download()
    assert np.array_equal(r2i, prob["r2i"]) and np.array_equal(i2c, prob["i2c"])
    rv, iv = dev["mat"].download()
    assert np.array_equal(rv, prob["rv0"]) and np.array_equal(iv, prob["iv0"])           # assembled values, before BC
    mat, nv = dev["mat"], prob["nv"]
    assert rel_err(dev["rhs"].download(), prob["rhs"]) < 1e-13                          # rhs = -K u0 (SpMV: 4 partial sums per row), fixed dofs zeroed
    mat.set_fixed_dof(1.0, prob["flag"])
    rv, iv = mat.download()
    assert np.array_equal(rv, prob["rv"]) and np.array_equal(iv, prob["iv"])            # after set_fixed_bc
    pc = sparse_ilu.Preconditioner("jacobi")
    pc.initialize(mat)
    sparse_ilu.copy_value(pc, mat)
    sparse_ilu.decompose(pc)
    x, pr, p = dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3), dfb.Vec(ctx, nv, 3)
    dev["rhs"].upload(prob["rhs"])                                                      # both sides start from the same bits
    hist = dfb.preconditioned_conjugate_gradient(dev["rhs"], x, pr, p, 1e-8, 20000, mat, pc)
    ilu = orc.Ilu("jacobi", 3, prob["r2i"], prob["i2c"])
    ilu.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    ilu.decompose()
    ro = prob["rhs"].copy()
    xo, pro, po = np.zeros_like(ro), np.zeros_like(ro), np.zeros_like(ro)
    hist_o = orc.preconditioned_conjugate_gradient(ro, xo, pro, po, 1e-8, 20000, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"], ilu)
    it_gpu, it_cpu = len(hist) - 1, len(hist_o) - 1
    assert abs(it_gpu - it_cpu) <= 1, (it_gpu, it_cpu)
    assert it_cpu == bench.EXPECTED_ITERS[n], (it_cpu, bench.EXPECTED_ITERS[n])
    m = min(len(hist), len(hist_o)) // 2
    assert np.allclose(hist[:m], hist_o[:m], rtol=1e-6)
    assert rel_err(x.download(), xo) < 1e-6
    ax = np.zeros_like(xo)
    xs = x.download()
    orc.mult_vec(ax, 0.0, 1.0, xs, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])     # true residual on the oracle side
    assert np.linalg.norm(prob["rhs"] - ax) / np.linalg.norm(prob["rhs"]) < 2e-8
