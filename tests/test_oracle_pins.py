"""CPU tests that PIN the oracle: the reference's own property tests / golden constants re-run against the C++
restatement.  No GPU needed.  Where the reference has nothing to pin a kernel (tet, neo-Hookean, Jacobi:
none exist upstream) the mathematical properties of SURVEY appendix B are the definition ("parity unpinned")."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_constants.json")))


def csr_from_csrdia(r2i, i2c, rv, iv, n):
    """scipy matrix of a convention-0 (off-diagonal CSR + separate diagonal) block matrix, blocks column-major"""
    nb = r2i.size - 1
    blocks = iv.reshape(-1, n, n).transpose(0, 2, 1)
    A = sp.bsr_matrix((blocks, i2c.astype(np.int64), r2i.astype(np.int64)), shape=(n * nb, n * nb)).tocsr()
    D = sp.block_diag([sp.csr_matrix(b.reshape(n, n).T) for b in rv], format="csr")
    return (A + D).tocsr()


# ---------------------------------------------------------------------------------- del-fem-ls smoke pattern
def test_ls_smoke_pattern(orc):
    g = GOLD["ls_smoke_pattern"]
    r2i, i2c = np.array(g["row2idx"], dtype=np.uint64), np.array(g["idx2col"], dtype=np.uint64)
    rv, iv = np.zeros((4, 1)), np.zeros((10, 1))
    emat = np.array([1.0, 0.0, 0.0, 1.0]).reshape(2, 2, 1)
    orc.merge_for_array_blk(emat, np.array([0, 1]), r2i, i2c, rv, iv, flavour=1)
    assert rv[:, 0].tolist() == [1.0, 1.0, 0.0, 0.0]
    assert np.all(iv == 0.0)  # the explicit diagonal slots stay zero: the vertex-id test routes them to row2val
    lhs, rhs = np.zeros((4, 1)), np.zeros((4, 1))
    orc.mult_vec(lhs, 1.0, 1.0, rhs, r2i, i2c, iv, rv)
    assert np.all(lhs == 0.0)
    # blkcsr on the same pattern (diagonal inside the pattern)
    iv2 = np.zeros((10, 1))
    orc.merge_blkcsr(np.array([0, 1]), np.array([0, 1]), emat, r2i, i2c, iv2)
    assert iv2[:, 0].tolist() == [1.0, 0, 0, 1.0, 0, 0, 0, 0, 0, 0]


def test_merge_panics_like_the_reference(orc):
    g = GOLD["ls_smoke_pattern"]
    r2i, i2c = np.array(g["row2idx"], dtype=np.uint64), np.array(g["idx2col"], dtype=np.uint64)
    rv, iv = np.zeros((4, 1)), np.zeros((10, 1))
    with pytest.raises(orc.OraclePanic):  # (0,3) is not in the pattern: assert_ne!(idx0, usize::MAX)
        orc.merge_for_array_blk(np.ones((2, 2, 1)), np.array([0, 3]), r2i, i2c, rv, iv)
    with pytest.raises(orc.OraclePanic):  # row out of range: assert!(i_row < num_blk)
        orc.merge_csrdia(np.array([0, 9]), np.array([0, 9]), np.ones((2, 2, 1)), r2i, i2c, rv.reshape(-1), iv.reshape(-1))


# ---------------------------------------------------------------------------------- solid_linear_tri2 end-to-end test
def unit_square_problem(orc):
    g = GOLD["tri2_solid_test"]
    n = int(round(1.0 / g["edge"]))  # structured stand-in for meshing_from_polyloop2(edge 0.07)
    t2v, xy = orc.grid_tri_mesh(n, disk=False)
    xy = 0.5 * (xy + 1.0)
    nv = xy.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv, iv = orc.assemble("solid_linear_tri2", 0, t2v, xy, g["lambda"], g["myu"], r2i, i2c)
    flag = np.zeros((nv, 2), dtype=np.int32)
    flag[xy[:, 1] < g["fixed_if_y_below"]] = 1
    u = orc.splitmix_fill(0, 0, 2 * nv, -g["perturbation"], g["perturbation"]).reshape(nv, 2)
    u[flag != 0] = 0.0
    return dict(t2v=t2v, xy=xy, nv=nv, r2i=r2i, i2c=i2c, rv=rv, iv=iv, flag=flag, u=u, g=g)


def element_energy(orc, t2v, xy, u, lam, mu):
    eng = 0.0
    for tri in t2v.astype(np.int64):
        emat = orc.emat_tri2(lam, mu, *xy[tri]).reshape(3, 3, 2, 2).transpose(0, 1, 3, 2)  # [i][j][a][b]
        ue = u[tri]
        e = 0.5 * np.einsum("ia,ijab,jb->", ue, emat, ue)
        assert e >= -1e-10  # solid_linear_tri2.rs:74
        eng += e
    return eng


def test_energy_identity(orc):
    """solid_linear_tri2.rs:164-176: 0.5 u^T A u == sum of element energies (pins merge + mult_vec)"""
    p = unit_square_problem(orc)
    au = np.zeros_like(p["u"])
    orc.mult_vec(au, 0.0, 1.0, p["u"], p["r2i"], p["i2c"], p["iv"], p["rv"])
    pap = 0.5 * orc.dot(p["u"], au)
    eng = element_energy(orc, p["t2v"], p["xy"], p["u"], p["g"]["lambda"], p["g"]["myu"])
    assert abs(pap - eng) < p["g"]["bounds"]["energy_identity_abs"]


def test_solver_quality_bounds(orc):
    """solid_linear_tri2.rs:306-326: CG < 80 its, ILU0-PCG < 36, full-LU PCG history length == 2, dense LU solves"""
    p = unit_square_problem(orc)
    g, b = p["g"], p["g"]["bounds"]
    r2i, i2c, nv = p["r2i"], p["i2c"], p["nv"]
    rhs = np.zeros_like(p["u"])
    orc.mult_vec(rhs, 0.0, -1.0, p["u"], r2i, i2c, p["iv"], p["rv"])  # r = -K u  (per-element accumulation upstream)
    orc.set_fix_dof_to_rhs_vector(rhs, p["flag"])
    rv, iv = p["rv"].copy(), p["iv"].copy()
    orc.set_fixed_bc(1.0, p["flag"], rv, iv, r2i, i2c)

    def check(u_sol):
        r1 = np.zeros_like(u_sol)
        orc.mult_vec(r1, 0.0, 1.0, u_sol, r2i, i2c, iv, rv)
        res = np.max(np.linalg.norm(rhs - r1, axis=1))
        eng = element_energy(orc, p["t2v"], p["xy"], p["u"] + u_sol, g["lambda"], g["myu"])
        return res, eng

    # CG. The reference asserts "< 80 iterations" on ITS mesh (third-party Delaunay mesher, ChaCha perturbation): neither is
    # reproducible here and a right-triangle grid is worse conditioned, so that bound is not transferable. What IS pinned:
    # the oracle's CG is iteration-for-iteration a textbook CG on the same matrix (independent numpy implementation).
    K = csr_from_csrdia(r2i, i2c, rv, iv, 2)
    bvec = rhs.reshape(-1)
    xk, rk = np.zeros_like(bvec), bvec.copy()
    pk, rr = rk.copy(), rk @ rk
    rr0, ratios = rr, []
    for _ in range(300):
        Ap = K @ pk
        al = rr / (pk @ Ap)
        xk += al * pk
        rk -= al * Ap
        rn = rk @ rk
        ratios.append(np.sqrt(rn / rr0))
        if ratios[-1] < g["cg_tol"]:
            break
        pk = rk + (rn / rr) * pk
        rr = rn
    r = rhs.copy()
    u, ap, pp = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist = orc.conjugate_gradient(r, u, ap, pp, g["cg_tol"], 300, r2i, i2c, iv, rv)
    assert abs(len(hist) - len(ratios)) <= 1
    assert np.allclose(hist[:60], ratios[:60], rtol=1e-6)
    res, eng = check(u)
    assert res < b["cg_res_lt"] and eng < 10 * b["cg_energy_lt"], (res, eng, len(hist))
    r = rhs.copy()  # reference call: tol 1e-5, max 100 -> history is capped at 100 entries
    hist = orc.conjugate_gradient(r, np.zeros_like(r), np.zeros_like(r), np.zeros_like(r), g["cg_tol"], g["cg_max"], r2i, i2c, iv, rv)
    assert len(hist) <= g["cg_max"]

    ilu = orc.Ilu("ilu0", 2, r2i, i2c)
    ilu.copy_value(r2i, i2c, iv, rv)
    ilu.decompose()
    r = rhs.copy()
    x, pr, pp = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist = orc.preconditioned_conjugate_gradient(r, x, pr, pp, g["cg_tol"], g["cg_max"], r2i, i2c, iv, rv, ilu)
    res, eng = check(x)
    assert res < b["ilu0_res_lt"] and eng < 10 * b["ilu0_energy_lt"] and len(hist) < b["ilu0_iters_lt"], (res, eng, len(hist))

    full = orc.Ilu("full", 2, r2i, i2c)
    full.copy_value(r2i, i2c, iv, rv)
    full.decompose()
    r = rhs.copy()
    x, pr, pp = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist = orc.preconditioned_conjugate_gradient(r, x, pr, pp, g["cg_tol"], g["cg_max"], r2i, i2c, iv, rv, full)
    assert len(hist) == b["iluinf_hist_len"]
    res, eng = check(x)
    assert res < 1e-13 and eng < 1e-25
    xd = rhs.copy()  # SolverType::DENSE: one application of the full LU
    full.solve(xd)
    res, eng = check(xd)
    assert res < 1e-13 and eng < 1e-25
    # scipy cross-check of the whole path
    xs = spla.spsolve(K.tocsc(), rhs.reshape(-1)).reshape(nv, 2)
    assert np.linalg.norm(xs - xd) / np.linalg.norm(xs) < 1e-10


def test_pcg_history_semantics(orc):
    """A11/A12: CG history = ratios without k=0; PCG = absolute norms with k=0 and a trailing entry at max_it"""
    p = unit_square_problem(orc)
    r2i, i2c = p["r2i"], p["i2c"]
    rhs = np.zeros_like(p["u"])
    orc.mult_vec(rhs, 0.0, -1.0, p["u"], r2i, i2c, p["iv"], p["rv"])
    orc.set_fix_dof_to_rhs_vector(rhs, p["flag"])
    rv, iv = p["rv"].copy(), p["iv"].copy()
    orc.set_fixed_bc(1.0, p["flag"], rv, iv, r2i, i2c)
    z = np.zeros_like(rhs)
    assert len(orc.conjugate_gradient(z.copy(), z.copy(), z.copy(), z.copy(), 1e-5, 10, r2i, i2c, iv, rv)) == 0
    jac = orc.Ilu("block_jacobi", 2, r2i, i2c)
    jac.copy_value(r2i, i2c, iv, rv)
    jac.decompose()
    h = orc.preconditioned_conjugate_gradient(z.copy(), z.copy(), z.copy(), z.copy(), 1e-5, 10, r2i, i2c, iv, rv, jac)
    assert len(h) == 1 and h[0] == 0.0
    r = rhs.copy()
    h = orc.preconditioned_conjugate_gradient(r, z.copy(), z.copy(), z.copy(), 1e-30, 7, r2i, i2c, iv, rv, jac)
    assert len(h) == 9 and h[0] == np.sqrt(orc.dot(rhs, rhs)) and h[-1] == h[-2]
    r = rhs.copy()
    hc = orc.conjugate_gradient(r, z.copy(), z.copy(), z.copy(), 1e-30, 7, r2i, i2c, iv, rv)
    assert len(hc) == 7 and hc[0] < 1.0 + 1e-12
    # block-Jacobi == reference ILU code on an empty pattern: D^-1 applied blockwise
    v = np.random.default_rng(0).standard_normal(rhs.shape)
    w = v.copy()
    jac.solve(w)
    for i in range(0, p["nv"], 17):
        assert np.allclose(w[i], np.linalg.solve(rv[i].reshape(2, 2).T, v[i]), rtol=1e-12)
    pj = orc.Ilu("jacobi", 2, r2i, i2c)
    pj.copy_value(r2i, i2c, iv, rv)
    pj.decompose()
    w = v.copy()
    pj.solve(w)
    assert np.allclose(w, v / rv[:, [0, 3]], rtol=1e-14)


# ---------------------------------------------------------------------------------- Laplacian properties (tests/test_laplace.py:22-30)
@pytest.mark.parametrize("kind", ["laplace_tri2", "cot_laplace_tri3"])
def test_laplacian_properties(orc, kind):
    t2v, xy = orc.grid_tri_mesh(12, disk=True)
    if kind == "cot_laplace_tri3":
        xy = np.ascontiguousarray(np.concatenate([xy, np.zeros((xy.shape[0], 1))], axis=1))
    nv = xy.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv, iv = orc.assemble(kind, 0, t2v, xy, 1.0, 0.0, r2i, i2c)
    L = csr_from_csrdia(r2i, i2c, rv, iv, 1)
    assert abs(L - L.T).max() < 1e-10
    assert np.abs(L @ np.ones(nv)).max() < 1e-10  # reference bound 1e-4 (f32)
    # convention II (diagonal inside the pattern), as del_fem_numpy/LaplacianMatrix.py:8-20 builds it
    r2, c2 = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, True)
    _, iv2 = orc.assemble(kind, 1, t2v, xy, 1.0, 0.0, r2, c2)
    L2 = sp.csr_matrix((iv2[:, 0], c2.astype(np.int64), r2.astype(np.int64)), shape=(nv, nv))
    assert abs(L - L2).max() < 1e-13
    # generalized eigenpair Rayleigh identity with a lumped mass matrix
    area = np.zeros(nv)
    for tri in t2v.astype(np.int64):
        p = xy[tri][:, :2]
        a = 0.5 * abs((p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[2, 0] - p[0, 0]) * (p[1, 1] - p[0, 1]))
        area[tri] += a / 3.0
    M = sp.diags(area)
    w, v = spla.eigsh(L.tocsc(), k=4, M=M.tocsc(), sigma=-0.1)
    for k in range(4):
        assert abs(v[:, k] @ (L @ v[:, k]) - w[k] * (v[:, k] @ (M @ v[:, k]))) < 1e-8
    if kind == "laplace_tri2":  # planar mesh: the cotangent Laplacian equals the P1 Laplacian
        _, ivc = orc.assemble("cot_laplace_tri3", 0, t2v, np.ascontiguousarray(np.concatenate([xy, np.zeros((nv, 1))], 1)), 0, 0, r2i, i2c)
        assert np.abs(ivc - iv).max() < 1e-12


# ---------------------------------------------------------------------------------- derived tet kernels
def test_tet_kernel_properties(orc):
    rng = np.random.default_rng(4)
    for _ in range(10):
        p = rng.standard_normal((4, 3))
        if orc.tet_volume(*p) < 0:
            p[[1, 2]] = p[[2, 1]]
        vol = orc.tet_volume(*p)
        assert abs(vol - np.linalg.det(p[1:] - p[0]) / 6.0) < 1e-14
        lam, mu = 1.3, 0.7
        E = orc.emat_tet(lam, mu, *p).reshape(4, 4, 3, 3).transpose(0, 1, 3, 2)  # [i][j][a][b]
        assert np.allclose(E, E.transpose(1, 0, 3, 2), atol=1e-13)  # E_ij^T = E_ji
        K = E.transpose(0, 2, 1, 3).reshape(12, 12)
        for t in np.eye(3):  # rigid translations
            assert np.abs(K @ np.tile(t, 4)).max() < 1e-12
        w = rng.standard_normal(3)  # infinitesimal rotation
        assert np.abs(K @ np.cross(w, p).reshape(-1)).max() < 1e-11
        eps = rng.standard_normal((3, 3))
        eps = 0.5 * (eps + eps.T)
        u = (p @ eps.T).reshape(-1)  # uniform strain patch energy
        assert abs(0.5 * u @ K @ u - vol * (0.5 * lam * np.trace(eps) ** 2 + mu * np.sum(eps * eps))) < 1e-11
        assert np.linalg.eigvalsh(0.5 * (K + K.T)).min() > -1e-11
        Lp = orc.laplace_tet_ddw(2.0, *p)[:, :, 0]
        assert np.allclose(Lp, Lp.T, atol=1e-14) and np.abs(Lp.sum(axis=1)).max() < 1e-13
        f = rng.standard_normal(3)  # exact for linear fields: x^T L x = alpha V |grad|^2
        xlin = p @ f
        assert abs(xlin @ Lp @ xlin - 2.0 * vol * (f @ f)) < 1e-11


def test_kuhn_mesh_and_pattern(orc):
    n = 7
    e2v, xyz = orc.kuhn_tet_mesh(n)
    assert e2v.shape[0] == 6 * (n - 1) ** 3
    vols = np.array([orc.tet_volume(*xyz[t]) for t in e2v.astype(np.int64)])
    assert vols.min() > 0 and abs(vols.sum() - 1.0) < 1e-12
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, n ** 3, False)
    assert i2c.size == 2 * (3 * n * n * (n - 1) + 3 * n * (n - 1) ** 2 + (n - 1) ** 3)
    assert np.diff(r2i.astype(np.int64)).max() == 14
    # the per-row column SET is mesh adjacency (the order is first-seen: third-party, parity unpinned)
    adj = [set() for _ in range(n ** 3)]
    for t in e2v.astype(np.int64):
        for a in t:
            adj[a].update(int(b) for b in t if b != a)
    for v in range(0, n ** 3, 13):
        row = i2c[int(r2i[v]):int(r2i[v + 1])].astype(np.int64)
        assert len(set(row)) == len(row) and set(row) == adj[v]
    r2, c2 = orc.vtx2vtx_from_uniform_mesh(e2v, 4, n ** 3, True)
    assert c2.size == i2c.size + n ** 3


# ---------------------------------------------------------------------------------- neo-Hookean (FD protocol of the reference)
def _tet_from_hex_constants():
    X = np.array(GOLD["hex_node2xyz"]["value"])[[0, 1, 3, 4]]
    U = np.array(GOLD["hex_node2disp"]["value"])[[0, 1, 3, 4]]
    return X, U


def test_neohookean_density_fd(orc):
    g = GOLD["fd_eps_tol_density"]
    ij = GOLD["istdim2ij"]["value"]
    cv = np.array(GOLD["voigt_c"]["value"])

    def tens(v):
        return np.array([[v[0], v[3], v[5]], [v[3], v[1], v[4]], [v[5], v[4], v[2]]])  # dudx.rs:38-43

    c0 = tens(cv)
    w0, dw0, ddw0 = orc.wr_dwrdc_ddwrddc_neohookean(1.0, 10.0, c0)
    for d in range(6):
        c1 = c0.copy()
        c1[ij[d][0], ij[d][1]] += g["eps"]
        w1, _, _ = orc.wr_dwrdc_ddwrddc_neohookean(1.0, 10.0, c1)
        assert abs((w1 - w0) / g["eps"] - dw0[d]) < 10 * g["tol_dw"]
    for d in range(6):  # symmetrised derivative (solid_incompressive_hex.rs:70-82)
        c1 = c0.copy()
        c1[ij[d][0], ij[d][1]] += 0.5 * g["eps"]
        c1[ij[d][1], ij[d][0]] += 0.5 * g["eps"]
        _, dw1, _ = orc.wr_dwrdc_ddwrddc_neohookean(1.0, 10.0, c1)
        assert np.abs((dw1 - dw0) / g["eps"] - ddw0[d]).max() < 20 * g["tol_ddw"]


def test_neohookean_tet_fd_and_limits(orc):
    g = GOLD["fd_eps_tol_hex"]
    X, U = _tet_from_hex_constants()
    if orc.tet_volume(*X) < 0:
        X[[1, 2]], U[[1, 2]] = X[[2, 1]], U[[2, 1]]
    mu, lam = 1.0, 10.0
    w0, dw0, ddw0 = orc.neohookean_tet_wdwddw(mu, lam, X, U)
    for i in range(4):
        for a in range(3):
            U1 = U.copy()
            U1[i, a] += g["eps"]
            w1, dw1, _ = orc.neohookean_tet_wdwddw(mu, lam, X, U1)
            assert abs((w1 - w0) / g["eps"] - dw0[i, a]) < 20 * g["tol"]
            num = (dw1 - dw0) / g["eps"]  # [j][b]
            ana = ddw0[i].reshape(4, 3, 3)[:, a, :]  # ddw[i][j][a*3+b]
            assert np.abs(num - ana).max() < 50 * g["tol"]
    # zero state, and the Hessian at u = 0 is the linear-elastic tet matrix with the same (lambda, mu)
    w, dw, ddw = orc.neohookean_tet_wdwddw(mu, lam, X, np.zeros((4, 3)))
    assert abs(w) < 1e-14 and np.abs(dw).max() < 1e-13
    lin = orc.emat_tet(lam, mu, *X).reshape(4, 4, 3, 3).transpose(0, 1, 3, 2).reshape(4, 4, 9)  # a + 3b -> a*3 + b
    assert np.abs(ddw - lin).max() < 1e-12


def test_spring3_fd(orc):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 3))
    w0, dw0, ddw0 = orc.spring3_wdwddw(1.7, x, 0.8)
    eps = 1e-6
    for i in range(2):
        for a in range(3):
            x1 = x.copy()
            x1[i, a] += eps
            w1, dw1, _ = orc.spring3_wdwddw(1.7, x1, 0.8)
            assert abs((w1 - w0) / eps - dw0[i, a]) < 1e-5
            ana = ddw0[i].reshape(2, 3, 3)[:, :, a]  # column-major: entry (b,a) at b + 3a ; symmetric
            assert np.abs((dw1 - dw0) / eps - ana).max() < 1e-5


# ---------------------------------------------------------------------------------- cloth elements (SURVEY §8(c) golden vectors)
def test_stvk_tri3_strain_fd(orc):
    """stvk_tri3.rs:121-141: FD of the 2-D Green-Lagrange strain, the reference's triangle, eps 1e-6, tol 1e-5"""
    p_ini = np.array([[0.1, 0.2], [0.2, 0.1], [0.3, 0.4]])
    p0 = np.array([[0.31, 0.04, 0.38], [0.13, 0.13, 0.07], [0.03, 0.42, 0.35]])
    w0, dw = orc.stvk_tri3_wdw(p_ini, p0)
    eps = 1.0e-6
    for i_no in range(3):
        for i_dim in range(3):
            p1 = p0.copy()
            p1[i_no, i_dim] += eps
            w1, _ = orc.stvk_tri3_wdw(p_ini, p1)
            assert np.abs(dw[:, i_no, i_dim] - (w1 - w0) / eps).max() < 1.0e-5


STVK_TRIS = [
    ([[1.2, 2.1, 3.4], [3.5, 5.2, 4.3], [3.4, 4.8, 2.4]], [[3.1, 2.2, 1.5], [4.3, 3.6, 2.0], [5.2, 4.5, 3.4]]),
    ([[3.1, 5.0, 1.3], [2.0, 1.8, 3.4], [5.6, 2.4, 3.3]], [[2.5, 1.0, 3.2], [0.3, 4.6, 1.2], [3.2, 5.5, 1.4]]),
]


@pytest.mark.parametrize("case", [0, 1])
def test_stvk_tri3_wdwddw_fd(orc, case):
    """stvk_tri3.rs:333-362 test_wdwddw_cst: lambda 1.3, myu 1.9, eps 1e-5, tol 1e-4, the reference's two triangles"""
    lam, myu = 1.3, 1.9
    pos0, pos1 = (np.array(a) for a in STVK_TRIS[case])
    w0, dw0, ddw0 = orc.stvk_tri3_wdwddw(pos0, pos1, lam, myu)
    eps = 1.0e-5
    for ino in range(3):
        for idim in range(3):
            pa = pos1.copy()
            pa[ino, idim] += eps
            w1, dw1, _ = orc.stvk_tri3_wdwddw(pos0, pa, lam, myu)
            assert abs((w1 - w0) / eps - dw0[ino, idim]) < 1.0e-4
            num = (dw1 - dw0) / eps  # [jno][jdim]
            ana = ddw0[:, ino, :].reshape(3, 3, 3)[:, :, idim]  # ddw[jno][ino][jdim*3 + idim]
            assert np.abs(num - ana).max() < 1.0e-4
    # rigid motion of the rest triangle: zero energy and gradient; Hessian is symmetric as a 9x9 matrix
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    w, dw, _ = orc.stvk_tri3_wdwddw(pos0, pos0 @ R.T + 0.3, lam, myu)
    assert abs(w) < 1e-12 and np.abs(dw).max() < 1e-11
    H = ddw0.reshape(3, 3, 3, 3).transpose(0, 2, 1, 3).reshape(9, 9)
    assert np.abs(H - H.T).max() < 1e-12 * np.abs(H).max()


def test_quadratic_bending_fd(orc):
    """quadratic_bending.rs:45-79: the reference's hinge quad, eps 1e-5, tol 1e-6; plus the derived energy's FD"""
    p_ini = np.array([[0.1, 0.2, 0.3], [0.2, 0.1, 0.5], [0.3, 0.4, 0.3], [0.5, 0.2, 0.4]])
    p0 = np.array([[0.11, 0.04, 0.38], [0.22, 0.13, 0.07], [0.03, 0.42, 0.35], [0.54, 0.01, 0.46]])
    c0, dcdp, k = orc.quadratic_bending_wdw(p_ini, p0)
    eps = 1.0e-5
    for i_no in range(4):
        for i_dim in range(3):
            p1 = p0.copy()
            p1[i_no, i_dim] += eps
            c1, _, _ = orc.quadratic_bending_wdw(p_ini, p1)
            assert np.abs(dcdp[:, i_no, i_dim] - (c1 - c0) / eps).max() < 1.0e-6
    assert abs(k.sum()) < 1e-12  # translation invariance of the hinge vector
    w0, dw0, ddw0 = orc.quadratic_bending_wdwddw(2.5, p_ini, p0)
    assert abs(w0 - 0.5 * 2.5 * (c0 @ c0)) < 1e-14
    for i_no in range(4):
        for i_dim in range(3):
            p1 = p0.copy()
            p1[i_no, i_dim] += eps
            w1, dw1, _ = orc.quadratic_bending_wdwddw(2.5, p_ini, p1)
            assert abs((w1 - w0) / eps - dw0[i_no, i_dim]) < 1e-3 * max(1.0, abs(dw0[i_no, i_dim]))
            ana = ddw0[:, i_no, :].reshape(4, 3, 3)[:, :, i_dim]  # block (j,i), entry (b, a) at b + 3a
            assert np.abs((dw1 - dw0) / eps - ana).max() < 1e-6 * max(1.0, np.abs(ana).max())


def test_assemble_cloth_matches_dense(orc):
    """whole-mesh StVK / hinge loops: merged matrix == dense sum of element Hessians; gradient == sum of element gradients"""
    from del_fem_b200.cloth import hinges_of_tri_mesh
    t2v, xy = orc.grid_tri_mesh(5)
    nv = xy.shape[0]
    xyz = np.concatenate([xy, np.zeros((nv, 1))], axis=1)
    rng = np.random.default_rng(3)
    disp = 0.05 * rng.standard_normal((nv, 3))
    quads = hinges_of_tri_mesh(t2v)
    for kind, e2v, nn, p0, p1 in (("stvk_tri3", t2v, 3, 1.3, 1.9), ("bending_quad", quads, 4, 0.7, 0.0)):
        r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, nn, nv, False)
        w, dw, rv, iv = orc.assemble_cloth(kind, e2v, xyz, disp, p0, p1, r2i, i2c)
        K = np.zeros((nv, 3, nv, 3))
        g = np.zeros((nv, 3))
        wsum = 0.0
        for nvs in np.asarray(e2v, dtype=np.int64):
            rest, dfm = xyz[nvs], xyz[nvs] + disp[nvs]
            if kind == "stvk_tri3":
                we, dwe, ddwe = orc.stvk_tri3_wdwddw(rest, dfm, p0, p1)
                blk = ddwe.reshape(3, 3, 3, 3)  # [i][j][a][b]
            else:
                we, dwe, ddwe = orc.quadratic_bending_wdwddw(p0, rest, dfm)
                blk = ddwe.reshape(4, 4, 3, 3).transpose(0, 1, 3, 2)  # a + 3b -> [a][b]
            wsum += we
            g[nvs] += dwe
            for i in range(nn):
                for j in range(nn):
                    K[nvs[i], :, nvs[j], :] += blk[i, j]
        assert abs(w - wsum) < 1e-12 * max(1.0, abs(wsum))
        assert np.abs(dw - g).max() < 1e-12 * max(1.0, np.abs(g).max())
        D = np.zeros_like(K)
        for r in range(nv):
            D[r, :, r, :] = rv[r].reshape(3, 3).T  # column-major a + 3b
            for q in range(r2i[r], r2i[r + 1]):
                D[r, :, int(i2c[q]), :] = iv[q].reshape(3, 3).T
        assert np.abs(D - K).max() < 1e-12 * np.abs(K).max()


# ---------------------------------------------------------------------------------- A9' mass
def test_mass_kernels(orc):
    """lumped mass sums to rho * measure, the consistent tet mass has the lumped mass as its row sum, and the plate-bending
    lumped mass (mitc_tri3.rs:304-332) is (A/3 rho t, A/3 rho t^3/12, A/3 rho t^3/12) per corner"""
    e2v, xyz = orc.kuhn_tet_mesh(5, 4, 6, 0.25)
    nv = xyz.shape[0]
    m = orc.lumped_mass(e2v, 4, xyz, 3, 2.5)
    assert abs(m.sum() - 2.5 * (4 * 0.25) * (3 * 0.25) * (5 * 0.25)) < 1e-12
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv, iv = orc.assemble("mass_tet", 0, e2v, xyz, 2.5, 0.0, r2i, i2c)
    rows = rv[:, 0].copy()
    for r in range(nv):
        rows[r] += iv[r2i[r]:r2i[r + 1], 0].sum()
    assert np.abs(rows - m).max() < 1e-14
    assert np.abs(rv[:, [1, 2, 3, 5, 6, 7]]).max() == 0.0 and np.array_equal(rv[:, 0], rv[:, 4]) and np.array_equal(rv[:, 0], rv[:, 8])
    em = orc.emat_mass_tet(3.0, *xyz[e2v[0].astype(np.int64)])
    vol = orc.tet_volume(*xyz[e2v[0].astype(np.int64)])
    assert abs(em[:, :, 0].sum() - 3.0 * vol) < 1e-15 and abs(em[1, 1, 4] - 2 * em[1, 2, 4]) < 1e-18
    t2v, xy = orc.grid_tri_mesh(6)
    ma = orc.lumped_mass(t2v, 3, xy, 2, 1.0)
    assert abs(ma.sum() - 4.0) < 1e-12  # the square [-1, 1]^2
    xyz3 = np.c_[xy, 0.3 * xy[:, 0]]
    m3 = orc.lumped_mass(t2v, 3, xyz3, 3, 1.0)
    assert abs(m3.sum() - 4.0 * np.sqrt(1 + 0.09)) < 1e-12
    thick, rho = 0.05, 7.0
    mp = orc.mass_lumped_plate_bending(t2v, xy, thick, rho)
    assert np.allclose(mp[:, 0], ma * rho * thick, rtol=1e-14) and np.allclose(mp[:, 1], ma * rho * thick ** 3 / 12, rtol=1e-14)
    assert np.array_equal(mp[:, 1], mp[:, 2])


# ---------------------------------------------------------------------------------- config 1: 2-D Poisson on the disk
def test_config1_poisson_disk(orc):
    t2v, xy = orc.grid_tri_mesh(100, disk=True)
    nv = xy.shape[0]
    assert nv == 10201 and t2v.shape[0] == 20000
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv, iv = orc.assemble("laplace_tri2", 0, t2v, xy, 1.0, 0.0, r2i, i2c)
    idx = np.arange(nv)
    i, j = idx % 101, idx // 101
    flag = ((i == 0) | (i == 100) | (j == 0) | (j == 100)).astype(np.int32).reshape(nv, 1)
    rhs = np.zeros((nv, 1))
    for tri in t2v.astype(np.int64):  # lumped f = 1
        p = xy[tri]
        a = 0.5 * ((p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[2, 0] - p[0, 0]) * (p[1, 1] - p[0, 1]))
        rhs[tri, 0] += a / 3.0
    orc.set_fix_dof_to_rhs_vector(rhs, flag)
    orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
    r = rhs.copy()
    u, ap, p = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
    hist = orc.conjugate_gradient(r, u, ap, p, 1e-10, 2000, r2i, i2c, iv, rv)
    assert hist[-1] < 1e-10
    K = csr_from_csrdia(r2i, i2c, rv, iv, 1)
    us = spla.spsolve(K.tocsc(), rhs[:, 0])
    assert np.linalg.norm(u[:, 0] - us) / np.linalg.norm(us) < 1e-8
    # -lap u = 1 on the unit disk, u = 0 on the boundary: u(0) = 1/4
    assert abs(u[50 + 101 * 50, 0] - 0.25) < 2e-3


def test_ls_mult_mat(orc):
    """del-fem-ls mult_mat (sparse_square.rs:134-164): scalar matrix applied to interleaved components"""
    t2v, xy = orc.grid_tri_mesh(6)
    nv = xy.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv, iv = orc.assemble("laplace_tri2", 0, t2v, xy, 1.0, 0.0, r2i, i2c)
    x = np.random.default_rng(0).standard_normal((nv, 3))
    y = np.ones((nv, 3))
    orc.ls_mult_mat(y, 0.5, 2.0, x, r2i, i2c, iv.reshape(-1), rv.reshape(-1))
    L = csr_from_csrdia(r2i, i2c, rv, iv, 1)
    assert np.allclose(y, 0.5 + 2.0 * (L @ x), atol=1e-12)


def test_threaded_baseline_pcg_agrees_with_the_oracle_pcg(orc):
    """bench.py's row-parallel CPU baseline (orc_pcg_jacobi3_mt) is the same iteration as the A12 restatement with Jacobi"""
    nx = 10
    e2v, xyz = orc.kuhn_tet_mesh(nx)
    nv = xyz.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(e2v, 4, nv, False)
    rv, iv = orc.assemble("solid_linear_tet", 0, e2v, xyz, 1.0, 1.0, r2i, i2c)
    flag = np.zeros((nv, 3), dtype=np.int32)
    flag[: nx * nx] = 1
    u0 = orc.splitmix_fill(0, 0, 3 * nv, -0.1, 0.1).reshape(nv, 3)
    u0[flag != 0] = 0.0
    rhs = np.zeros((nv, 3))
    orc.mult_vec(rhs, 0.0, -1.0, u0, r2i, i2c, iv, rv)
    orc.set_fix_dof_to_rhs_vector(rhs, flag)
    orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
    ilu = orc.Ilu("jacobi", 3, r2i, i2c)
    ilu.copy_value(r2i, i2c, iv, rv)
    ilu.decompose()
    r, x, pr, p = rhs.copy(), np.zeros_like(rhs), np.zeros_like(rhs), np.zeros_like(rhs)
    h1 = orc.preconditioned_conjugate_gradient(r, x, pr, p, 1e-8, 1000, r2i, i2c, iv, rv, ilu)
    for nt in (1, 3):
        r2, x2, pr2, p2 = rhs.copy(), np.zeros_like(rhs), np.zeros_like(rhs), np.zeros_like(rhs)
        h2 = orc.pcg_jacobi3_mt(nt, r2, x2, pr2, p2, 1e-8, 1000, r2i, i2c, iv, rv)
        assert len(h2) == len(h1)
        assert np.abs(h2 - h1).max() <= 1e-9 * h1[0]
        assert np.abs(x2 - x).max() <= 1e-9 * np.abs(x).max()


def test_symbolic_iluk_and_iluk_preconditioner(orc):
    """sparse_ilu.rs:221-312: level-of-fill pattern. ILU(k=0) is the ILU(0) pattern (sorted columns); a huge k (the commented-out
    `initialize_iluk(&sparse, 10000)` of linearsystem.rs:55) reproduces the fill of an exact LU, so one PCG step solves the system;
    fill grows and PCG iterations shrink monotonically with k."""
    import scipy.sparse as sp
    t2v, xy = orc.grid_tri_mesh(10)
    nv = xy.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv, iv = orc.assemble("solid_linear_tri2", 0, t2v, xy, 1.0, 1.0, r2i, i2c)
    flag = np.zeros((nv, 2), dtype=np.int32)
    flag[xy[:, 1] == xy[:, 1].min()] = 1
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal((nv, 2))
    orc.set_fix_dof_to_rhs_vector(rhs, flag)
    orc.set_fixed_bc(1.0, flag, rv, iv, r2i, i2c)
    # k = 0: same column sets as A, sorted
    r0, c0, d0 = orc.symbolic_iluk(r2i, i2c, 0)
    assert np.array_equal(r0, r2i)
    for i in range(nv):
        assert np.array_equal(c0[r0[i]:r0[i + 1]], np.sort(i2c[r2i[i]:r2i[i + 1]]))
        up = c0[r0[i]:r0[i + 1]] > i
        assert d0[i] == (r0[i] + int(np.argmax(up)) if up.any() else r0[i + 1])
    # exact symbolic LU fill (dense boolean elimination) == ILU(k -> infinity)
    A = np.zeros((nv, nv), dtype=bool)
    for i in range(nv):
        A[i, i2c[r2i[i]:r2i[i + 1]].astype(np.int64)] = True
        A[i, i] = True
    F = A.copy()
    for k in range(nv):
        rows = np.nonzero(F[k + 1:, k])[0] + k + 1
        F[np.ix_(rows, np.nonzero(F[k, k + 1:])[0] + k + 1)] = True
    rinf, cinf, _ = orc.symbolic_iluk(r2i, i2c, 10000)
    for i in (0, 1, nv // 3, nv // 2, nv - 2, nv - 1):
        want = np.nonzero(F[i])[0]
        assert np.array_equal(cinf[rinf[i]:rinf[i + 1]].astype(np.int64), want[want != i])
    sizes, iters = [], []
    for k in (0, 1, 2, 4, 10000):
        ilu = orc.Ilu(("iluk", k), 2, r2i, i2c)
        ilu.copy_value(r2i, i2c, iv, rv)
        ilu.decompose()
        r = rhs.copy()
        x, pr, pp = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
        hist = orc.preconditioned_conjugate_gradient(r, x, pr, pp, 1e-10, 200, r2i, i2c, iv, rv, ilu)
        y = np.zeros_like(x)
        orc.mult_vec(y, 0.0, 1.0, x, r2i, i2c, iv, rv)
        assert np.abs(y - rhs).max() < 1e-7
        sizes.append(int(orc.symbolic_iluk(r2i, i2c, k)[1].size))
        iters.append(len(hist))
    assert sizes == sorted(sizes) and sizes[0] == i2c.size and sizes[-1] > 3 * sizes[0]
    assert iters == sorted(iters, reverse=True) and iters[-1] == 2, iters
    # ILU(0) through the k = 0 path is the ILU(0) preconditioner bit for bit
    a, b = orc.Ilu("ilu0", 2, r2i, i2c), orc.Ilu(("iluk", 0), 2, r2i, i2c)
    for s in (a, b):
        s.copy_value(r2i, i2c, iv, rv)
        s.decompose()
    v1, v2 = rhs.copy(), rhs.copy()
    a.solve(v1)
    b.solve(v2)
    assert np.array_equal(v1, v2)


def test_mult_square_matrices(orc):
    """del-fem-ls sparse_matrix_multiplication.rs:109-188: C = M0 * M1 of two scalar csrdia matrices equals the scipy product;
    the pattern is the set of structurally reachable columns (diagonal kept apart), in discovery order"""
    t2v, xy = orc.grid_tri_mesh(7, disk=True)
    nv = xy.shape[0]
    r2i, i2c = orc.vtx2vtx_from_uniform_mesh(t2v, 3, nv, False)
    rv0, iv0 = orc.assemble("laplace_tri2", 0, t2v, xy, 1.0, 0.0, r2i, i2c)
    rng = np.random.default_rng(1)
    rv1, iv1 = rng.standard_normal(rv0.shape), rng.standard_normal(iv0.shape)  # M1: same pattern, non-symmetric values
    A = csr_from_csrdia(r2i, i2c, rv0, iv0, 1)
    B = csr_from_csrdia(r2i, i2c, rv1, iv1, 1)
    r, c, iv, rv = orc.mult_square_matrices(r2i, i2c, iv0, rv0, r2i, i2c, iv1, rv1)
    Cm = csr_from_csrdia(r, c, rv.reshape(-1, 1), iv.reshape(-1, 1), 1)
    want = (A @ B).toarray()
    assert np.abs(Cm.toarray() - want).max() < 1e-12 * np.abs(want).max()
    S = ((abs(A) > 0).astype(int) @ (abs(B) > 0).astype(int)).toarray() > 0   # structural product (no cancellation)
    for i in (0, 3, nv // 2, nv - 1):
        cols = c[r[i]:r[i + 1]].astype(np.int64)
        assert len(set(cols)) == cols.size and i not in cols
        assert set(cols) == set(np.nonzero(S[i])[0]) - {i}
        # discovery order: first the columns reached through A's first off-diagonal entry k (B's row k, then k itself), ...
        k = int(i2c[r2i[i]])
        first = [int(j) for j in i2c[r2i[k]:r2i[k + 1]] if int(j) != i] + [k]
        assert list(cols[: len(first)]) == first
