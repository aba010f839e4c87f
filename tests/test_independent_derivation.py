"""A SECOND, code-independent derivation of every DERIVED element of the hot path (none of them exists in the
reference, so the oracle's hand-written formulas are otherwise their only definition).

Nothing here shares a line with `oracle/` or with the CUDA kernels: the continuum energy of ONE linear tet is written down from the
textbook definition in torch fp64 — shape-function gradients from a matrix inverse (not the oracle's cofactor formulas), strain /
deformation gradient from them, volume from a determinant — and its gradient and Hessian come from reverse-mode autodiff
(`torch.autograd.functional.hessian`), i.e. no hand-derived chain rule at all. The oracle's element matrices must agree to
rounding.  This narrows "parity unpinned" for the derived kernels to "two independent derivations agree"; it cannot make the
oracle <-> reference link green (no Rust toolchain here).

Also: an independent Jacobi-PCG on a scipy BSR matrix (textbook recurrences, scipy SpMV) against the oracle's 3x3 PCG (SURVEY A.6/A.7,
which the reference only has for 2x2 blocks + ILU)."""
import numpy as np
import pytest
import torch

torch.set_default_dtype(torch.float64)

RNG = np.random.default_rng(20261020)


def random_tet():
    """a well-shaped, positively oriented tet with random vertices"""
    while True:
        x = RNG.standard_normal((4, 3))
        vol = np.linalg.det(x[1:] - x[0]) / 6.0
        if vol > 0.05:
            return x


def shape_gradients(X):
    """g[i] = grad N_i of the linear tet with vertices X (4x3 tensor): N_i(x) = a_i + g_i.x with N_i(X_j) = delta_ij, i.e. the
    rows 1..3 of inverse([1 X]) transposed — solved as one 4x4 linear system."""
    A = torch.cat([torch.ones(4, 1), X], dim=1)          # A[j] = [1, X_j]
    coef = torch.linalg.inv(A)                           # column i = [a_i, g_i]
    return coef[1:, :].T                                 # [4][3]


def volume(X):
    return torch.linalg.det(X[1:] - X[0]) / 6.0


def hessian_blocks(energy, u0):
    """d2W/du_i,a du_j,b as [i][j][a][b] and dW/du as [i][a]"""
    u0 = u0.clone().requires_grad_(True)
    g = torch.autograd.grad(energy(u0), u0, create_graph=False)[0]
    H = torch.autograd.functional.hessian(energy, u0.detach())
    return g.detach().numpy(), H.permute(0, 2, 1, 3).detach().numpy()   # hessian() returns [i][a][j][b]


def test_linear_elastic_tet_matches_autodiff_of_the_strain_energy(orc):
    """E[i][j][a + 3b] must be the Hessian of W(u) = V (lambda/2 tr(eps)^2 + mu eps:eps), eps = sym(grad u)"""
    for lam, mu in ((1.0, 1.0), (2.3, 0.7), (0.0, 1.0)):
        Xn = random_tet()
        X = torch.tensor(Xn)
        G, V = shape_gradients(X), volume(X)

        def W(u):
            gu = u.T @ G                                  # grad u [a][b] = sum_i u_i[a] g_i[b]
            eps = 0.5 * (gu + gu.T)
            return V * (0.5 * lam * torch.trace(eps) ** 2 + mu * (eps * eps).sum())

        _, H = hessian_blocks(W, torch.zeros(4, 3))
        E = orc.emat_tet(lam, mu, *Xn).reshape(4, 4, 3, 3).transpose(0, 1, 3, 2)   # stored column-major: [i][j][b][a] -> [i][j][a][b]
        assert np.abs(E - H).max() <= 1e-12 * np.abs(H).max()
        # and the volume the oracle uses is the determinant one
        assert abs(orc.tet_volume(*Xn) - float(V)) <= 1e-14 * abs(float(V))


def test_laplace_tet_matches_autodiff_of_the_dirichlet_energy(orc):
    """E[i][j] = Hessian of alpha/2 * V * |grad phi|^2"""
    Xn = random_tet()
    X = torch.tensor(Xn)
    G, V = shape_gradients(X), volume(X)
    alpha = 1.7

    def W(phi):
        g = phi @ G
        return 0.5 * alpha * V * (g * g).sum()

    H = torch.autograd.functional.hessian(W, torch.zeros(4)).numpy()
    E = orc.laplace_tet_ddw(alpha, *Xn).reshape(4, 4)
    assert np.abs(E - H).max() <= 1e-12 * np.abs(H).max()


def test_consistent_tet_mass_matches_quadrature(orc):
    """M[i][j] = rho * integral N_i N_j, integrated with the degree-2-exact 4-point Gauss rule of the tet
    (barycentric points (a,b,b,b), a = 0.5854..., weight V/4) instead of the closed form rho V/20 (1 + delta_ij)"""
    Xn = random_tet()
    V = float(volume(torch.tensor(Xn)))
    rho = 2.5
    a, b = (5.0 + 3.0 * np.sqrt(5.0)) / 20.0, (5.0 - np.sqrt(5.0)) / 20.0
    pts = np.full((4, 4), b) + (a - b) * np.eye(4)         # barycentric coordinates = shape-function values
    M = rho * (V / 4.0) * np.einsum("qi,qj->ij", pts, pts)
    E = orc.emat_mass_tet(rho, *Xn).reshape(4, 4, 3, 3)
    for i in range(4):
        for j in range(4):
            assert np.allclose(E[i, j], M[i, j] * np.eye(3), rtol=1e-13, atol=0.0)


@pytest.mark.parametrize("mu,lam", [(1.0, 10.0), (0.8, 2.0)])
def test_neohookean_tet_matches_autodiff_of_the_energy(orc, mu, lam):
    """W = V [ mu/2 (tr C - 3) - mu ln J + lambda/2 (ln J)^2 ], F = I + grad u, C = F^T F, J = det F.
    Energy, gradient and Hessian of the oracle's chain rule (solid_incompressive_hex.rs:86-148 restated for 4 nodes) against
    autodiff of that scalar, at rest and at a finite random deformation."""
    Xn = random_tet()
    X = torch.tensor(Xn)
    G, V = shape_gradients(X), volume(X)

    def W(u):
        F = torch.eye(3) + u.T @ G
        J = torch.linalg.det(F)
        Cm = F.T @ F
        lnj = torch.log(J)
        return V * (0.5 * mu * (torch.trace(Cm) - 3.0) - mu * lnj + 0.5 * lam * lnj * lnj)

    for scale in (0.0, 0.15):
        un = scale * RNG.standard_normal((4, 3))
        w_o, dw_o, ddw_o = orc.neohookean_tet_wdwddw(mu, lam, Xn, un)
        g, H = hessian_blocks(W, torch.tensor(un))
        w = float(W(torch.tensor(un)))
        assert abs(w_o - w) <= 1e-12 * max(1.0, abs(w))
        assert np.abs(dw_o - g).max() <= 1e-11 * max(1.0, np.abs(g).max())
        Ho = ddw_o.reshape(4, 4, 3, 3)                      # the reference convention a*3+b: already [i][j][a][b]
        assert np.abs(Ho - H).max() <= 1e-11 * np.abs(H).max()
    # at rest the tangent is linear elasticity with the same (lambda, mu)
    _, _, ddw0 = orc.neohookean_tet_wdwddw(mu, lam, Xn, np.zeros((4, 3)))
    E = orc.emat_tet(lam, mu, *Xn).reshape(4, 4, 3, 3).transpose(0, 1, 3, 2)
    assert np.abs(ddw0.reshape(4, 4, 3, 3) - E).max() <= 1e-12 * np.abs(E).max()


def test_jacobi_pcg_3x3_matches_an_independent_scipy_implementation(orc):
    """SURVEY A.6 + A.7 (DERIVED for 3x3 blocks and point Jacobi): textbook PCG written on a scipy BSR matrix, r updated
    recursively, convergence on ||r_k|| / ||r_0|| — same iteration count and iterates as the oracle's restatement"""
    import scipy.sparse as sp
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import oracle_elasticity_problem
    prob = oracle_elasticity_problem(orc, 8)
    r2i, i2c, nv = prob["r2i"].astype(np.int64), prob["i2c"].astype(np.int64), prob["nv"]
    # scipy wants row-major blocks and the diagonal inside the pattern
    off = sp.bsr_matrix((prob["iv"].reshape(-1, 3, 3).transpose(0, 2, 1), i2c, r2i), shape=(3 * nv, 3 * nv))
    dia = sp.bsr_matrix((prob["rv"].reshape(-1, 3, 3).transpose(0, 2, 1), np.arange(nv), np.arange(nv + 1)), shape=(3 * nv, 3 * nv))
    A = (off.tocsr() + dia).tocsr()
    b = prob["rhs"].reshape(-1)
    minv = 1.0 / A.diagonal()
    tol, max_it = 1e-8, 5000
    x = np.zeros_like(b)
    r = b.copy()
    rr0 = r @ r
    z = minv * r
    p = z.copy()
    rho = r @ z
    hist = [np.sqrt(rr0)]
    for _ in range(max_it):
        q = A @ p
        alpha = rho / (p @ q)
        r -= alpha * q
        x += alpha * p
        s = r @ r
        hist.append(np.sqrt(s))
        if np.sqrt(s / rr0) < tol:
            break
        z = minv * r
        rho_new = r @ z
        p = z + (rho_new / rho) * p
        rho = rho_new
    ilu = orc.Ilu("jacobi", 3, prob["r2i"], prob["i2c"])
    ilu.copy_value(prob["r2i"], prob["i2c"], prob["iv"], prob["rv"])
    ilu.decompose()
    ro = prob["rhs"].copy()
    xo, pro, po = np.zeros_like(ro), np.zeros_like(ro), np.zeros_like(ro)
    hist_o = orc.preconditioned_conjugate_gradient(ro, xo, pro, po, tol, max_it, prob["r2i"], prob["i2c"], prob["iv"], prob["rv"], ilu)
    assert abs(len(hist_o) - len(hist)) <= 1
    m = min(len(hist), len(hist_o)) // 2
    assert np.allclose(hist_o[:m], hist[:m], rtol=1e-8)
    assert np.linalg.norm(xo.reshape(-1) - x) <= 1e-7 * np.linalg.norm(x)
    assert np.linalg.norm(b - A @ x) <= 2e-8 * np.linalg.norm(b)
