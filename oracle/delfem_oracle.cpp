// =============================================================================================
// delfem_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
// Single-threaded fp64 restatement (C++, compiled `g++ -O3 -march=native -ffp-contract=off`) of the
// del-fem hot path: element kernels -> merge into "CSR + separate diagonal" / BSR -> BC -> (P)CG.
// Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
// load this library, and only as the checker / reported CPU baseline.  The product
// (del-fem_b200/csrc) never links, loads or calls anything in this directory.
//
// PARITY STATUS (see DESIGN.md "Oracle"):
//   * The reference cannot be built here (no cargo/rustc) => no oracle/_ref.  This restatement is pinned
//     by re-running the reference's own property tests (tests/test_oracle_*.py): energy identity
//     `solid_linear_tri2.rs:164-176`, CG/ILU0/ILU-full iteration bounds `:306-326`, FD protocol of
//     `solid_incompressive_hex.rs:188-231`, the 4-row del-fem-ls pattern `sparse_square.rs:169-170`.
//   * Third-party arithmetic (del-geo-core 0.1.36: tri2::{area,dldx}, tri3::emat_cotangent_laplacian,
//     matn_col_major::*; del-msh-cpu =0.1.42: vtx2vtx::from_uniform_mesh) is NOT under /root/reference and
//     is restated from its mathematical definition + call-site shapes: "parity unpinned" for the
//     within-row column order of the pattern and for the last-bit rounding of those helpers.
//   * DERIVED kernels (tet Laplacian, tet linear solid, tet mass, tet neo-Hookean, scalar/block Jacobi,
//     3x3 PCG) do not exist in the reference: "parity unpinned"; this file is their definition.
//
// All indices are `usize` (uint64_t), all nested Rust arrays are flat row-major, blocks column-major
// (entry (a,b) of an NxN block at a + N*b) exactly as `del_geo_core::matn_col_major`.
// =============================================================================================
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <thread>
#include <vector>

typedef uint64_t usize;
static const usize USIZE_MAX = std::numeric_limits<usize>::max();

#define ORC_API extern "C" __attribute__((visibility("default")))

// error convention of the oracle: the reference panics; we return non-zero.
enum { ORC_OK = 0, ORC_PANIC = 1 };

// ---------------------------------------------------------------------------------------------
// third-party restatements: del_geo_core::{vecn, matn_col_major}
// ---------------------------------------------------------------------------------------------
template <int N>
static inline double vecn_dot(const double *x, const double *y) {
  double s = 0.0;
  for (int i = 0; i < N; ++i) s = s + x[i] * y[i];
  return s;
}

// y = A x, A column-major (a + N*b)
template <int N>
static inline void matn_mult_vec(double *y, const double *a, const double *x) {
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int j = 0; j < N; ++j) s = s + a[i + N * j] * x[j];
    y[i] = s;
  }
}

// C = A B (column-major)
template <int N>
static inline void matn_mult_mat(double *c, const double *a, const double *b) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      for (int k = 0; k < N; ++k) s = s + a[i + N * k] * b[k + N * j];
      c[i + N * j] = s;
    }
}

// Gauss-Jordan inverse with partial pivoting on a column-major NxN block. returns false if singular.
template <int N>
static inline bool matn_try_inverse(double *inv, const double *a) {
  double m[N][2 * N];
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) {
      m[i][j] = a[i + N * j];
      m[i][N + j] = (i == j) ? 1.0 : 0.0;
    }
  }
  for (int c = 0; c < N; ++c) {
    int p = c;
    for (int r = c + 1; r < N; ++r)
      if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
    if (m[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 2 * N; ++j) std::swap(m[p][j], m[c][j]);
    double d = 1.0 / m[c][c];
    for (int j = 0; j < 2 * N; ++j) m[c][j] *= d;
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      double f = m[r][c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * N; ++j) m[r][j] -= f * m[c][j];
    }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) inv[i + N * j] = m[i][N + j];
  return true;
}

// ---------------------------------------------------------------------------------------------
// third-party restatements: del_geo_core::{tri2, tri3}; tet analogue (derived)
// ---------------------------------------------------------------------------------------------
static inline double tri2_area(const double *p0, const double *p1, const double *p2) {
  return 0.5 * ((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]));
}

// dldx[d][i] = dL_i/dx_d  (shape proven by solid_linear_tri2.rs:5,29-31)
static inline void tri2_dldx(double dldx[2][3], const double *p0, const double *p1, const double *p2) {
  const double a = tri2_area(p0, p1, p2);
  const double t = 1.0 / (2.0 * a);
  dldx[0][0] = t * (p1[1] - p2[1]);
  dldx[0][1] = t * (p2[1] - p0[1]);
  dldx[0][2] = t * (p0[1] - p1[1]);
  dldx[1][0] = t * (p2[0] - p1[0]);
  dldx[1][1] = t * (p0[0] - p2[0]);
  dldx[1][2] = t * (p1[0] - p0[0]);
}

// tet: volume and barycentric gradients g[i][d] = dN_i/dx_d  
static inline double tet_vol_grad(double g[4][3], const double *p0, const double *p1, const double *p2,
                                  const double *p3) {
  double d1[3], d2[3], d3[3];
  for (int k = 0; k < 3; ++k) {
    d1[k] = p1[k] - p0[k];
    d2[k] = p2[k] - p0[k];
    d3[k] = p3[k] - p0[k];
  }
  // cofactor rows: c1 = d2 x d3, c2 = d3 x d1, c3 = d1 x d2 ; det = d1 . c1
  double c1[3] = {d2[1] * d3[2] - d2[2] * d3[1], d2[2] * d3[0] - d2[0] * d3[2], d2[0] * d3[1] - d2[1] * d3[0]};
  double c2[3] = {d3[1] * d1[2] - d3[2] * d1[1], d3[2] * d1[0] - d3[0] * d1[2], d3[0] * d1[1] - d3[1] * d1[0]};
  double c3[3] = {d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0]};
  const double det = d1[0] * c1[0] + d1[1] * c1[1] + d1[2] * c1[2];
  const double inv = 1.0 / det;
  for (int k = 0; k < 3; ++k) {
    g[1][k] = c1[k] * inv;
    g[2][k] = c2[k] * inv;
    g[3][k] = c3[k] * inv;
    g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
  }
  return det / 6.0;
}

// ---------------------------------------------------------------------------------------------
// A1  laplace_tri2::ddw_            del-fem-cpu/src/laplace_tri2.rs:3-17
// out: [3][3][1]
// ---------------------------------------------------------------------------------------------
ORC_API void orc_laplace_tri2_ddw(double alpha, const double *p0, const double *p1, const double *p2, double *ddw) {
  const double area = tri2_area(p0, p1, p2);
  double dldx[2][3];
  tri2_dldx(dldx, p0, p1, p2);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ddw[i * 3 + j] = alpha * area * (dldx[0][i] * dldx[0][j] + dldx[1][i] * dldx[1][j]);
}

// A2  del_geo_core::tri3::emat_cotangent_laplacian (third-party, restated; call site laplace_tri3.rs:18)
// E[i][i] = (cot_j + cot_k)/2 , E[i][j] = -cot_k/2
ORC_API void orc_cot_laplace_tri3(const double *p0, const double *p1, const double *p2, double *emat) {
  const double *p[3] = {p0, p1, p2};
  double cot[3];
  for (int k = 0; k < 3; ++k) {
    const double *a = p[k], *b = p[(k + 1) % 3], *c = p[(k + 2) % 3];
    double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double cr[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
    double sn = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
    double cs = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
    cot[k] = cs / sn;
  }
  for (int i = 0; i < 3; ++i) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    emat[i * 3 + i] = 0.5 * (cot[j] + cot[k]);
    emat[i * 3 + j] = -0.5 * cot[k];
    emat[i * 3 + k] = -0.5 * cot[j];
  }
}

// A3  solid_linear_tri2::{add_weighted_emat_2d, emat_tri2}   del-fem-cpu/src/solid_linear_tri2.rs:1-33
// out: [3][3][4], block column-major a + 2b
ORC_API void orc_emat_tri2(double lambda, double myu, const double *p0, const double *p1, const double *p2,
                           double *emat) {
  const double w = tri2_area(p0, p1, p2);
  double dldx[2][3];
  tri2_dldx(dldx, p0, p1, p2);
  for (int q = 0; q < 36; ++q) emat[q] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double *e = emat + (i * 3 + j) * 4;
      e[0] += w * (lambda + myu) * dldx[0][i] * dldx[0][j];
      e[2] += w * (lambda * dldx[0][i] * dldx[1][j] + myu * dldx[0][j] * dldx[1][i]);
      e[1] += w * (lambda * dldx[1][i] * dldx[0][j] + myu * dldx[1][j] * dldx[0][i]);
      e[3] += w * (lambda + myu) * dldx[1][i] * dldx[1][j];
      const double dtmp1 = w * myu * (dldx[1][i] * dldx[1][j] + dldx[0][i] * dldx[0][j]);
      e[0] += dtmp1;
      e[3] += dtmp1;
    }
}

// A1' (DERIVED) tet scalar Laplacian: E[i][j] = alpha V g_i.g_j      generalises laplace_tri2.rs:12-15
ORC_API void orc_laplace_tet_ddw(double alpha, const double *p0, const double *p1, const double *p2,
                                 const double *p3, double *ddw) {
  double g[4][3];
  const double vol = tet_vol_grad(g, p0, p1, p2, p3);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      ddw[i * 4 + j] = alpha * vol * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
}

// A3' (DERIVED) tet linear solid: E[i][j][a+3b] = V (lambda g_i[a] g_j[b] + myu g_j[a] g_i[b] + d_ab myu g_i.g_j)
// generalises solid_linear_tri2.rs:11-19 (index convention) and solid_linear_hex.rs:18-25 (3-D formula)
ORC_API void orc_emat_tet(double lambda, double myu, const double *p0, const double *p1, const double *p2,
                          const double *p3, double *emat) {
  double g[4][3];
  const double vol = tet_vol_grad(g, p0, p1, p2, p3);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double *e = emat + (i * 4 + j) * 9;
      double dtmp1 = 0.0;
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) e[a + 3 * b] = vol * (lambda * g[i][a] * g[j][b] + myu * g[j][a] * g[i][b]);
        dtmp1 += g[i][a] * g[j][a];
      }
      for (int a = 0; a < 3; ++a) e[a + 3 * a] += vol * myu * dtmp1;
    }
}

// A9' (DERIVED) tet volume (for lumped mass rho V / 4 per node, cf. mitc_tri3.rs:324-329)
ORC_API double orc_tet_volume(const double *p0, const double *p1, const double *p2, const double *p3) {
  double g[4][3];
  return tet_vol_grad(g, p0, p1, p2, p3);
}

// A9' (DERIVED) consistent tet mass: block (i,j) = rho V / 20 (1 + delta_ij) I_3, column-major 3x3 blocks.  "parity unpinned".
ORC_API void orc_emat_mass_tet(double rho, const double *p0, const double *p1, const double *p2, const double *p3, double *emat) {
  const double vol = orc_tet_volume(p0, p1, p2, p3);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const double v = rho * vol / 20.0 * (i == j ? 2.0 : 1.0);
      double *e = emat + (i * 4 + j) * 9;
      for (int q = 0; q < 9; ++q) e[q] = (q == 0 || q == 4 || q == 8) ? v : 0.0;
    }
}

// mitc_tri3::mass_lumped_plate_bending   del-fem-cpu/src/mitc_tri3.rs:304-332 (3 dofs per vertex: w, theta_x, theta_y)
ORC_API void orc_mass_lumped_plate_bending(const usize *tri2vtx, usize num_tri, const double *vtx2xy, usize num_vtx, double thick,
                                           double rho, double *vtx2mass) {
  for (usize q = 0; q < num_vtx * 3; ++q) vtx2mass[q] = 0.0;
  for (usize t = 0; t < num_tri; ++t) {
    const usize *nv = tri2vtx + t * 3;
    const double a012 = tri2_area(vtx2xy + nv[0] * 2, vtx2xy + nv[1] * 2, vtx2xy + nv[2] * 2);
    const double m0 = a012 / 3.0 * rho * thick;
    const double m1 = a012 / 3.0 * rho * thick * thick * thick / 12.0;
    for (int n = 0; n < 3; ++n) {
      vtx2mass[nv[n] * 3] += m0;
      vtx2mass[nv[n] * 3 + 1] += m1;
      vtx2mass[nv[n] * 3 + 2] += m1;
    }
  }
}

// A9' (DERIVED) lumped mass of a uniform mesh, same loop shape as mass_lumped_plate_bending (:321-331): every node of an element
// receives measure / n_node * rho.  measure: tri2 area (n_node 3, ndim 2), tri3 area (3, 3), tet volume (4, 3).
ORC_API int orc_lumped_mass(const usize *elem2vtx, usize num_elem, usize n_node, const double *vtx2xyz, usize num_vtx, usize ndim,
                            double rho, double *vtx2mass) {
  for (usize q = 0; q < num_vtx; ++q) vtx2mass[q] = 0.0;
  for (usize e = 0; e < num_elem; ++e) {
    const usize *nv = elem2vtx + e * n_node;
    double meas;
    if (n_node == 3 && ndim == 2) {
      meas = tri2_area(vtx2xyz + nv[0] * 2, vtx2xyz + nv[1] * 2, vtx2xyz + nv[2] * 2);
    } else if (n_node == 3 && ndim == 3) {
      const double *a = vtx2xyz + nv[0] * 3, *b = vtx2xyz + nv[1] * 3, *c = vtx2xyz + nv[2] * 3;
      const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
      meas = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * 0.5;
    } else if (n_node == 4 && ndim == 3) {
      meas = orc_tet_volume(vtx2xyz + nv[0] * 3, vtx2xyz + nv[1] * 3, vtx2xyz + nv[2] * 3, vtx2xyz + nv[3] * 3);
    } else {
      return ORC_PANIC;
    }
    const double m0 = meas / (double)n_node * rho;
    for (usize n = 0; n < n_node; ++n) vtx2mass[nv[n]] += m0;
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------
// A15' (DERIVED) tet neo-Hookean  W = mu/2 (trC - 3) - mu ln J + lambda/2 (ln J)^2   
// energy density + chain rule of solid_incompressive_hex.rs:86-148 with NNODE=4, dndx=g, detwei=V.
// out: w (1), dw [4][3], ddw [4][4][9] with block index a*3+b (reference convention, :124)
// ---------------------------------------------------------------------------------------------
static const int ISTDIM2IJ[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {1, 2}, {2, 0}};  // dudx.rs:36

static void mat3_det_inv(double *det, double inv[3][3], const double c[3][3]) {
  const double d = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1]) - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                   c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
  const double t = 1.0 / d;
  inv[0][0] = t * (c[1][1] * c[2][2] - c[1][2] * c[2][1]);
  inv[0][1] = t * (c[0][2] * c[2][1] - c[0][1] * c[2][2]);
  inv[0][2] = t * (c[0][1] * c[1][2] - c[0][2] * c[1][1]);
  inv[1][0] = t * (c[1][2] * c[2][0] - c[1][0] * c[2][2]);
  inv[1][1] = t * (c[0][0] * c[2][2] - c[0][2] * c[2][0]);
  inv[1][2] = t * (c[0][2] * c[1][0] - c[0][0] * c[1][2]);
  inv[2][0] = t * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
  inv[2][1] = t * (c[0][1] * c[2][0] - c[0][0] * c[2][1]);
  inv[2][2] = t * (c[0][0] * c[1][1] - c[0][1] * c[1][0]);
  *det = d;
}

// (W, dW/dC [6], d2W/dC2 [6][6]) for compressible neo-Hookean (log-J form), Voigt order ISTDIM2IJ,
// second derivative symmetrised as the reference does at solid_incompressive_hex.rs:44-48
ORC_API void orc_wr_dwrdc_ddwrddc_neohookean(double myu, double lambda, const double *c_flat, double *wr,
                                             double *dwrdc, double *ddwrddc) {
  double c[3][3], ci[3][3], detc;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i][j] = c_flat[i * 3 + j];
  mat3_det_inv(&detc, ci, c);
  const double lnj = 0.5 * std::log(detc);
  *wr = 0.5 * myu * (c[0][0] + c[1][1] + c[2][2] - 3.0) - myu * lnj + 0.5 * lambda * lnj * lnj;
  const double s = 0.5 * lambda * lnj - 0.5 * myu;  // coefficient of C^-1
  for (int g = 0; g < 6; ++g) {
    const int i = ISTDIM2IJ[g][0], j = ISTDIM2IJ[g][1];
    dwrdc[g] = ((i == j) ? 0.5 * myu : 0.0) + s * ci[i][j];
  }
  const double t = 0.5 * myu - 0.5 * lambda * lnj;
  for (int g = 0; g < 6; ++g)
    for (int h = 0; h < 6; ++h) {
      const int i = ISTDIM2IJ[g][0], j = ISTDIM2IJ[g][1], k = ISTDIM2IJ[h][0], l = ISTDIM2IJ[h][1];
      const double v1 = 0.25 * lambda * ci[i][j] * ci[k][l];
      const double v2 = 0.5 * t * ci[i][l] * ci[j][k];
      const double v3 = 0.5 * t * ci[i][k] * ci[j][l];
      ddwrddc[g * 6 + h] = v1 + v2 + v3;
    }
}

// generic chain rule C -> nodal gradient / Hessian (solid_incompressive_hex.rs:86-148) for NNO nodes
template <int NNO>
static void add_wdwddw_from_energy_density_cauchy(double *w, double *dwdx /*[NNO][3]*/, double *ddwddx /*[NNO][NNO][9]*/,
                                                  const double dudx[3][3], const double (*dndx)[3], double wr,
                                                  const double *dwrdc, const double *ddwrddc, double detwei) {
  double dcdu0[NNO][3][6];
  double z[3][3];  // dudx::def_grad_tensor  (dudx.rs:1-14)
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) z[i][j] = dudx[i][j];
    z[i][i] += 1.0;
  }
  for (int ino = 0; ino < NNO; ++ino)
    for (int idim = 0; idim < 3; ++idim) {
      dcdu0[ino][idim][0] = dndx[ino][0] * z[idim][0];
      dcdu0[ino][idim][1] = dndx[ino][1] * z[idim][1];
      dcdu0[ino][idim][2] = dndx[ino][2] * z[idim][2];
      dcdu0[ino][idim][3] = dndx[ino][0] * z[idim][1] + dndx[ino][1] * z[idim][0];
      dcdu0[ino][idim][4] = dndx[ino][1] * z[idim][2] + dndx[ino][2] * z[idim][1];
      dcdu0[ino][idim][5] = dndx[ino][2] * z[idim][0] + dndx[ino][0] * z[idim][2];
    }
  for (int ino = 0; ino < NNO; ++ino)
    for (int jno = 0; jno < NNO; ++jno) {
      double *blk = ddwddx + (ino * NNO + jno) * 9;
      for (int idim = 0; idim < 3; ++idim)
        for (int jdim = 0; jdim < 3; ++jdim) {
          double dtmp1 = 0.0;
          for (int g = 0; g < 6; ++g)
            for (int h = 0; h < 6; ++h) dtmp1 += 2.0 * ddwrddc[g * 6 + h] * dcdu0[ino][idim][g] * dcdu0[jno][jdim][h];
          blk[idim * 3 + jdim] += 2.0 * detwei * dtmp1;
        }
      double dtmp2 = 0.0;
      dtmp2 += dwrdc[0] * dndx[ino][0] * dndx[jno][0];
      dtmp2 += dwrdc[1] * dndx[ino][1] * dndx[jno][1];
      dtmp2 += dwrdc[2] * dndx[ino][2] * dndx[jno][2];
      dtmp2 += dwrdc[3] * (dndx[ino][0] * dndx[jno][1] + dndx[ino][1] * dndx[jno][0]);
      dtmp2 += dwrdc[4] * (dndx[ino][1] * dndx[jno][2] + dndx[ino][2] * dndx[jno][1]);
      dtmp2 += dwrdc[5] * (dndx[ino][2] * dndx[jno][0] + dndx[ino][0] * dndx[jno][2]);
      for (int idim = 0; idim < 3; ++idim) blk[idim * 3 + idim] += 2.0 * detwei * dtmp2;
    }
  for (int ino = 0; ino < NNO; ++ino)
    for (int idim = 0; idim < 3; ++idim) {
      double dtmp1 = 0.0;
      for (int g = 0; g < 6; ++g) dtmp1 += dcdu0[ino][idim][g] * dwrdc[g];
      dwdx[ino * 3 + idim] += 2.0 * detwei * dtmp1;
    }
  *w += wr * detwei;
}

ORC_API void orc_neohookean_tet_wdwddw(double myu, double lambda, const double *node2xyz /*[4][3]*/,
                                       const double *node2disp /*[4][3]*/, double *w, double *dw, double *ddw) {
  double g[4][3];
  const double vol = tet_vol_grad(g, node2xyz, node2xyz + 3, node2xyz + 6, node2xyz + 9);
  // dndx::disp_grad_tensor (dndx.rs:2-21): dudx[i][j] = sum_n disp[n][i] dndx[n][j]
  double dudx[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
      for (int n = 0; n < 4; ++n) t += node2disp[n * 3 + i] * g[n][j];
      dudx[i][j] = t;
    }
  // dudx::right_cauchy_green_tensor (dudx.rs:17-34)
  double c[9];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double t = dudx[i][j] + dudx[j][i];
      for (int k = 0; k < 3; ++k) t += dudx[k][i] * dudx[k][j];
      c[i * 3 + j] = t;
    }
    c[i * 3 + i] += 1.0;
  }
  double wr, dwrdc[6], ddwrddc[36];
  orc_wr_dwrdc_ddwrddc_neohookean(myu, lambda, c, &wr, dwrdc, ddwrddc);
  *w = 0.0;
  for (int q = 0; q < 12; ++q) dw[q] = 0.0;
  for (int q = 0; q < 144; ++q) ddw[q] = 0.0;
  add_wdwddw_from_energy_density_cauchy<4>(w, dw, ddw, dudx, g, wr, dwrdc, ddwrddc, vol);
}

// spring3::wdwddw_squared_length_difference   del-fem-cpu/src/spring3.rs:1-28
// out: w, dw [2][3], ddw [2][2][9] (column-major 3x3; the matrix is symmetric)
ORC_API void orc_spring3_wdwddw(double stiffness, const double *node2xyz_def /*[2][3]*/, double edge_length_ini,
                                double *w, double *dw, double *ddw) {
  double v[3] = {node2xyz_def[0] - node2xyz_def[3], node2xyz_def[1] - node2xyz_def[4], node2xyz_def[2] - node2xyz_def[5]};
  const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double c = edge_length_ini - l;
  for (int k = 0; k < 3; ++k) {
    dw[k] = v[k] * (-c * stiffness / l);
    dw[3 + k] = v[k] * (c * stiffness / l);
  }
  const double t0 = stiffness * edge_length_ini / (l * l * l);
  const double t1 = stiffness * (l - edge_length_ini) / l;
  double m[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) m[a + 3 * b] = (v[a] * v[b]) * t0 + ((a == b) ? t1 : 0.0);
  for (int q = 0; q < 9; ++q) {
    ddw[0 * 9 + q] = m[q];
    ddw[1 * 9 + q] = -m[q];
    ddw[2 * 9 + q] = -m[q];
    ddw[3 * 9 + q] = m[q];
  }
  *w = 0.5 * stiffness * c * c;
}

static inline void orc_cross3(double *o, const double *a, const double *b) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double orc_dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// stvk_tri3::wdw   del-fem-cpu/src/stvk_tri3.rs:9-98 — 2-D Green-Lagrange strain of a triangle (rest shape in the plane)
// and its derivative w.r.t. the deformed corner positions.  out: w[3], dw[3][3][3] = dw[strain][node][dim].
// del_geo_core::vec3::{cross,norm,scale,normalize} restated from their definitions (source unseen).
ORC_API void orc_stvk_tri3_wdw(const double *p_ini /*[3][2]*/, const double *p_def /*[3][3]*/, double *w, double *dw) {
  const double gd0_x[3] = {p_ini[2] - p_ini[0], p_ini[3] - p_ini[1], 0.0};
  const double gd0_y[3] = {p_ini[4] - p_ini[0], p_ini[5] - p_ini[1], 0.0};
  double gd0_z[3];
  orc_cross3(gd0_z, gd0_x, gd0_y);
  const double area0 = std::sqrt(orc_dot3(gd0_z, gd0_z)) * 0.5;
  const double sc = 1.0 / (area0 * 2.0);
  for (int k = 0; k < 3; ++k) gd0_z[k] *= sc;
  double gu0_x[3], gu0_y[3];
  orc_cross3(gu0_x, gd0_y, gd0_z);
  {
    const double il = 1.0 / std::sqrt(orc_dot3(gu0_x, gu0_x));
    for (int k = 0; k < 3; ++k) gu0_x[k] *= il;
  }
  orc_cross3(gu0_y, gd0_z, gd0_x);
  {
    const double il = 1.0 / std::sqrt(orc_dot3(gu0_y, gu0_y));
    for (int k = 0; k < 3; ++k) gu0_y[k] *= il;
  }
  const double gd1_x[3] = {p_def[3] - p_def[0], p_def[4] - p_def[1], p_def[5] - p_def[2]};
  const double gd1_y[3] = {p_def[6] - p_def[0], p_def[7] - p_def[1], p_def[8] - p_def[2]};
  const double gl[3] = {0.5 * (orc_dot3(gd1_x, gd1_x) - orc_dot3(gd0_x, gd0_x)),
                        0.5 * (orc_dot3(gd1_y, gd1_y) - orc_dot3(gd0_y, gd0_y)),
                        orc_dot3(gd1_x, gd1_y) - orc_dot3(gd0_x, gd0_y)};
  const double gg[3][3] = {
      {gu0_x[0] * gu0_x[0], gu0_y[0] * gu0_y[0], gu0_x[0] * gu0_y[0]},
      {gu0_x[1] * gu0_x[1], gu0_y[1] * gu0_y[1], gu0_x[1] * gu0_y[1]},
      {2.0 * gu0_x[0] * gu0_x[1], 2.0 * gu0_y[0] * gu0_y[1], gu0_x[0] * gu0_y[1] + gu0_x[1] * gu0_y[0]}};
  for (int s = 0; s < 3; ++s) {
    w[s] = gl[0] * gg[s][0] + gl[1] * gg[s][1] + gl[2] * gg[s][2];
    const double ab[3][2] = {{-(gg[s][0] + gg[s][2]), -(gg[s][1] + gg[s][2])}, {gg[s][0], gg[s][2]}, {gg[s][2], gg[s][1]}};
    for (int n = 0; n < 3; ++n)
      for (int d = 0; d < 3; ++d) dw[(s * 3 + n) * 3 + d] = ab[n][0] * gd1_x[d] + ab[n][1] * gd1_y[d];
  }
}

// stvk_tri3::wdwddw_   del-fem-cpu/src/stvk_tri3.rs:152-330 — St.Venant-Kirchhoff membrane energy of a 3-D triangle,
// gradient dw[3][3] and Hessian ddw[3][3][9] with the reference's index  ddw[ino][jno][idim*3 + jdim].
// del_geo_core::tri3::unit_normal_area (source unseen) = (n/|n|, |n|/2) with n = (p1-p0) x (p2-p0).
ORC_API void orc_stvk_tri3_wdwddw(const double *p0 /*[3][3] rest*/, const double *p1 /*[3][3] deformed*/, double lambda, double myu,
                                  double *w, double *dw, double *ddw) {
  double gd0[3][3];
  for (int k = 0; k < 3; ++k) {
    gd0[0][k] = p0[3 + k] - p0[k];
    gd0[1][k] = p0[6 + k] - p0[k];
  }
  double nrm[3];
  orc_cross3(nrm, gd0[0], gd0[1]);
  const double area0 = std::sqrt(orc_dot3(nrm, nrm)) * 0.5;
  {
    const double inv = 0.5 / area0;
    for (int k = 0; k < 3; ++k) gd0[2][k] = nrm[k] * inv;
  }
  double gu0[2][3];
  orc_cross3(gu0[0], gd0[1], gd0[2]);
  {
    const double inv = 1.0 / orc_dot3(gu0[0], gd0[0]);
    for (int k = 0; k < 3; ++k) gu0[0][k] *= inv;
  }
  orc_cross3(gu0[1], gd0[2], gd0[0]);
  {
    const double inv = 1.0 / orc_dot3(gu0[1], gd0[1]);
    for (int k = 0; k < 3; ++k) gu0[1][k] *= inv;
  }
  double gd[2][3];
  for (int k = 0; k < 3; ++k) {
    gd[0][k] = p1[3 + k] - p1[k];
    gd[1][k] = p1[6 + k] - p1[k];
  }
  const double gl[3] = {0.5 * (orc_dot3(gd[0], gd[0]) - orc_dot3(gd0[0], gd0[0])),
                        0.5 * (orc_dot3(gd[1], gd[1]) - orc_dot3(gd0[1], gd0[1])),
                        1.0 * (orc_dot3(gd[0], gd[1]) - orc_dot3(gd0[0], gd0[1]))};
  const double g[3] = {orc_dot3(gu0[0], gu0[0]), orc_dot3(gu0[1], gu0[1]), orc_dot3(gu0[1], gu0[0])};
  const double E[3][3] = {
      {lambda * g[0] * g[0] + 2.0 * myu * (g[0] * g[0]), lambda * g[0] * g[1] + 2.0 * myu * (g[2] * g[2]),
       lambda * g[0] * g[2] + 2.0 * myu * (g[0] * g[2])},
      {lambda * g[1] * g[0] + 2.0 * myu * (g[2] * g[2]), lambda * g[1] * g[1] + 2.0 * myu * (g[1] * g[1]),
       lambda * g[1] * g[2] + 2.0 * myu * (g[2] * g[1])},
      {lambda * g[2] * g[0] + 2.0 * myu * (g[0] * g[2]), lambda * g[2] * g[1] + 2.0 * myu * (g[2] * g[1]),
       lambda * g[2] * g[2] + 1.0 * myu * (g[0] * g[1] + g[2] * g[2])}};
  double S[3];
  for (int s = 0; s < 3; ++s) S[s] = E[s][0] * gl[0] + E[s][1] * gl[1] + E[s][2] * gl[2];
  *w = 0.5 * area0 * (gl[0] * S[0] + gl[1] * S[1] + gl[2] * S[2]);
  static const double dN[3][2] = {{-1.0, -1.0}, {1.0, 0.0}, {0.0, 1.0}};
  for (int i = 0; i < 3; ++i)
    for (int a = 0; a < 3; ++a)
      dw[i * 3 + a] = area0 * (S[0] * gd[0][a] * dN[i][0] + S[2] * gd[0][a] * dN[i][1] + S[2] * gd[1][a] * dN[i][0] +
                               S[1] * gd[1][a] * dN[i][1]);
  // the 16 material terms of :282-313 in their order: (row edge p, row dN column r, E entry, col edge q, col dN column s)
  static const int T[16][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 1, 1, 1}, {0, 0, 0, 2, 0, 1}, {0, 0, 0, 2, 1, 0},
                               {1, 1, 1, 0, 0, 0}, {1, 1, 1, 1, 1, 1}, {1, 1, 1, 2, 0, 1}, {1, 1, 1, 2, 1, 0},
                               {0, 1, 2, 0, 0, 0}, {0, 1, 2, 1, 1, 1}, {0, 1, 2, 2, 0, 1}, {0, 1, 2, 2, 1, 0},
                               {1, 0, 2, 0, 0, 0}, {1, 0, 2, 1, 1, 1}, {1, 0, 2, 2, 0, 1}, {1, 0, 2, 2, 1, 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double t = 0.0;
          for (int k = 0; k < 16; ++k) t += gd[T[k][0]][a] * dN[i][T[k][1]] * E[T[k][2]][T[k][3]] * gd[T[k][4]][b] * dN[j][T[k][5]];
          ddw[(i * 3 + j) * 9 + a * 3 + b] = t * area0;
        }
      const double t1 = area0 * (S[0] * dN[i][0] * dN[j][0] + S[2] * dN[i][0] * dN[j][1] + S[2] * dN[i][1] * dN[j][0] +
                                 S[1] * dN[i][1] * dN[j][1]);
      ddw[(i * 3 + j) * 9 + 0] += t1;
      ddw[(i * 3 + j) * 9 + 4] += t1;
      ddw[(i * 3 + j) * 9 + 8] += t1;
    }
}

// quadratic_bending::wdw   del-fem-cpu/src/quadratic_bending.rs:1-43 — hinge "bending vector" w[3] = sum_n k_n p_def[n] of the two
// triangles (0,2,3),(1,3,2) sharing edge 2-3, and dwdp[3][4][3] (= k_n on the diagonal).  Also returns k[4].
// del_geo_core::tri3::area = |(b-a)x(c-a)|/2, edge::length = |b-a| (source unseen).
ORC_API void orc_quadratic_bending_wdw(const double *p_ini /*[4][3]*/, const double *p_def /*[4][3]*/, double *w, double *dwdp,
                                       double *k_out) {
  auto sub = [](double *o, const double *a, const double *b) {
    for (int k = 0; k < 3; ++k) o[k] = a[k] - b[k];
  };
  auto tri_area = [&](const double *a, const double *b, const double *c) {
    double u[3], v[3], n[3];
    sub(u, b, a);
    sub(v, c, a);
    orc_cross3(n, u, v);
    return std::sqrt(orc_dot3(n, n)) * 0.5;
  };
  const double *q0 = p_ini, *q1 = p_ini + 3, *q2 = p_ini + 6, *q3 = p_ini + 9;
  const double area0 = tri_area(q0, q2, q3);
  const double area1 = tri_area(q1, q3, q2);
  double e23[3], e02[3], e03[3], e12[3], e13[3];
  sub(e23, q3, q2);
  sub(e02, q2, q0);
  sub(e03, q3, q0);
  sub(e12, q2, q1);
  sub(e13, q3, q1);
  const double len = std::sqrt(orc_dot3(e23, e23));
  const double h0 = 2.0 * area0 / len;
  const double h1 = 2.0 * area1 / len;
  const double cot023 = -orc_dot3(e02, e23) / h0;
  const double cot032 = orc_dot3(e03, e23) / h0;
  const double cot123 = -orc_dot3(e12, e23) / h1;
  const double cot132 = orc_dot3(e13, e23) / h1;
  const double tmp0 = std::sqrt(3.0) / (std::sqrt(area0 + area1) * len);
  const double k[4] = {(-cot023 - cot032) * tmp0, (-cot123 - cot132) * tmp0, (cot032 + cot132) * tmp0, (cot023 + cot123) * tmp0};
  for (int d = 0; d < 3; ++d) w[d] = k[0] * p_def[d] + k[1] * p_def[3 + d] + k[2] * p_def[6 + d] + k[3] * p_def[9 + d];
  if (dwdp) {
    for (int q = 0; q < 36; ++q) dwdp[q] = 0.0;
    for (int d = 0; d < 3; ++d)
      for (int n = 0; n < 4; ++n) dwdp[(d * 4 + n) * 3 + d] = k[n];
  }
  if (k_out)
    for (int n = 0; n < 4; ++n) k_out[n] = k[n];
}

// DERIVED (not in the reference): energy W = stiffness/2 |w|^2 of the hinge vector above.  w is linear in p_def, so the
// Hessian block (i,j) is exactly stiffness k_i k_j I_3 and the gradient stiffness k_i w.  "parity unpinned".
ORC_API void orc_quadratic_bending_wdwddw(double stiffness, const double *p_ini, const double *p_def, double *w, double *dw,
                                          double *ddw /*[4][4][9]*/) {
  double c[3], k[4];
  orc_quadratic_bending_wdw(p_ini, p_def, c, nullptr, k);
  *w = 0.5 * stiffness * (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  for (int n = 0; n < 4; ++n)
    for (int d = 0; d < 3; ++d) dw[n * 3 + d] = (stiffness * k[n]) * c[d];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const double v = (stiffness * k[i]) * k[j];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) ddw[(i * 4 + j) * 9 + a + 3 * b] = (a == b) ? v : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// synthetic meshes — test/bench inputs, not part of the reference
// ---------------------------------------------------------------------------------------------
// Kuhn-6 structured tet grid, vertex id i + nx (j + ny k), coordinates (i,j,k)*h.
ORC_API void orc_kuhn_tet_mesh(usize nx, usize ny, usize nz, double h, usize *elem2vtx, double *vtx2xyz) {
  for (usize k = 0; k < nz; ++k)
    for (usize j = 0; j < ny; ++j)
      for (usize i = 0; i < nx; ++i) {
        const usize v = i + nx * (j + ny * k);
        vtx2xyz[v * 3 + 0] = (double)i * h;
        vtx2xyz[v * 3 + 1] = (double)j * h;
        vtx2xyz[v * 3 + 2] = (double)k * h;
      }
  // axis permutations in lexicographic order and their parity (odd => swap middle two nodes)
  static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  static const int odd[6] = {0, 1, 1, 0, 0, 1};
  const usize stride[3] = {1, nx, nx * ny};
  usize e = 0;
  for (usize k = 0; k + 1 < nz; ++k)
    for (usize j = 0; j + 1 < ny; ++j)
      for (usize i = 0; i + 1 < nx; ++i) {
        const usize o = i + nx * (j + ny * k);
        for (int t = 0; t < 6; ++t, ++e) {
          usize a = o + stride[perm[t][0]];
          usize b = a + stride[perm[t][1]];
          usize c = o + stride[0] + stride[1] + stride[2];
          if (odd[t]) std::swap(a, b);
          elem2vtx[e * 4 + 0] = o;
          elem2vtx[e * 4 + 1] = a;
          elem2vtx[e * 4 + 2] = b;
          elem2vtx[e * 4 + 3] = c;
        }
      }
}

// (n+1)^2 grid on [-1,1]^2, each cell split along the (0,0)-(1,1) diagonal; optional elliptical disk map
ORC_API void orc_grid_tri_mesh(usize n, int map_to_disk, usize *tri2vtx, double *vtx2xy) {
  const usize m = n + 1;
  for (usize j = 0; j < m; ++j)
    for (usize i = 0; i < m; ++i) {
      const double s = -1.0 + 2.0 * (double)i / (double)n, t = -1.0 + 2.0 * (double)j / (double)n;
      double x = s, y = t;
      if (map_to_disk) {
        x = s * std::sqrt(1.0 - 0.5 * t * t);
        y = t * std::sqrt(1.0 - 0.5 * s * s);
      }
      vtx2xy[(i + m * j) * 2 + 0] = x;
      vtx2xy[(i + m * j) * 2 + 1] = y;
    }
  usize e = 0;
  for (usize j = 0; j < n; ++j)
    for (usize i = 0; i < n; ++i) {
      const usize v00 = i + m * j, v10 = v00 + 1, v01 = v00 + m, v11 = v01 + 1;
      tri2vtx[e * 3 + 0] = v00; tri2vtx[e * 3 + 1] = v10; tri2vtx[e * 3 + 2] = v11; ++e;
      tri2vtx[e * 3 + 0] = v00; tri2vtx[e * 3 + 1] = v11; tri2vtx[e * 3 + 2] = v01; ++e;
    }
}

// splitmix64 hashed by index: uniform in [0,1)   (stand-in for ChaCha seed 0 at solid_linear_tri2.rs:120-129)
static inline double splitmix_uniform(usize seed, usize idx) {
  usize z = (seed + idx + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
ORC_API void orc_splitmix_fill(usize seed, usize first_idx, usize n, double lo, double hi, double *out) {
  for (usize i = 0; i < n; ++i) out[i] = lo + (hi - lo) * splitmix_uniform(seed, first_idx + i);
}

// ---------------------------------------------------------------------------------------------
// A.8 pattern: del_msh_cpu::vtx2vtx::from_uniform_mesh (third-party, restated; "parity unpinned" for
// within-row order).  vtx2elem by counting sort (elements ascending per vertex); per vertex: elements
// ascending x nodes in order, first-seen order, stamp array; is_self=false pre-stamps the vertex itself,
// is_self=true emits the vertex the first time it is seen like any other neighbour.
// Two-call protocol: idx2col == NULL => only row2idx is filled; returns nnz.
// ---------------------------------------------------------------------------------------------
ORC_API usize orc_vtx2vtx_from_uniform_mesh(const usize *elem2vtx, usize num_elem, usize num_node, usize num_vtx,
                                            int is_self, usize *row2idx, usize *idx2col) {
  std::vector<usize> v2e_ptr(num_vtx + 1, 0), v2e;
  for (usize q = 0; q < num_elem * num_node; ++q) v2e_ptr[elem2vtx[q] + 1] += 1;
  for (usize v = 0; v < num_vtx; ++v) v2e_ptr[v + 1] += v2e_ptr[v];
  v2e.resize(v2e_ptr[num_vtx]);
  {
    std::vector<usize> cur(v2e_ptr.begin(), v2e_ptr.end() - 1);
    for (usize e = 0; e < num_elem; ++e)
      for (usize n = 0; n < num_node; ++n) v2e[cur[elem2vtx[e * num_node + n]]++] = e;
  }
  std::vector<usize> stamp(num_vtx, USIZE_MAX);
  usize nnz = 0;
  row2idx[0] = 0;
  for (usize v = 0; v < num_vtx; ++v) {
    if (!is_self) stamp[v] = v;
    for (usize q = v2e_ptr[v]; q < v2e_ptr[v + 1]; ++q) {
      const usize e = v2e[q];
      for (usize n = 0; n < num_node; ++n) {
        const usize j = elem2vtx[e * num_node + n];
        if (stamp[j] == v) continue;
        stamp[j] = v;
        if (idx2col) idx2col[nnz] = j;
        ++nnz;
      }
    }
    row2idx[v + 1] = nnz;
  }
  return nnz;
}

// ---------------------------------------------------------------------------------------------
// A5 / A6 / A7 / A7ls  merge (single element)
// ---------------------------------------------------------------------------------------------
// merge::csrdia   del-fem-cpu/src/merge.rs:3-48  (uses only emat[i][j][0]; vertex-id diagonal test)
ORC_API int orc_merge_csrdia(usize blksize, usize n_node, const usize *node2row, const usize *node2col,
                             const double *emat, const usize *row2idx, usize num_blk, const usize *idx2col,
                             usize num_idx, double *row2val, double *idx2val, usize *col2idx /*[num_blk], MAX*/) {
  for (usize in = 0; in < n_node; ++in) {
    const usize i_row = node2row[in];
    if (i_row >= num_blk) return ORC_PANIC;
    for (usize k = row2idx[i_row]; k < row2idx[i_row + 1]; ++k) {
      if (k >= num_idx) return ORC_PANIC;
      col2idx[idx2col[k]] = k;
    }
    int rc = ORC_OK;
    for (usize jn = 0; jn < n_node; ++jn) {
      const usize j_col = node2col[jn];
      if (j_col >= num_blk) { rc = ORC_PANIC; break; }
      const double e = emat[(in * n_node + jn) * blksize];
      if (i_row == j_col) {
        row2val[i_row] += e;
      } else {
        const usize k = col2idx[j_col];
        if (k >= num_idx || idx2col[k] != j_col) { rc = ORC_PANIC; break; }
        idx2val[k] += e;
      }
    }
    for (usize k = row2idx[i_row]; k < row2idx[i_row + 1]; ++k) col2idx[idx2col[k]] = USIZE_MAX;
    if (rc) return rc;
  }
  return ORC_OK;
}

// merge::blkcsr   del-fem-cpu/src/merge.rs:51-88  (diagonal inside the pattern, all BLKSIZE entries)
ORC_API int orc_merge_blkcsr(usize blksize, usize n_node, const usize *node2row, const usize *node2col,
                             const double *emat, const usize *row2idx, usize num_row, const usize *idx2col,
                             usize num_idx, double *idx2val, usize *col2idx) {
  for (usize in = 0; in < n_node; ++in) {
    const usize i_row = node2row[in];
    if (i_row >= num_row) return ORC_PANIC;
    for (usize k = row2idx[i_row]; k < row2idx[i_row + 1]; ++k) {
      if (k >= num_idx) return ORC_PANIC;
      col2idx[idx2col[k]] = k;
    }
    int rc = ORC_OK;
    for (usize jn = 0; jn < n_node; ++jn) {
      const usize j_col = node2col[jn];
      if (j_col >= num_row) { rc = ORC_PANIC; break; }
      const usize k = col2idx[j_col];
      if (k >= num_idx || idx2col[k] != j_col) { rc = ORC_PANIC; break; }
      for (usize q = 0; q < blksize; ++q) idx2val[k * blksize + q] += emat[(in * n_node + jn) * blksize + q];
    }
    for (usize k = row2idx[i_row]; k < row2idx[i_row + 1]; ++k) col2idx[idx2col[k]] = USIZE_MAX;
    if (rc) return rc;
  }
  return ORC_OK;
}

// Matrix::merge_for_array_blk   del-fem-cpu/src/sparse_square.rs:180-214  (node-index diagonal test)
// flavour: 0 = node-index test (cpu `merge_for_array_blk`), 1 = vertex-id test (ls `merge`, sparse_square.rs:66-104)
template <int NN>
static int merge_for_array_blk_t(usize n_node, const usize *node2vtx, const double *emat, const usize *row2idx,
                                 usize num_blk, const usize *idx2col, double *row2val, double *idx2val,
                                 usize *col2idx, int flavour) {
  for (usize in = 0; in < n_node; ++in) {
    const usize i_vtx = node2vtx[in];
    if (i_vtx >= num_blk) return ORC_PANIC;
    for (usize k = row2idx[i_vtx]; k < row2idx[i_vtx + 1]; ++k) col2idx[idx2col[k]] = k;
    int rc = ORC_OK;
    for (usize jn = 0; jn < n_node; ++jn) {
      const double *e = emat + (in * n_node + jn) * NN;
      const bool is_dia = flavour == 0 ? (in == jn) : (i_vtx == node2vtx[jn]);
      if (is_dia) {
        double *d = row2val + i_vtx * NN;
        for (int q = 0; q < NN; ++q) d[q] += e[q];
      } else {
        const usize k = col2idx[node2vtx[jn]];
        if (k == USIZE_MAX) { rc = ORC_PANIC; break; }
        double *d = idx2val + k * NN;
        for (int q = 0; q < NN; ++q) d[q] += e[q];
      }
    }
    for (usize k = row2idx[i_vtx]; k < row2idx[i_vtx + 1]; ++k) col2idx[idx2col[k]] = USIZE_MAX;
    if (rc) return rc;
  }
  return ORC_OK;
}

ORC_API int orc_merge_for_array_blk(usize nn, usize n_node, const usize *node2vtx, const double *emat,
                                    const usize *row2idx, usize num_blk, const usize *idx2col, double *row2val,
                                    double *idx2val, usize *col2idx, int flavour) {
  switch (nn) {
    case 1: return merge_for_array_blk_t<1>(n_node, node2vtx, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx, flavour);
    case 4: return merge_for_array_blk_t<4>(n_node, node2vtx, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx, flavour);
    case 9: return merge_for_array_blk_t<9>(n_node, node2vtx, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx, flavour);
    case 16: return merge_for_array_blk_t<16>(n_node, node2vtx, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx, flavour);
  }
  return ORC_PANIC;
}

// ---------------------------------------------------------------------------------------------
// whole-mesh assembly loops (caller loops of the reference: solid_linear_tri2.rs:144-161, laplace_tri3.rs:13-29,
// del-fem-numpy/src/{laplacian.rs:71-103, linear_solid.rs:17-51}); elements in ascending order.
// kind: 0 tri2 laplace (N=1), 1 tri3 cot-laplace (N=1), 2 tri2 linear solid (N=2),
//       3 tet laplace (N=1), 4 tet linear solid (N=3), 9 consistent tet mass (N=3, p0 = rho; numbered like DFB_ELEM_MASS_TET)
// convention: 0 = csrdia (row2val + idx2val), 1 = blkcsr (diag inside idx2val; row2val ignored)
// ---------------------------------------------------------------------------------------------
ORC_API int orc_assemble(int kind, int convention, const usize *elem2vtx, usize num_elem, const double *vtx2xyz,
                         double p0, double p1, const usize *row2idx, usize num_blk, const usize *idx2col, usize num_idx,
                         double *row2val, double *idx2val) {
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  int n_node, ndim_x, nn;
  switch (kind) {
    case 0: n_node = 3; ndim_x = 2; nn = 1; break;
    case 1: n_node = 3; ndim_x = 3; nn = 1; break;
    case 2: n_node = 3; ndim_x = 2; nn = 4; break;
    case 3: n_node = 4; ndim_x = 3; nn = 1; break;
    case 4: n_node = 4; ndim_x = 3; nn = 9; break;
    case 9: n_node = 4; ndim_x = 3; nn = 9; break;
    default: return ORC_PANIC;
  }
  double emat[16 * 9];
  for (usize e = 0; e < num_elem; ++e) {
    const usize *nv = elem2vtx + e * n_node;
    const double *p[4];
    for (int n = 0; n < n_node; ++n) {
      if (nv[n] >= num_blk) return ORC_PANIC;
      p[n] = vtx2xyz + nv[n] * ndim_x;
    }
    switch (kind) {
      case 0: orc_laplace_tri2_ddw(p0, p[0], p[1], p[2], emat); break;
      case 1: orc_cot_laplace_tri3(p[0], p[1], p[2], emat); break;
      case 2: orc_emat_tri2(p0, p1, p[0], p[1], p[2], emat); break;
      case 3: orc_laplace_tet_ddw(p0, p[0], p[1], p[2], p[3], emat); break;
      case 4: orc_emat_tet(p0, p1, p[0], p[1], p[2], p[3], emat); break;
      case 9: orc_emat_mass_tet(p0, p[0], p[1], p[2], p[3], emat); break;
    }
    int rc;
    if (convention == 1) {
      rc = orc_merge_blkcsr(nn, n_node, nv, nv, emat, row2idx, num_blk, idx2col, num_idx, idx2val, col2idx.data());
    } else if (nn == 1) {
      rc = orc_merge_csrdia(1, n_node, nv, nv, emat, row2idx, num_blk, idx2col, num_idx, row2val, idx2val, col2idx.data());
    } else {
      rc = orc_merge_for_array_blk(nn, n_node, nv, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx.data(), 0);
    }
    if (rc) return rc;
  }
  return ORC_OK;
}

// neo-Hookean whole-mesh: energy, gradient dw[num_vtx][3], Hessian merged (transposed to column-major)
ORC_API int orc_assemble_neohookean_tet(const usize *elem2vtx, usize num_elem, const double *vtx2xyz,
                                        const double *vtx2disp, double myu, double lambda, const usize *row2idx,
                                        usize num_blk, const usize *idx2col, double *row2val, double *idx2val,
                                        double *dw_out, double *w_out) {
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  double wsum = 0.0;
  for (usize e = 0; e < num_elem; ++e) {
    const usize *nv = elem2vtx + e * 4;
    double xyz[12], dsp[12], w, dw[12], ddw[144], emat[144];
    for (int n = 0; n < 4; ++n)
      for (int k = 0; k < 3; ++k) {
        xyz[n * 3 + k] = vtx2xyz[nv[n] * 3 + k];
        dsp[n * 3 + k] = vtx2disp[nv[n] * 3 + k];
      }
    orc_neohookean_tet_wdwddw(myu, lambda, xyz, dsp, &w, dw, ddw);
    wsum += w;
    for (int n = 0; n < 4; ++n)
      for (int k = 0; k < 3; ++k) dw_out[nv[n] * 3 + k] += dw[n * 3 + k];
    // a*3+b  ->  a + 3b (explicit transpose into the column-major BSR convention)
    for (int ij = 0; ij < 16; ++ij)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) emat[ij * 9 + a + 3 * b] = ddw[ij * 9 + a * 3 + b];
    int rc = orc_merge_for_array_blk(9, 4, nv, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx.data(), 0);
    if (rc) return rc;
  }
  *w_out = wsum;
  return ORC_OK;
}

// mass-spring whole-mesh loop: the edge part of wdwddw_hair_system (del-fem-cpu/src/rod3_darboux.rs:760-793) with 3x3 blocks:
// per edge spring3::wdwddw_squared_length_difference -> w += , dw[node] += , ddw.merge_for_array_blk. Rest positions = vtx2xyz,
// deformed = vtx2xyz + vtx2disp, rest length = |rest_0 - rest_1|.
ORC_API int orc_assemble_spring3(const usize *edge2vtx, usize num_edge, const double *vtx2xyz, const double *vtx2disp,
                                 double stiffness, const usize *row2idx, usize num_blk, const usize *idx2col, double *row2val,
                                 double *idx2val, double *dw_out, double *w_out) {
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  double wsum = 0.0;
  for (usize e = 0; e < num_edge; ++e) {
    const usize *nv = edge2vtx + e * 2;
    double def[6], r[3];
    for (int k = 0; k < 3; ++k) {
      const double a0 = vtx2xyz[nv[0] * 3 + k], a1 = vtx2xyz[nv[1] * 3 + k];
      r[k] = a0 - a1;
      def[k] = a0 + vtx2disp[nv[0] * 3 + k];
      def[3 + k] = a1 + vtx2disp[nv[1] * 3 + k];
    }
    const double l0 = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    double w, dw[6], ddw[36];
    orc_spring3_wdwddw(stiffness, def, l0, &w, dw, ddw);
    wsum += w;
    for (int n = 0; n < 2; ++n)
      for (int k = 0; k < 3; ++k) dw_out[nv[n] * 3 + k] += dw[n * 3 + k];
    int rc = orc_merge_for_array_blk(9, 2, nv, ddw, row2idx, num_blk, idx2col, row2val, idx2val, col2idx.data(), 0);
    if (rc) return rc;
  }
  *w_out = wsum;
  return ORC_OK;
}

// cloth membrane / hinge whole-mesh loops (the "remaining element kernels" of SURVEY §8(f)-4; same loop shape as above):
// kind 0: stvk_tri3::wdwddw_ per triangle (p0 = lambda, p1 = myu), Hessian transposed a*3+b -> a+3b when merged;
// kind 1: DERIVED quadratic-bending energy per hinge quad (p0 = stiffness).
// Rest positions = vtx2xyz, deformed = vtx2xyz + vtx2disp.
ORC_API int orc_assemble_cloth(int kind, const usize *elem2vtx, usize num_elem, const double *vtx2xyz, const double *vtx2disp,
                               double p0, double p1, const usize *row2idx, usize num_blk, const usize *idx2col, double *row2val,
                               double *idx2val, double *dw_out, double *w_out) {
  if (kind != 0 && kind != 1) return ORC_PANIC;
  const int nn = kind == 0 ? 3 : 4;
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  double wsum = 0.0;
  for (usize e = 0; e < num_elem; ++e) {
    const usize *nv = elem2vtx + e * nn;
    double rest[12], def[12], w, dw[12], ddw[144], emat[144];
    for (int n = 0; n < nn; ++n)
      for (int k = 0; k < 3; ++k) {
        rest[n * 3 + k] = vtx2xyz[nv[n] * 3 + k];
        def[n * 3 + k] = vtx2xyz[nv[n] * 3 + k] + vtx2disp[nv[n] * 3 + k];
      }
    if (kind == 0) {
      orc_stvk_tri3_wdwddw(rest, def, p0, p1, &w, dw, ddw);
      for (int ij = 0; ij < 9; ++ij)
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) emat[ij * 9 + a + 3 * b] = ddw[ij * 9 + a * 3 + b];
    } else {
      orc_quadratic_bending_wdwddw(p0, rest, def, &w, dw, emat);
    }
    wsum += w;
    for (int n = 0; n < nn; ++n)
      for (int k = 0; k < 3; ++k) dw_out[nv[n] * 3 + k] += dw[n * 3 + k];
    int rc = orc_merge_for_array_blk(9, nn, nv, emat, row2idx, num_blk, idx2col, row2val, idx2val, col2idx.data(), 0);
    if (rc) return rc;
  }
  *w_out = wsum;
  return ORC_OK;
}

// element -> nonzero index map (the artefact the GPU assembly plan must reproduce bit-exactly):
// slot of (e,i,j): off-diagonal -> index into idx2val; diagonal (convention 0) -> num_idx + row.
// flavour as in merge_for_array_blk (0 node-index diagonal test, 1 vertex-id test).
ORC_API int orc_elem2nz(int convention, int flavour, const usize *elem2vtx, usize num_elem, usize n_node,
                        const usize *row2idx, usize num_blk, const usize *idx2col, usize num_idx, uint32_t *elem2nz) {
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  for (usize e = 0; e < num_elem; ++e) {
    const usize *nv = elem2vtx + e * n_node;
    for (usize in = 0; in < n_node; ++in) {
      const usize r = nv[in];
      if (r >= num_blk) return ORC_PANIC;
      for (usize k = row2idx[r]; k < row2idx[r + 1]; ++k) col2idx[idx2col[k]] = k;
      int rc = ORC_OK;
      for (usize jn = 0; jn < n_node; ++jn) {
        const bool is_dia = convention == 0 && (flavour == 0 ? in == jn : r == nv[jn]);
        usize slot;
        if (is_dia) slot = num_idx + r;
        else {
          slot = col2idx[nv[jn]];
          if (slot == USIZE_MAX) { rc = ORC_PANIC; break; }
        }
        elem2nz[(e * n_node + in) * n_node + jn] = (uint32_t)slot;
      }
      for (usize k = row2idx[r]; k < row2idx[r + 1]; ++k) col2idx[idx2col[k]] = USIZE_MAX;
      if (rc) return rc;
    }
  }
  return ORC_OK;
}

// ---------------------------------------------------------------------------------------------
// A8  set_fixed_bc / set_fix_dof_to_rhs_vector     del-fem-cpu/src/sparse_square.rs:3-70
// ---------------------------------------------------------------------------------------------
ORC_API void orc_set_fixed_bc(usize n, double val_dia, const int32_t *bc_flag, double *row2val, double *idx2val,
                              const usize *row2idx, usize num_blk, const usize *idx2col, usize num_idx) {
  const usize nn = n * n;
  for (usize ib = 0; ib < num_blk; ++ib)
    for (usize id = 0; id < n; ++id) {
      if (bc_flag[ib * n + id] == 0) continue;
      for (usize jd = 0; jd < n; ++jd) {
        row2val[ib * nn + id + n * jd] = 0.0;
        row2val[ib * nn + jd + n * id] = 0.0;
      }
      row2val[ib * nn + id + n * id] = val_dia;
    }
  for (usize ib = 0; ib < num_blk; ++ib)
    for (usize k = row2idx[ib]; k < row2idx[ib + 1]; ++k)
      for (usize id = 0; id < n; ++id) {
        if (bc_flag[ib * n + id] == 0) continue;
        for (usize jd = 0; jd < n; ++jd) idx2val[k * nn + id + n * jd] = 0.0;
      }
  for (usize k = 0; k < num_idx; ++k) {
    const usize jb = idx2col[k];
    for (usize jd = 0; jd < n; ++jd) {
      if (bc_flag[jb * n + jd] == 0) continue;
      for (usize id = 0; id < n; ++id) idx2val[k * nn + id + n * jd] = 0.0;
    }
  }
}

ORC_API void orc_set_fix_dof_to_rhs_vector(usize n, double *rhs, const int32_t *flag, usize num_blk) {
  for (usize q = 0; q < num_blk * n; ++q)
    if (flag[q] != 0) rhs[q] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// A9  mult_vec     del-fem-cpu/src/sparse_square.rs:419-447 ; A10 slice_of_array.rs:1-47
// ---------------------------------------------------------------------------------------------
template <int N>
static void mult_vec_t(double *y, double beta, double alpha, const double *x, const usize *row2idx, usize num_blk,
                       const usize *idx2col, const double *idx2val, const double *row2val) {
  const int NN = N * N;
  for (usize q = 0; q < num_blk * N; ++q) y[q] = y[q] * beta;
  for (usize i = 0; i < num_blk; ++i) {
    double a[N];
    for (usize k = row2idx[i]; k < row2idx[i + 1]; ++k) {
      matn_mult_vec<N>(a, idx2val + k * NN, x + idx2col[k] * N);
      for (int d = 0; d < N; ++d) y[i * N + d] += a[d] * alpha;
    }
    matn_mult_vec<N>(a, row2val + i * NN, x + i * N);
    for (int d = 0; d < N; ++d) y[i * N + d] += a[d] * alpha;
  }
}

ORC_API int orc_mult_vec(usize n, double *y, double beta, double alpha, const double *x, const usize *row2idx,
                         usize num_blk, const usize *idx2col, const double *idx2val, const double *row2val) {
  switch (n) {
    case 1: mult_vec_t<1>(y, beta, alpha, x, row2idx, num_blk, idx2col, idx2val, row2val); return ORC_OK;
    case 2: mult_vec_t<2>(y, beta, alpha, x, row2idx, num_blk, idx2col, idx2val, row2val); return ORC_OK;
    case 3: mult_vec_t<3>(y, beta, alpha, x, row2idx, num_blk, idx2col, idx2val, row2val); return ORC_OK;
    case 4: mult_vec_t<4>(y, beta, alpha, x, row2idx, num_blk, idx2col, idx2val, row2val); return ORC_OK;
  }
  return ORC_PANIC;
}

// del-fem-ls mult_mat (scalar matrix applied to num_dim interleaved components)  del-fem-ls/src/sparse_square.rs:134-164
ORC_API void orc_ls_mult_mat(usize num_dim, double *y, double beta, double alpha, const double *x, const usize *row2idx,
                             usize num_row, const usize *idx2col, const double *idx2val, const double *row2val) {
  for (usize q = 0; q < num_row * num_dim; ++q) y[q] *= beta;
  for (usize i = 0; i < num_row; ++i) {
    for (usize k = row2idx[i]; k < row2idx[i + 1]; ++k) {
      const usize j = idx2col[k];
      for (usize d = 0; d < num_dim; ++d) y[i * num_dim + d] += alpha * idx2val[k] * x[j * num_dim + d];
    }
    for (usize d = 0; d < num_dim; ++d) y[i * num_dim + d] += alpha * row2val[i] * x[i * num_dim + d];
  }
}

// slice_of_array::dot — strict left fold over blocks of an inner N-dot (slice_of_array.rs:8-16)
template <int N>
static double dot_t(const double *a, const double *b, usize num_blk) {
  double s = 0.0;
  for (usize i = 0; i < num_blk; ++i) s = s + vecn_dot<N>(a + i * N, b + i * N);
  return s;
}
static double dot_n(usize n, const double *a, const double *b, usize num_blk) {
  switch (n) {
    case 1: return dot_t<1>(a, b, num_blk);
    case 2: return dot_t<2>(a, b, num_blk);
    case 3: return dot_t<3>(a, b, num_blk);
    case 4: return dot_t<4>(a, b, num_blk);
  }
  return std::nan("");
}
ORC_API double orc_dot(usize n, const double *a, const double *b, usize num_blk) { return dot_n(n, a, b, num_blk); }

static void add_scaled_vector(double *u, double alpha, const double *p, usize len) {
  for (usize q = 0; q < len; ++q) u[q] += p[q] * alpha;
}
static void scale_and_add_vec(double *p, double beta, const double *r, usize len) {
  for (usize q = 0; q < len; ++q) p[q] = r[q] + p[q] * beta;
}

// ---------------------------------------------------------------------------------------------
// A11  conjugate_gradient   del-fem-cpu/src/sparse_square.rs:218-261 (eps_sq = DBL_EPSILON) /
//                           del-fem-ls/src/solver_sparse.rs:5-70      (eps_sq = 1e-20, asserts pAp>=0)
// returns history length (ratios, k=0 excluded); -1 on ls-flavour pAp<0 panic
// ---------------------------------------------------------------------------------------------
ORC_API int64_t orc_conjugate_gradient(usize n, double *r, double *u, double *ap, double *p, double tol, usize max_it,
                                       const usize *row2idx, usize num_blk, const usize *idx2col, const double *idx2val,
                                       const double *row2val, double eps_sq, int assert_pap, double *hist) {
  const usize len = num_blk * n;
  usize nh = 0;
  for (usize q = 0; q < len; ++q) u[q] = 0.0;
  double sqnorm_res = dot_n(n, r, r, num_blk);
  if (sqnorm_res < eps_sq) return 0;
  const double inv_sqnorm_res_ini = 1.0 / sqnorm_res;
  std::memcpy(p, r, len * sizeof(double));
  for (usize it = 0; it < max_it; ++it) {
    orc_mult_vec(n, ap, 0.0, 1.0, p, row2idx, num_blk, idx2col, idx2val, row2val);
    const double pap = dot_n(n, p, ap, num_blk);
    if (assert_pap && !(pap >= 0.0)) return -1;
    const double alpha = sqnorm_res / pap;
    add_scaled_vector(u, alpha, p, len);
    add_scaled_vector(r, -alpha, ap, len);
    const double sqnorm_res_new = dot_n(n, r, r, num_blk);
    const double conv_ratio = std::sqrt(sqnorm_res_new * inv_sqnorm_res_ini);
    hist[nh++] = conv_ratio;
    if (conv_ratio < tol) return (int64_t)nh;
    const double beta = sqnorm_res_new / sqnorm_res;
    sqnorm_res = sqnorm_res_new;
    scale_and_add_vec(p, beta, r, len);
  }
  return (int64_t)nh;
}

// ---------------------------------------------------------------------------------------------
// A13  sparse_ilu::Preconditioner      del-fem-cpu/src/sparse_ilu.rs:3-219
// ---------------------------------------------------------------------------------------------
struct OrcIlu {
  usize n;  // block dim
  usize num_blk;
  std::vector<usize> row2idx, idx2col, row2idx_dia;
  std::vector<double> idx2val, row2val;
  int point_jacobi;  // DERIVED: scalar Jacobi (A.7 last paragraph): z[i][d] = r[i][d] / D[i][d + N d]
};

static void ilu_set_dia(OrcIlu *s) {
  s->row2idx_dia.assign(s->num_blk, 0);
  for (usize i = 0; i < s->num_blk; ++i) {
    s->row2idx_dia[i] = s->row2idx[i + 1];
    for (usize k = s->row2idx[i]; k < s->row2idx[i + 1]; ++k)
      if (s->idx2col[k] > i) {
        s->row2idx_dia[i] = k;
        break;
      }
  }
}

// symbolic_iluk   del-fem-cpu/src/sparse_ilu.rs:221-312 — level-of-fill pattern: row i starts from A's (sorted) columns at
// level 0; for every strictly-lower column k taken in ascending order (std BTreeSet), the strictly-upper part of the already
// finished row k whose level allows it (lev_ik + 1 <= lev_fill and lev_kj + 1 <= lev_fill) is merged in with level
// max(lev_ik, lev_kj) + 1 (minimum over all paths); the diagonal is never stored.
static void symbolic_iluk(const usize *a_row2idx, const usize *a_idx2col, usize num_row, usize lev_fill, std::vector<usize> &row2idx,
                          std::vector<usize> &idx2col, std::vector<usize> &row2idx_dia) {
  row2idx.assign(num_row + 1, 0);
  row2idx_dia.assign(num_row, 0);
  std::vector<usize> col, lev;  // idx2collev
  for (usize i = 0; i < num_row; ++i) {
    std::map<usize, usize> col2lev;
    std::set<usize> que;
    for (usize k = a_row2idx[i]; k < a_row2idx[i + 1]; ++k) {
      col2lev[a_idx2col[k]] = 0;  // BTreeMap::insert overwrites: duplicates collapse
      if (a_idx2col[k] < i) que.insert(a_idx2col[k]);
    }
    while (!que.empty()) {
      const usize k = *que.begin();
      que.erase(que.begin());
      const usize ik_lev = col2lev[k];
      if (ik_lev + 1 > lev_fill) continue;
      for (usize kj = row2idx_dia[k]; kj < row2idx[k + 1]; ++kj) {
        const usize kj_lev = lev[kj];
        if (kj_lev + 1 > lev_fill) continue;
        const usize j = col[kj];
        if (j == i) continue;
        const usize mx = ik_lev > kj_lev ? ik_lev : kj_lev;
        if (j < i) que.insert(j);
        auto it = col2lev.find(j);
        if (it == col2lev.end()) col2lev[j] = mx + 1;
        else if (mx + 1 < it->second) it->second = mx + 1;
      }
    }
    for (auto &cl : col2lev) {
      col.push_back(cl.first);
      lev.push_back(cl.second);
    }
    row2idx[i + 1] = row2idx[i] + col2lev.size();
    row2idx_dia[i] = row2idx[i + 1];
    for (usize q = row2idx[i]; q < row2idx[i + 1]; ++q)
      if (col[q] > i) {
        row2idx_dia[i] = q;
        break;
      }
  }
  idx2col = col;
}

// the pattern alone, for tests and for the device side's own symbolic phase to be compared with
ORC_API usize orc_symbolic_iluk(const usize *a_row2idx, const usize *a_idx2col, usize num_row, usize lev_fill, usize *row2idx_out,
                                usize *idx2col_out, usize cap, usize *row2idx_dia_out) {
  std::vector<usize> r, c, d;
  symbolic_iluk(a_row2idx, a_idx2col, num_row, lev_fill, r, c, d);
  std::copy(r.begin(), r.end(), row2idx_out);
  if (row2idx_dia_out) std::copy(d.begin(), d.end(), row2idx_dia_out);
  if (c.size() <= cap) std::copy(c.begin(), c.end(), idx2col_out);
  return c.size();
}

// mode 0: initialize_ilu0 (:28-56)   mode 1: initialize_full (:66-84)
// mode 2: empty off-diagonal pattern = block-Jacobi   mode 3: point (scalar) Jacobi (DERIVED)
// mode 4 + k: initialize_iluk (:58-64) with lev_fill = k
ORC_API OrcIlu *orc_ilu_new(int mode, usize n, const usize *row2idx, usize num_blk, const usize *idx2col) {
  OrcIlu *s = new OrcIlu();
  s->n = n;
  s->num_blk = num_blk;
  s->point_jacobi = (mode == 3);
  if (mode == 0) {
    s->row2idx.assign(row2idx, row2idx + num_blk + 1);
    s->idx2col.assign(idx2col, idx2col + row2idx[num_blk]);
    for (usize i = 0; i < num_blk; ++i) std::sort(s->idx2col.begin() + s->row2idx[i], s->idx2col.begin() + s->row2idx[i + 1]);
  } else if (mode == 1) {
    s->row2idx.resize(num_blk + 1);
    for (usize i = 0; i <= num_blk; ++i) s->row2idx[i] = i * (num_blk - 1);
    s->idx2col.assign(num_blk * (num_blk - 1), 0);
    for (usize i = 0; i < num_blk; ++i) {
      for (usize j = 0; j < i; ++j) s->idx2col[i * (num_blk - 1) + j] = j;
      for (usize j = i + 1; j < num_blk; ++j) s->idx2col[i * (num_blk - 1) + j - 1] = j;
    }
  } else if (mode >= 4) {
    std::vector<usize> dia;
    symbolic_iluk(row2idx, idx2col, num_blk, (usize)(mode - 4), s->row2idx, s->idx2col, dia);
  } else {
    s->row2idx.assign(num_blk + 1, 0);
  }
  ilu_set_dia(s);
  s->idx2val.assign(s->idx2col.size() * n * n, 0.0);
  s->row2val.assign(num_blk * n * n, 0.0);
  return s;
}
ORC_API void orc_ilu_free(OrcIlu *s) { delete s; }

// copy_value (:87-125). returns ORC_PANIC if A has an entry outside the ILU pattern (mode 0/1 only).
ORC_API int orc_ilu_copy_value(OrcIlu *s, const usize *a_row2idx, const usize *a_idx2col, const double *a_idx2val,
                               const double *a_row2val) {
  const usize nn = s->n * s->n;
  std::copy(a_row2val, a_row2val + s->num_blk * nn, s->row2val.begin());
  std::fill(s->idx2val.begin(), s->idx2val.end(), 0.0);
  if (s->idx2col.empty()) return ORC_OK;  // Jacobi flavours: A's off-diagonals are simply not copied
  std::vector<usize> col2idx(s->num_blk, USIZE_MAX);
  for (usize i = 0; i < s->num_blk; ++i) {
    for (usize k = s->row2idx[i]; k < s->row2idx[i + 1]; ++k) col2idx[s->idx2col[k]] = k;
    for (usize ka = a_row2idx[i]; ka < a_row2idx[i + 1]; ++ka) {
      const usize k = col2idx[a_idx2col[ka]];
      if (k == USIZE_MAX) return ORC_PANIC;
      std::copy(a_idx2val + ka * nn, a_idx2val + (ka + 1) * nn, s->idx2val.begin() + k * nn);
    }
    for (usize k = s->row2idx[i]; k < s->row2idx[i + 1]; ++k) col2idx[s->idx2col[k]] = USIZE_MAX;
  }
  return ORC_OK;
}

template <int N>
static int ilu_decompose_t(OrcIlu *s) {
  const int NN = N * N;
  if (s->point_jacobi) {
    for (usize i = 0; i < s->num_blk; ++i) {
      double *d = s->row2val.data() + i * NN;
      double inv[N];
      for (int a = 0; a < N; ++a) {
        if (d[a + N * a] == 0.0) return ORC_PANIC;
        inv[a] = 1.0 / d[a + N * a];
      }
      for (int q = 0; q < NN; ++q) d[q] = 0.0;
      for (int a = 0; a < N; ++a) d[a + N * a] = inv[a];
    }
    return ORC_OK;
  }
  std::vector<usize> col2idx(s->num_blk, USIZE_MAX);
  double ikkj[NN], tmp[NN];
  for (usize i = 0; i < s->num_blk; ++i) {
    for (usize k = s->row2idx[i]; k < s->row2idx[i + 1]; ++k) col2idx[s->idx2col[k]] = k;
    for (usize ik = s->row2idx[i]; ik < s->row2idx_dia[i]; ++ik) {
      const usize kc = s->idx2col[ik];
      double ikv[NN];
      std::copy(s->idx2val.begin() + ik * NN, s->idx2val.begin() + (ik + 1) * NN, ikv);
      for (usize kj = s->row2idx_dia[kc]; kj < s->row2idx[kc + 1]; ++kj) {
        const usize jc = s->idx2col[kj];
        matn_mult_mat<N>(ikkj, ikv, s->idx2val.data() + kj * NN);
        if (jc != i) {
          const usize ij = col2idx[jc];
          if (ij == USIZE_MAX) continue;
          for (int q = 0; q < NN; ++q) s->idx2val[ij * NN + q] -= ikkj[q];
        } else {
          for (int q = 0; q < NN; ++q) s->row2val[i * NN + q] -= ikkj[q];
        }
      }
    }
    if (!matn_try_inverse<N>(tmp, s->row2val.data() + i * NN)) return ORC_PANIC;
    std::copy(tmp, tmp + NN, s->row2val.begin() + i * NN);
    for (usize ij = s->row2idx_dia[i]; ij < s->row2idx[i + 1]; ++ij) {
      matn_mult_mat<N>(tmp, s->row2val.data() + i * NN, s->idx2val.data() + ij * NN);
      std::copy(tmp, tmp + NN, s->idx2val.begin() + ij * NN);
    }
    for (usize k = s->row2idx[i]; k < s->row2idx[i + 1]; ++k) col2idx[s->idx2col[k]] = USIZE_MAX;
  }
  return ORC_OK;
}
// decompose (:127-181)
ORC_API int orc_ilu_decompose(OrcIlu *s) {
  switch (s->n) {
    case 1: return ilu_decompose_t<1>(s);
    case 2: return ilu_decompose_t<2>(s);
    case 3: return ilu_decompose_t<3>(s);
    case 4: return ilu_decompose_t<4>(s);
  }
  return ORC_PANIC;
}

template <int N>
static void ilu_solve_t(double *vec, const OrcIlu *s) {
  const int NN = N * N;
  for (usize i = 0; i < s->num_blk; ++i) {
    double t[N], v[N];
    for (int d = 0; d < N; ++d) t[d] = vec[i * N + d];
    for (usize k = s->row2idx[i]; k < s->row2idx_dia[i]; ++k) {
      matn_mult_vec<N>(v, s->idx2val.data() + k * NN, vec + s->idx2col[k] * N);
      for (int d = 0; d < N; ++d) t[d] -= v[d];
    }
    matn_mult_vec<N>(vec + i * N, s->row2val.data() + i * NN, t);
  }
  for (usize ii = s->num_blk; ii-- > 0;) {
    double t[N], v[N];
    for (int d = 0; d < N; ++d) t[d] = vec[ii * N + d];
    for (usize k = s->row2idx_dia[ii]; k < s->row2idx[ii + 1]; ++k) {
      matn_mult_vec<N>(v, s->idx2val.data() + k * NN, vec + s->idx2col[k] * N);
      for (int d = 0; d < N; ++d) t[d] -= v[d];
    }
    for (int d = 0; d < N; ++d) vec[ii * N + d] = t[d];
  }
}
// solve_preconditioning_vec (:183-219)
ORC_API void orc_ilu_solve(double *vec, const OrcIlu *s) {
  switch (s->n) {
    case 1: ilu_solve_t<1>(vec, s); break;
    case 2: ilu_solve_t<2>(vec, s); break;
    case 3: ilu_solve_t<3>(vec, s); break;
    case 4: ilu_solve_t<4>(vec, s); break;
  }
}
ORC_API const double *orc_ilu_row2val(const OrcIlu *s) { return s->row2val.data(); }

// ---------------------------------------------------------------------------------------------
// A12  preconditioned_conjugate_gradient   del-fem-cpu/src/sparse_square.rs:264-347 (generalised to N)
// history: absolute ||r_k|| including k=0, +1 trailing entry if max_it exhausted. returns history length.
// ---------------------------------------------------------------------------------------------
ORC_API int64_t orc_preconditioned_conjugate_gradient(usize n, double *r, double *x, double *pr, double *p, double tol,
                                                      usize max_it, const usize *row2idx, usize num_blk,
                                                      const usize *idx2col, const double *idx2val,
                                                      const double *row2val, const OrcIlu *ilu, double eps_sq,
                                                      double *hist) {
  const usize len = num_blk * n;
  usize nh = 0;
  for (usize q = 0; q < len; ++q) x[q] = 0.0;
  double inv_sqnorm_res0;
  {
    const double sqnorm_res0 = dot_n(n, r, r, num_blk);
    hist[nh++] = std::sqrt(sqnorm_res0);
    if (sqnorm_res0 < eps_sq) return (int64_t)nh;
    inv_sqnorm_res0 = 1.0 / sqnorm_res0;
  }
  std::memcpy(pr, r, len * sizeof(double));
  orc_ilu_solve(pr, ilu);
  std::memcpy(p, pr, len * sizeof(double));
  double rpr = dot_n(n, r, pr, num_blk);
  for (usize it = 0; it < max_it; ++it) {
    orc_mult_vec(n, pr, 0.0, 1.0, p, row2idx, num_blk, idx2col, idx2val, row2val);
    {
      const double pap = dot_n(n, p, pr, num_blk);
      const double alpha = rpr / pap;
      add_scaled_vector(r, -alpha, pr, len);
      add_scaled_vector(x, alpha, p, len);
    }
    {
      const double sqnorm_res = dot_n(n, r, r, num_blk);
      hist[nh++] = std::sqrt(sqnorm_res);
      const double conv_ratio = std::sqrt(sqnorm_res * inv_sqnorm_res0);
      if (conv_ratio < tol) return (int64_t)nh;
    }
    {
      std::memcpy(pr, r, len * sizeof(double));
      orc_ilu_solve(pr, ilu);
      const double rpr1 = dot_n(n, r, pr, num_blk);
      const double beta = rpr1 / rpr;
      rpr = rpr1;
      scale_and_add_vec(p, beta, pr, len);
    }
  }
  hist[nh++] = std::sqrt(dot_n(n, r, r, num_blk));
  return (int64_t)nh;
}

// ---------------------------------------------------------------------------------------------
// BASELINE-TIMING AID (not a restatement): the same Jacobi-PCG (A12 history semantics) with the rows split over `nthreads`
// std::threads — what a maintainer would get from a row-parallel (rayon-style) version of the reference loop.  The reference
// itself is single-threaded on this path; bench.py reports this beside the faithful 1-thread number.  Dots are summed per
// thread then over threads in thread order, so results differ from the strict left fold in the last bits.
// ---------------------------------------------------------------------------------------------
namespace {
struct SpinBarrier {
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  int n;
  explicit SpinBarrier(int n_) : n(n_) {}
  void wait() {
    const int g = gen.load(std::memory_order_acquire);
    if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
      count.store(0, std::memory_order_relaxed);
      gen.fetch_add(1, std::memory_order_release);
    } else {
      int spins = 0;
      while (gen.load(std::memory_order_acquire) == g)
        if (++spins > 2000) std::this_thread::yield();
    }
  }
};
}  // namespace

ORC_API int64_t orc_pcg_jacobi3_mt(int nthreads, double *r, double *x, double *pr, double *p, double tol, usize max_it,
                                   const usize *row2idx, usize num_blk, const usize *idx2col, const double *idx2val,
                                   const double *row2val, double eps_sq, double *hist) {
  if (nthreads < 1) nthreads = 1;
  const int T = nthreads;
  // row ranges balanced by the number of blocks
  std::vector<usize> rb(T + 1, num_blk);
  rb[0] = 0;
  {
    const usize total = row2idx[num_blk] + num_blk;
    usize i = 0;
    for (int t = 1; t < T; ++t) {
      const usize target = total / T * t;
      while (i < num_blk && row2idx[i] + i < target) ++i;
      rb[t] = i;
    }
  }
  std::vector<double> part(3 * (size_t)T * 8, 0.0);  // padded: one cache line per (thread, slot)
  std::vector<double> dinv(num_blk * 3);
  SpinBarrier bar(T);
  std::atomic<int64_t> nh_out{0};
  auto worker = [&](int t) {
    const usize r0 = rb[t], r1 = rb[t + 1];
    auto reduce = [&](int slot, double v) {
      part[((size_t)slot * T + t) * 8] = v;
      bar.wait();
      double s = 0.0;
      for (int q = 0; q < T; ++q) s += part[((size_t)slot * T + q) * 8];
      return s;
    };
    int64_t nh = 0;
    double a0 = 0.0, a1 = 0.0;
    for (usize i = r0; i < r1; ++i)
      for (int d = 0; d < 3; ++d) {
        const usize q = i * 3 + d;
        dinv[q] = 1.0 / row2val[i * 9 + d + 3 * d];
        x[q] = 0.0;
        a0 += r[q] * r[q];
        pr[q] = dinv[q] * r[q];
        p[q] = pr[q];
        a1 += r[q] * pr[q];
      }
    const double sq0 = reduce(0, a0);
    double rpr = reduce(1, a1);
    if (t == 0) hist[nh] = std::sqrt(sq0);
    ++nh;
    if (sq0 < eps_sq) {
      if (t == 0) nh_out = nh;
      return;
    }
    const double inv0 = 1.0 / sq0;
    for (usize it = 0; it < max_it; ++it) {
      double pap_t = 0.0;
      for (usize i = r0; i < r1; ++i) {
        double y[3] = {0.0, 0.0, 0.0};
        for (usize k = row2idx[i]; k < row2idx[i + 1]; ++k) {
          const double *b = idx2val + k * 9;
          const double *xv = p + idx2col[k] * 3;
          for (int a = 0; a < 3; ++a) y[a] += b[a] * xv[0] + b[a + 3] * xv[1] + b[a + 6] * xv[2];
        }
        const double *b = row2val + i * 9;
        const double *xv = p + i * 3;
        for (int a = 0; a < 3; ++a) {
          y[a] += b[a] * xv[0] + b[a + 3] * xv[1] + b[a + 6] * xv[2];
          pr[i * 3 + a] = y[a];
          pap_t += xv[a] * y[a];
        }
      }
      const double pap = reduce(2, pap_t);
      const double alpha = rpr / pap;
      double rr_t = 0.0, rz_t = 0.0;
      for (usize q = r0 * 3; q < r1 * 3; ++q) {
        r[q] -= alpha * pr[q];
        x[q] += alpha * p[q];
        rr_t += r[q] * r[q];
        pr[q] = dinv[q] * r[q];
        rz_t += r[q] * pr[q];
      }
      const double rr = reduce(0, rr_t);
      const double rz = reduce(1, rz_t);
      if (t == 0) hist[nh] = std::sqrt(rr);
      ++nh;
      if (std::sqrt(rr * inv0) < tol) {
        if (t == 0) nh_out = nh;
        return;
      }
      const double beta = rz / rpr;
      rpr = rz;
      for (usize q = r0 * 3; q < r1 * 3; ++q) p[q] = pr[q] + beta * p[q];
      bar.wait();  // p complete before the next SpMV reads it
    }
    double rr_t = 0.0;
    for (usize q = r0 * 3; q < r1 * 3; ++q) rr_t += r[q] * r[q];
    const double rr = reduce(0, rr_t);
    if (t == 0) {
      hist[nh] = std::sqrt(rr);
      nh_out = nh + 1;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto &q : th) q.join();
  return nh_out.load();
}

// ---------------------------------------------------------------------------------------------
// del-fem-ls sparse_matrix_multiplication.rs
// symbolic_multiplication :7-107 — pattern of C = A * B in discovery order; *_is_square: the matrix has an (unlisted) diagonal,
// c_is_square: the diagonal of C is kept out of the pattern.
// ---------------------------------------------------------------------------------------------
static void symbolic_multiplication(const usize *a_r2i, const usize *a_i2c, usize nrow_a, bool a_sq, const usize *b_r2i,
                                    const usize *b_i2c, bool b_sq, usize b_ncol, bool c_sq, std::vector<usize> &c_r2i,
                                    std::vector<usize> &c_i2c) {
  c_r2i.assign(nrow_a + 1, 0);
  std::vector<usize> flag(b_ncol, USIZE_MAX);
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      for (usize i = 0; i < nrow_a; ++i) c_r2i[i + 1] += c_r2i[i];
      c_i2c.assign(c_r2i[nrow_a], 0);
      std::fill(flag.begin(), flag.end(), USIZE_MAX);
    }
    // pass 1 advances c_r2i[i] while filling and shifts it back afterwards (:62-106)
    for (usize i = 0; i < nrow_a; ++i) {
      auto add = [&](usize j) {
        if (flag[j] == i || (j == i && c_sq)) return;
        if (pass == 0) {
          c_r2i[i + 1] += 1;
        } else {
          c_i2c[c_r2i[i]] = j;
          c_r2i[i] += 1;
        }
        flag[j] = i;
      };
      for (usize ka = a_r2i[i]; ka < a_r2i[i + 1]; ++ka) {
        const usize k = a_i2c[ka];
        for (usize kb = b_r2i[k]; kb < b_r2i[k + 1]; ++kb) add(b_i2c[kb]);
        if (b_sq) add(k);
      }
      if (a_sq) {
        for (usize kb = b_r2i[i]; kb < b_r2i[i + 1]; ++kb) add(b_i2c[kb]);
        if (b_sq) add(i);
      }
    }
  }
  for (usize i = nrow_a; i-- > 1;) c_r2i[i] = c_r2i[i - 1];
  c_r2i[0] = 0;
}

// mult_square_matrices :109-188 — C = M0 * M1 for scalar "CSR + separate diagonal" matrices. Two calls: c_idx2col == NULL
// returns the number of off-diagonal entries (and fills c_row2idx), the second fills everything.
ORC_API usize orc_mult_square_matrices(usize num_blk, const usize *r0, const usize *c0, const double *v0, const double *d0,
                                       const usize *r1, const usize *c1, const double *v1, const double *d1, usize *c_row2idx,
                                       usize *c_idx2col, double *c_idx2val, double *c_row2val) {
  std::vector<usize> r2i, i2c;
  symbolic_multiplication(r0, c0, num_blk, true, r1, c1, true, num_blk, true, r2i, i2c);
  std::copy(r2i.begin(), r2i.end(), c_row2idx);
  if (!c_idx2col) return i2c.size();
  std::copy(i2c.begin(), i2c.end(), c_idx2col);
  std::vector<double> iv(i2c.size(), 0.0), rv(num_blk, 0.0);
  std::vector<usize> col2idx(num_blk, USIZE_MAX);
  for (usize i = 0; i < num_blk; ++i) {
    for (usize q = r2i[i]; q < r2i[i + 1]; ++q) col2idx[i2c[q]] = q;
    for (usize k0 = r0[i]; k0 < r0[i + 1]; ++k0) {
      const usize k = c0[k0];
      for (usize k1 = r1[k]; k1 < r1[k + 1]; ++k1) {
        const usize j = c1[k1];
        if (i == j) rv[i] += v0[k0] * v1[k1];
        else iv[col2idx[j]] += v0[k0] * v1[k1];
      }
      if (i == k) rv[i] += v0[k0] * d1[k];
      else iv[col2idx[k]] += v0[k0] * d1[k];
    }
    for (usize k1 = r1[i]; k1 < r1[i + 1]; ++k1) {
      const usize j = c1[k1];
      if (i == j) rv[i] += d0[i] * v1[k1];
      else iv[col2idx[j]] += d0[i] * v1[k1];
    }
    rv[i] += d0[i] * d1[i];
    for (usize q = r2i[i]; q < r2i[i + 1]; ++q) col2idx[i2c[q]] = USIZE_MAX;
  }
  std::copy(iv.begin(), iv.end(), c_idx2val);
  std::copy(rv.begin(), rv.end(), c_row2val);
  return i2c.size();
}

ORC_API const char *orc_version(void) { return "delfem-oracle 0.1 (restates del-fem 0.1.9 hot path)"; }
