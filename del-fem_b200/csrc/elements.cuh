// elements.cuh — element formulas shared by the assembly translation units (all compiled --fmad=false: the formulas are written
// operation-for-operation like the reference's / the oracle's, and rustc never contracts a*b+c, so linear kernels reproduce the
// serial CPU result bit for bit).
#pragma once
#include "dfb_internal.h"

// ------------------------------------------------------------------------------------------------ element formulas
struct ElemInfo {
  int n_node, ndim, N;
};
__host__ __device__ inline ElemInfo elem_info(int kind) {
  switch (kind) {
    case DFB_ELEM_LAPLACE_TRI2: return {3, 2, 1};
    case DFB_ELEM_COT_LAPLACE_TRI3: return {3, 3, 1};
    case DFB_ELEM_SOLID_LINEAR_TRI2: return {3, 2, 2};
    case DFB_ELEM_LAPLACE_TET: return {4, 3, 1};
    case DFB_ELEM_SOLID_LINEAR_TET: return {4, 3, 3};
    case DFB_ELEM_NEOHOOKEAN_TET: return {4, 3, 3};
    case DFB_ELEM_SPRING3: return {2, 3, 3};
    case DFB_ELEM_STVK_TRI3: return {3, 3, 3};
    case DFB_ELEM_BENDING_QUAD: return {4, 3, 3};
    case DFB_ELEM_MASS_TET: return {4, 3, 3};
  }
  return {0, 0, 0};
}

// del_geo_core::tri2::{area, dldx} (third-party; call sites laplace_tri2.rs:9-10, solid_linear_tri2.rs:28-29)
__device__ __forceinline__ double tri2_area(const double *p0, const double *p1, const double *p2) {
  return 0.5 * ((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]));
}
__device__ __forceinline__ void tri2_dldx(double (&dldx)[2][3], const double *p0, const double *p1, const double *p2) {
  const double a = tri2_area(p0, p1, p2);
  const double t = 1.0 / (2.0 * a);
  dldx[0][0] = t * (p1[1] - p2[1]);
  dldx[0][1] = t * (p2[1] - p0[1]);
  dldx[0][2] = t * (p0[1] - p1[1]);
  dldx[1][0] = t * (p2[0] - p1[0]);
  dldx[1][1] = t * (p0[0] - p2[0]);
  dldx[1][2] = t * (p1[0] - p0[0]);
}

// tet volume + barycentric gradients g[i][d] = dN_i/dx_d
__device__ __forceinline__ double tet_vol_grad(double (&g)[4][3], const double *p0, const double *p1, const double *p2,
                                               const double *p3) {
  double d1[3], d2[3], d3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d1[k] = p1[k] - p0[k];
    d2[k] = p2[k] - p0[k];
    d3[k] = p3[k] - p0[k];
  }
  const double c1[3] = {d2[1] * d3[2] - d2[2] * d3[1], d2[2] * d3[0] - d2[0] * d3[2], d2[0] * d3[1] - d2[1] * d3[0]};
  const double c2[3] = {d3[1] * d1[2] - d3[2] * d1[1], d3[2] * d1[0] - d3[0] * d1[2], d3[0] * d1[1] - d3[1] * d1[0]};
  const double c3[3] = {d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0]};
  const double det = d1[0] * c1[0] + d1[1] * c1[1] + d1[2] * c1[2];
  const double inv = 1.0 / det;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g[1][k] = c1[k] * inv;
    g[2][k] = c2[k] * inv;
    g[3][k] = c3[k] * inv;
    g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
  }
  return det / 6.0;
}

// Per-element precomputation shared by all (i,j) blocks of one element.
template <int KIND> struct ElemGeo;

template <> struct ElemGeo<DFB_ELEM_LAPLACE_TRI2> {  // laplace_tri2::ddw_  laplace_tri2.rs:3-17
  static constexpr int NNODE = 3, NDIM = 2, N = 1;
  double area, dldx[2][3];
  __device__ void init(const double *const *p, double, double) {
    area = tri2_area(p[0], p[1], p[2]);
    tri2_dldx(dldx, p[0], p[1], p[2]);
  }
  __device__ void block(int i, int j, double alpha, double, double *e) const {
    e[0] = alpha * area * (dldx[0][i] * dldx[0][j] + dldx[1][i] * dldx[1][j]);
  }
};

template <> struct ElemGeo<DFB_ELEM_COT_LAPLACE_TRI3> {  // tri3::emat_cotangent_laplacian (call site laplace_tri3.rs:18)
  static constexpr int NNODE = 3, NDIM = 3, N = 1;
  double cot[3];
  __device__ void init(const double *const *p, double, double) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double *a = p[k], *b = p[(k + 1) % 3], *c = p[(k + 2) % 3];
      const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
      const double v[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      const double cr[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
      const double sn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
      const double cs = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
      cot[k] = cs / sn;
    }
  }
  __device__ void block(int i, int j, double, double, double *e) const {
    const int j1 = (i + 1) % 3, k1 = (i + 2) % 3;
    if (j == i) e[0] = 0.5 * (cot[j1] + cot[k1]);
    else if (j == j1) e[0] = -0.5 * cot[k1];
    else e[0] = -0.5 * cot[j1];
  }
};

template <> struct ElemGeo<DFB_ELEM_SOLID_LINEAR_TRI2> {  // solid_linear_tri2::emat_tri2  solid_linear_tri2.rs:1-33
  static constexpr int NNODE = 3, NDIM = 2, N = 2;
  double w, dldx[2][3];
  __device__ void init(const double *const *p, double, double) {
    w = tri2_area(p[0], p[1], p[2]);
    tri2_dldx(dldx, p[0], p[1], p[2]);
  }
  __device__ void block(int i, int j, double lambda, double myu, double *e) const {
    e[0] = 0.0; e[1] = 0.0; e[2] = 0.0; e[3] = 0.0;
    e[0] += w * (lambda + myu) * dldx[0][i] * dldx[0][j];
    e[2] += w * (lambda * dldx[0][i] * dldx[1][j] + myu * dldx[0][j] * dldx[1][i]);
    e[1] += w * (lambda * dldx[1][i] * dldx[0][j] + myu * dldx[1][j] * dldx[0][i]);
    e[3] += w * (lambda + myu) * dldx[1][i] * dldx[1][j];
    const double dtmp1 = w * myu * (dldx[1][i] * dldx[1][j] + dldx[0][i] * dldx[0][j]);
    e[0] += dtmp1;
    e[3] += dtmp1;
  }
};

template <> struct ElemGeo<DFB_ELEM_LAPLACE_TET> {  // DERIVED from laplace_tri2.rs:12-15
  static constexpr int NNODE = 4, NDIM = 3, N = 1;
  double vol, g[4][3];
  __device__ void init(const double *const *p, double, double) { vol = tet_vol_grad(g, p[0], p[1], p[2], p[3]); }
  __device__ void block(int i, int j, double alpha, double, double *e) const {
    e[0] = alpha * vol * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
  }
};

template <> struct ElemGeo<DFB_ELEM_SOLID_LINEAR_TET> {  // DERIVED from solid_linear_tri2.rs:11-19, solid_linear_hex.rs:18-25
  static constexpr int NNODE = 4, NDIM = 3, N = 3;
  double vol, g[4][3];
  __device__ void init(const double *const *p, double, double) { vol = tet_vol_grad(g, p[0], p[1], p[2], p[3]); }
  __device__ void block(int i, int j, double lambda, double myu, double *e) const {
    double dtmp1 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) e[a + 3 * b] = vol * (lambda * g[i][a] * g[j][b] + myu * g[j][a] * g[i][b]);
      dtmp1 += g[i][a] * g[j][a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a + 3 * a] += vol * myu * dtmp1;
  }
};

template <> struct ElemGeo<DFB_ELEM_MASS_TET> {  // DERIVED (SURVEY A9'): consistent tet mass rho V / 20 (1 + delta_ij) I_3
  static constexpr int NNODE = 4, NDIM = 3, N = 3;
  double vol;
  __device__ void init(const double *const *p, double, double) {
    double g[4][3];
    vol = tet_vol_grad(g, p[0], p[1], p[2], p[3]);
  }
  __device__ void block(int i, int j, double rho, double, double *e) const {
    const double v = rho * vol / 20.0 * (i == j ? 2.0 : 1.0);
#pragma unroll
    for (int q = 0; q < 9; ++q) e[q] = (q == 0 || q == 4 || q == 8) ? v : 0.0;
  }
};

// ---- node records: the per-(element, node) quarter of the element geometry, one aligned 32-byte sector each, so that a
// contribution (e,i,j) needs exactly two sectors (two pairs of 128-bit loads). The block formulas repeat ElemGeo<KIND>::block
// operation for operation (bit-identical results).
struct __align__(32) NodeRec {
  double v[4];
};

template <int KIND> struct NodeOps;  // make(geo, i) -> NodeRec ; block(ri, rj, p0, p1, e)

template <> struct NodeOps<DFB_ELEM_LAPLACE_TRI2> {
  static constexpr bool OK = true;
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_LAPLACE_TRI2> &g, int i) { return {{g.dldx[0][i], g.dldx[1][i], g.area, 0.0}}; }
  __device__ static void block(const NodeRec &a, const NodeRec &b, double alpha, double, double *e) {
    e[0] = alpha * a.v[2] * (a.v[0] * b.v[0] + a.v[1] * b.v[1]);
  }
};
template <> struct NodeOps<DFB_ELEM_SOLID_LINEAR_TRI2> {
  static constexpr bool OK = true;
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_SOLID_LINEAR_TRI2> &g, int i) { return {{g.dldx[0][i], g.dldx[1][i], g.w, 0.0}}; }
  __device__ static void block(const NodeRec &a, const NodeRec &b, double lambda, double myu, double *e) {
    const double w = a.v[2];
    e[0] = 0.0; e[1] = 0.0; e[2] = 0.0; e[3] = 0.0;
    e[0] += w * (lambda + myu) * a.v[0] * b.v[0];
    e[2] += w * (lambda * a.v[0] * b.v[1] + myu * b.v[0] * a.v[1]);
    e[1] += w * (lambda * a.v[1] * b.v[0] + myu * b.v[1] * a.v[0]);
    e[3] += w * (lambda + myu) * a.v[1] * b.v[1];
    const double dtmp1 = w * myu * (a.v[1] * b.v[1] + a.v[0] * b.v[0]);
    e[0] += dtmp1;
    e[3] += dtmp1;
  }
};
template <> struct NodeOps<DFB_ELEM_LAPLACE_TET> {
  static constexpr bool OK = true;
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_LAPLACE_TET> &g, int i) { return {{g.g[i][0], g.g[i][1], g.g[i][2], g.vol}}; }
  __device__ static void block(const NodeRec &a, const NodeRec &b, double alpha, double, double *e) {
    e[0] = alpha * a.v[3] * (a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]);
  }
};
template <> struct NodeOps<DFB_ELEM_SOLID_LINEAR_TET> {
  static constexpr bool OK = true;
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_SOLID_LINEAR_TET> &g, int i) { return {{g.g[i][0], g.g[i][1], g.g[i][2], g.vol}}; }
  __device__ static void block(const NodeRec &ri, const NodeRec &rj, double lambda, double myu, double *e) {
    const double vol = ri.v[3];
    double dtmp1 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int b = 0; b < 3; ++b) e[a + 3 * b] = vol * (lambda * ri.v[a] * rj.v[b] + myu * rj.v[a] * ri.v[b]);
      dtmp1 += ri.v[a] * rj.v[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a + 3 * a] += vol * myu * dtmp1;
  }
};
template <> struct NodeOps<DFB_ELEM_MASS_TET> {
  static constexpr bool OK = true;
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_MASS_TET> &g, int i) { return {{(double)i, 0.0, 0.0, g.vol}}; }
  __device__ static void block(const NodeRec &a, const NodeRec &b, double rho, double, double *e) {
    const double v = rho * a.v[3] / 20.0 * (a.v[0] == b.v[0] ? 2.0 : 1.0);
#pragma unroll
    for (int q = 0; q < 9; ++q) e[q] = (q == 0 || q == 4 || q == 8) ? v : 0.0;
  }
};
template <> struct NodeOps<DFB_ELEM_COT_LAPLACE_TRI3> {
  static constexpr bool OK = false;  // the cotangent block of (i,j) needs the angle at the THIRD corner: stays on the ElemGeo path
  __device__ static NodeRec make(const ElemGeo<DFB_ELEM_COT_LAPLACE_TRI3> &, int) { return {{0, 0, 0, 0}}; }
  __device__ static void block(const NodeRec &, const NodeRec &, double, double, double *) {}
};

