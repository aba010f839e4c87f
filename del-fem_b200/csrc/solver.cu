// solver.cu — conjugate_gradient (del-fem-cpu/src/sparse_square.rs:218-261, del-fem-ls/src/solver_sparse.rs:5-70) and
// preconditioned_conjugate_gradient (sparse_square.rs:264-347, solver_sparse.rs:73-175) with Jacobi / block-Jacobi
// (sparse_ilu.rs:87-219 on an empty off-diagonal pattern).
//
// One iteration = three kernels, no host round trip:
//   K1  q = A p            fused  p.q                       (bsr.cu; deterministic last-block reduction -> red[0])
//   K2  alpha = rho/p.q;   r -= alpha q;  x += alpha p;     fused  r.r, r.(M^-1 r)  -> red[1], red[2]
//   K3  beta = r.z/rho;    p = M^-1 r + beta p;             thread 0 records ||r||^2, tests convergence, swaps rho
// alpha/beta live on the device; the host polls a convergence flag every `iters_per_sync` iterations (one chunk of
// look-ahead); iterations enqueued past convergence exit at their first instruction, so x, r and the history are
// exactly those of the reference's early return.  Multi-GPU: red[] is all-reduced between the kernels.
#include <math.h>

#include <vector>

#include "dfb_internal.h"
#include "reduce.cuh"

dfb_status dfb_spmv_full(const dfb_bsr *m, double *y, double beta, double alpha, double *x, int fuse_dot, int parity,
                         unsigned long long post_seq);
bool dfb_prof_begin(dfb_ctx *c);
void dfb_prof_end(dfb_ctx *c);
dfb_status dfb_prof_flush(dfb_ctx *c, size_t max_pairs);

// ------------------------------------------------------------------------------------------------ preconditioner
template <int N>
__device__ __forceinline__ bool invert_block(double *inv, const double *a) {
  // Gauss-Jordan with partial pivoting on a column-major NxN block
  double m[N][2 * N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      m[i][j] = a[i + N * j];
      m[i][N + j] = (i == j) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int c = 0; c < N; ++c) {
    int p = c;
#pragma unroll
    for (int r = c + 1; r < N; ++r)
      if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    if (m[p][c] == 0.0) return false;
    if (p != c) {
#pragma unroll
      for (int j = 0; j < 2 * N; ++j) {
        const double t = m[p][j];
        m[p][j] = m[c][j];
        m[c][j] = t;
      }
    }
    const double d = 1.0 / m[c][c];
#pragma unroll
    for (int j = 0; j < 2 * N; ++j) m[c][j] *= d;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = m[r][c];
#pragma unroll
      for (int j = 0; j < 2 * N; ++j) m[r][j] -= f * m[c][j];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) inv[i + N * j] = m[i][N + j];
  return true;
}

// diagonal block of row i (separate diagonal, or searched inside the pattern for convention 1)
__device__ __forceinline__ const double *find_diag(uint64_t i, int NN, const double *row2val, const double *idx2val,
                                                   const uint32_t *row2idx, const uint32_t *idx2col) {
  if (row2val) return row2val + i * NN;
  for (uint32_t k = row2idx[i]; k < row2idx[i + 1]; ++k)
    if (idx2col[k] == (uint32_t)i) return idx2val + (uint64_t)k * NN;
  return nullptr;
}

template <int N>
__global__ void k_precond_update(int kind, uint64_t num_blk, const double *__restrict__ row2val, const double *__restrict__ idx2val,
                                 const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col,
                                 double *__restrict__ inv, int *err) {
  constexpr int NN = N * N;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    const double *D = find_diag(i, NN, row2val, idx2val, row2idx, idx2col);
    if (!D) {
      atomicExch(err, 1);
      continue;
    }
    if (kind == DFB_PRECOND_JACOBI) {
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const double v = D[d + N * d];
        if (v == 0.0) atomicExch(err, 1);
        inv[i * N + d] = 1.0 / v;
      }
    } else {
      double a[NN], b[NN];
#pragma unroll
      for (int q = 0; q < NN; ++q) a[q] = D[q];
      if (!invert_block<N>(b, a)) {
        atomicExch(err, 1);
        continue;
      }
#pragma unroll
      for (int q = 0; q < NN; ++q) inv[i * NN + q] = b[q];
    }
  }
}

static dfb_status precond_create(dfb_ctx *ctx, int32_t kind, const dfb_bsr *m, uint64_t lev_fill, dfb_precond **out) {
  DFB_REQUIRE(ctx && m && out, "dfb_precond_create: NULL argument");
  DFB_REQUIRE(kind >= DFB_PRECOND_NONE && kind <= DFB_PRECOND_ILU0, "dfb_precond_create: unknown kind %d", kind);
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_precond *pc = new dfb_precond();
  pc->ctx = ctx;
  pc->kind = kind;
  pc->blk_dim = m->blk_dim;
  pc->num_blk = m->pat->num_blk;
  pc->d_inv = nullptr;
  if (kind != DFB_PRECOND_NONE) {
    const uint64_t per = kind == DFB_PRECOND_JACOBI ? (uint64_t)m->blk_dim : (uint64_t)m->blk_dim * m->blk_dim;
    dfb_status st = dfb_dev_alloc_t(&pc->d_inv, pc->num_blk * per + 8);
    if (st == DFB_OK && kind == DFB_PRECOND_ILU0) {
      st = dfb_ilu_create(ctx, m, lev_fill, &pc->ilu);
      if (st == DFB_OK) {
        pc->pat = m->pat;
        m->pat->refcount += 1;
      }
    }
    if (st != DFB_OK) {
      dfb_dev_free(pc->d_inv);
      delete pc;
      return st;
    }
  }
  *out = pc;
  return DFB_OK;
}

DFB_API dfb_status dfb_precond_create(dfb_ctx *ctx, int32_t kind, const dfb_bsr *m, dfb_precond **out) {
  return precond_create(ctx, kind, m, 0, out);
}

// initialize_iluk (sparse_ilu.rs:58-64): the ILU preconditioner on the level-of-fill pattern; lev_fill = 0 is ILU(0)
DFB_API dfb_status dfb_precond_create_iluk(dfb_ctx *ctx, const dfb_bsr *m, uint64_t lev_fill, dfb_precond **out) {
  return precond_create(ctx, DFB_PRECOND_ILU0, m, lev_fill, out);
}

DFB_API dfb_status dfb_precond_update(dfb_precond *pc, const dfb_bsr *m) {
  DFB_REQUIRE(pc && m, "dfb_precond_update: NULL argument");
  DFB_REQUIRE(pc->num_blk == m->pat->num_blk && pc->blk_dim == m->blk_dim, "dfb_precond_update: shape mismatch");
  if (pc->kind == DFB_PRECOND_NONE) return DFB_OK;
  dfb_ctx *c = pc->ctx;
  if (pc->kind == DFB_PRECOND_ILU0) {
    DFB_REQUIRE(pc->pat == m->pat, "dfb_precond_update: the ILU(0) was initialised for another pattern");
    return dfb_ilu_update(c, pc->ilu, m, pc->d_inv);
  }
  const int grid = c->sm_count * 8;
#define PC_CASE(NV)                                                                                                       \
  case NV:                                                                                                                \
    DFB_LAUNCH(c, k_precond_update<NV>, grid, 128, 0, pc->kind, pc->num_blk, m->d_row2val, m->d_idx2val, m->pat->d_row2idx, \
               m->pat->d_idx2col, pc->d_inv, c->d_errflag + 2);                                                           \
    break;
  switch (pc->blk_dim) {
    PC_CASE(1)
    PC_CASE(2)
    PC_CASE(3)
    PC_CASE(4)
  }
#undef PC_CASE
  DFB_CHECK_LAUNCH();
  return dfb_check_error_flag(c, 2, DFB_ERR_SINGULAR_DIAG, "dfb_precond_update: singular (or missing) diagonal block (try_inverse().unwrap())");
}

template <int N, int KIND>
__device__ __forceinline__ void apply_minv(double (&z)[N], const double *__restrict__ inv, uint64_t i, const double (&r)[N]) {
  if (KIND == DFB_PRECOND_NONE) {
#pragma unroll
    for (int d = 0; d < N; ++d) z[d] = r[d];
  } else if (KIND == DFB_PRECOND_JACOBI) {
#pragma unroll
    for (int d = 0; d < N; ++d) z[d] = inv[i * N + d] * r[d];
  } else if (KIND == DFB_PRECOND_ILU0) {
    // explicit z: `inv` points at the vector M^-1 r computed beforehand by the triangular sweeps
#pragma unroll
    for (int d = 0; d < N; ++d) z[d] = inv[i * N + d];
  } else {
    const double *B = inv + i * N * N;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      double s = 0.0;
#pragma unroll
      for (int b = 0; b < N; ++b) s += B[a + N * b] * r[b];
      z[a] = s;
    }
  }
}

template <int N, int KIND>
__global__ void k_precond_apply(double *__restrict__ v, const double *__restrict__ inv, uint64_t num_blk) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    double r[N], z[N];
#pragma unroll
    for (int d = 0; d < N; ++d) r[d] = v[i * N + d];
    apply_minv<N, KIND>(z, inv, i, r);
#pragma unroll
    for (int d = 0; d < N; ++d) v[i * N + d] = z[d];
  }
}

#define DISPATCH_N_KIND(NV, KV, MACRO) \
  if (n == NV && kind == KV) { MACRO(NV, KV); }
#define DISPATCH_ALL(MACRO)                                 \
  DISPATCH_N_KIND(1, 0, MACRO) DISPATCH_N_KIND(1, 1, MACRO) DISPATCH_N_KIND(1, 2, MACRO) \
  DISPATCH_N_KIND(2, 0, MACRO) DISPATCH_N_KIND(2, 1, MACRO) DISPATCH_N_KIND(2, 2, MACRO) \
  DISPATCH_N_KIND(3, 0, MACRO) DISPATCH_N_KIND(3, 1, MACRO) DISPATCH_N_KIND(3, 2, MACRO) \
  DISPATCH_N_KIND(4, 0, MACRO) DISPATCH_N_KIND(4, 1, MACRO) DISPATCH_N_KIND(4, 2, MACRO)

DFB_API dfb_status dfb_precond_apply(const dfb_precond *pc, dfb_vec *v) {
  DFB_REQUIRE(pc && v, "dfb_precond_apply: NULL argument");
  DFB_REQUIRE(v->n_blk == pc->num_blk && v->blk_dim == pc->blk_dim, "dfb_precond_apply: shape mismatch");
  dfb_ctx *c = pc->ctx;
  const int n = pc->blk_dim, kind = pc->kind;
  if (kind == DFB_PRECOND_ILU0) return dfb_ilu_apply(c, pc->ilu, pc->pat, n, pc->d_inv, v->d);
  const int grid = c->sm_count * 8;
#define APPLY(NV, KV) DFB_LAUNCH(c, (k_precond_apply<NV, KV>), grid, 256, 0, v->d, pc->d_inv, pc->num_blk)
  DISPATCH_ALL(APPLY)
#undef APPLY
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

DFB_API dfb_status dfb_precond_destroy(dfb_precond *pc) {
  if (!pc) return DFB_OK;
  cudaStreamSynchronize(pc->ctx->stream);
  dfb_dev_free(pc->d_inv);
  dfb_ilu_destroy(pc->ilu);
  if (pc->pat) dfb_pattern_destroy(pc->pat);
  delete pc;
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ (P)CG kernels
// K0: x = 0 ; p = M^-1 r ; red[1] = r.r ; red[2] = r.(M^-1 r)
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_init(const double *__restrict__ r, double *__restrict__ x, double *__restrict__ p,
                                                  const double *__restrict__ inv, uint64_t num_blk, DfbSolverState *st,
                                                  double *partials, unsigned int *ticket, const DfbPeerView pv_out) {
  __shared__ double s_buf[64];
  double acc[2] = {0.0, 0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    double rv[N], z[N];
#pragma unroll
    for (int d = 0; d < N; ++d) rv[d] = r[i * N + d];
    apply_minv<N, KIND>(z, inv, i, rv);
#pragma unroll
    for (int d = 0; d < N; ++d) {
      x[i * N + d] = 0.0;
      p[i * N + d] = z[d];
      acc[0] += rv[d] * rv[d];
      acc[1] += rv[d] * z[d];
    }
  }
  if (grid_sum_last_block<2>(acc, partials, ticket, s_buf)) {
    if (threadIdx.x == 0) {
      st->red[1] = acc[0];
      st->red[2] = acc[1];
      s_buf[0] = acc[0];
      s_buf[1] = acc[1];
    }
    if (pv_out.nranks > 1 && pv_out.seq != 0) peer_post_block<2>(pv_out, s_buf);
  }
}

// after the (all-reduced) initial dots: set up the state. hist[0] = r0.r0
__global__ void k_pcg_init_finalize(DfbSolverState *st, double *hist, double tol, double eps_sq, int check_pap,
                                    const DfbPeerView pv_in) {
  double g[2] = {st->red[1], st->red[2]};
  int perr = 0;
  if (pv_in.nranks > 1 && pv_in.seq != 0) peer_gather<2>(pv_in, g, &perr);
  if (threadIdx.x != 0) return;
  const double rr0 = g[0];
  st->rr0 = rr0;
  st->inv_rr0 = 1.0 / rr0;
  st->rho[0] = g[1];
  st->rho[1] = 0.0;
  st->tol = tol;
  st->eps_sq = eps_sq;
  st->done[0] = (rr0 < eps_sq) ? 1 : 0;
  st->done[1] = st->done[0];
  st->iter = 0;
  st->err = perr;
  st->check_pap = check_pap;
  hist[0] = rr0;
}

// K2: alpha = rho / (p.q) ; r -= alpha q ; x += alpha p ; red[1] = r.r ; red[2] = r.(M^-1 r)
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_update(double *__restrict__ r, double *__restrict__ x, const double *__restrict__ q,
                                                    const double *__restrict__ p, const double *__restrict__ inv,
                                                    uint64_t num_blk, DfbSolverState *st, int parity, double *partials,
                                                    unsigned int *ticket, const DfbPeerView pv_in, const DfbPeerView pv_out) {
  __shared__ double s_buf[64];
  if (st->done[parity]) return;
  double pq = st->red[0];
  if (pv_in.nranks > 1) {
    double g[1];
    peer_gather<1>(pv_in, g, &st->err);
    pq = g[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->red[0] = pq;
  }
  const double alpha = st->rho[parity] / pq;
  double acc[2] = {0.0, 0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    double rv[N], z[N];
#pragma unroll
    for (int d = 0; d < N; ++d) {
      rv[d] = r[i * N + d] - alpha * q[i * N + d];
      r[i * N + d] = rv[d];
      x[i * N + d] += alpha * p[i * N + d];
    }
    apply_minv<N, KIND>(z, inv, i, rv);
#pragma unroll
    for (int d = 0; d < N; ++d) {
      acc[0] += rv[d] * rv[d];
      acc[1] += rv[d] * z[d];
    }
  }
  if (grid_sum_last_block<2>(acc, partials, ticket, s_buf)) {
    if (threadIdx.x == 0) {
      st->red[1] = acc[0];
      st->red[2] = acc[1];
      s_buf[0] = acc[0];
      s_buf[1] = acc[1];
    }
    if (pv_out.nranks > 1) peer_post_block<2>(pv_out, s_buf);
  }
}

// K3: beta = r.z / rho ; p = M^-1 r + beta p ; thread 0 of block 0: history, convergence, rho swap
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_direction(const double *__restrict__ r, double *__restrict__ p,
                                                       const double *__restrict__ inv, uint64_t num_blk, DfbSolverState *st,
                                                       int parity, double *hist, uint64_t hist_cap, const DfbPeerView pv_in) {
  if (st->done[parity]) {
    // converged earlier: propagate the flag into the slot the next iteration reads, change nothing else
    if (blockIdx.x == 0 && threadIdx.x == 0) st->done[parity ^ 1] = 1;
    return;
  }
  double rr = st->red[1], rz = st->red[2];
  if (pv_in.nranks > 1) {
    double g[2];
    peer_gather<2>(pv_in, g, &st->err);
    rr = g[0];
    rz = g[1];
  }
  const double beta = rz / st->rho[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it = st->iter + 1;
    st->iter = it;
    if ((uint64_t)it < hist_cap) hist[it] = rr;
    const double conv_ratio = sqrt(rr * st->inv_rr0);
    int done = (conv_ratio < st->tol) ? 1 : 0;
    if (st->check_pap && !(st->red[0] >= 0.0)) {
      st->err = 1;
      done = 1;
    }
    st->done[parity ^ 1] = done;
    st->rho[parity ^ 1] = rz;
  }
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    double rv[N], z[N];
#pragma unroll
    for (int d = 0; d < N; ++d) rv[d] = r[i * N + d];
    apply_minv<N, KIND>(z, inv, i, rv);
#pragma unroll
    for (int d = 0; d < N; ++d) p[i * N + d] = z[d] + beta * p[i * N + d];
  }
}

// ---- flat (element-wise) forms of K2 / K3 for the preconditioners that act per scalar (none, point Jacobi):
// 128-bit loads/stores over the n_dof scalars instead of 24-byte block rows.
template <int KIND>
__global__ void __launch_bounds__(256) k_pcg_update_flat(double *__restrict__ r, double *__restrict__ x, const double *__restrict__ q,
                                                         const double *__restrict__ p, const double *__restrict__ inv,
                                                         uint64_t n, DfbSolverState *st, int parity, double *partials,
                                                         unsigned int *ticket, const DfbPeerView pv_in, const DfbPeerView pv_out) {
  __shared__ double s_buf[64];
  if (st->done[parity]) return;
  double pq = st->red[0];
  if (pv_in.nranks > 1) {
    double g[1];
    peer_gather<1>(pv_in, g, &st->err);
    pq = g[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) st->red[0] = pq;
  }
  const double alpha = st->rho[parity] / pq;
  double acc[2] = {0.0, 0.0};
  const uint64_t n2 = n >> 1;
  double2 *r2 = reinterpret_cast<double2 *>(r);
  double2 *x2 = reinterpret_cast<double2 *>(x);
  const double2 *q2 = reinterpret_cast<const double2 *>(q);
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  const double2 *i2 = reinterpret_cast<const double2 *>(inv);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 rv = r2[i], xv = x2[i];
    const double2 qv = q2[i], pv = p2[i];
    rv.x -= alpha * qv.x;
    rv.y -= alpha * qv.y;
    xv.x += alpha * pv.x;
    xv.y += alpha * pv.y;
    r2[i] = rv;
    x2[i] = xv;
    double zx = rv.x, zy = rv.y;
    if (KIND == DFB_PRECOND_JACOBI) {
      const double2 iv = i2[i];
      zx *= iv.x;
      zy *= iv.y;
    }
    acc[0] += rv.x * rv.x + rv.y * rv.y;
    acc[1] += rv.x * zx + rv.y * zy;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint64_t i = n - 1;
    const double rv = r[i] - alpha * q[i];
    r[i] = rv;
    x[i] += alpha * p[i];
    const double z = (KIND == DFB_PRECOND_JACOBI) ? inv[i] * rv : rv;
    acc[0] += rv * rv;
    acc[1] += rv * z;
  }
  if (grid_sum_last_block<2>(acc, partials, ticket, s_buf)) {
    if (threadIdx.x == 0) {
      st->red[1] = acc[0];
      st->red[2] = acc[1];
      s_buf[0] = acc[0];
      s_buf[1] = acc[1];
    }
    if (pv_out.nranks > 1) peer_post_block<2>(pv_out, s_buf);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_pcg_direction_flat(const double *__restrict__ r, double *__restrict__ p,
                                                            const double *__restrict__ inv, uint64_t n, DfbSolverState *st,
                                                            int parity, double *hist, uint64_t hist_cap, const DfbPeerView pv_in) {
  if (st->done[parity]) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->done[parity ^ 1] = 1;
    return;
  }
  double rr = st->red[1], rz = st->red[2];
  if (pv_in.nranks > 1) {
    double g[2];
    peer_gather<2>(pv_in, g, &st->err);
    rr = g[0];
    rz = g[1];
  }
  const double beta = rz / st->rho[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int it = st->iter + 1;
    st->iter = it;
    if ((uint64_t)it < hist_cap) hist[it] = rr;
    const double conv_ratio = sqrt(rr * st->inv_rr0);
    int done = (conv_ratio < st->tol) ? 1 : 0;
    if (st->check_pap && !(st->red[0] >= 0.0)) {
      st->err = 1;
      done = 1;
    }
    st->done[parity ^ 1] = done;
    st->rho[parity ^ 1] = rz;
  }
  const uint64_t n2 = n >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  double2 *p2 = reinterpret_cast<double2 *>(p);
  const double2 *i2 = reinterpret_cast<const double2 *>(inv);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 z = r2[i], pv = p2[i];
    if (KIND == DFB_PRECOND_JACOBI) {
      const double2 iv = i2[i];
      z.x *= iv.x;
      z.y *= iv.y;
    }
    pv.x = z.x + beta * pv.x;
    pv.y = z.y + beta * pv.y;
    p2[i] = pv;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint64_t i = n - 1;
    const double z = (KIND == DFB_PRECOND_JACOBI) ? inv[i] * r[i] : r[i];
    p[i] = z + beta * p[i];
  }
}

// One warp collects the R partial sums of exchange pv.seq from the LOCAL mailbox and stores the rank-ordered total where the
// single-GPU kernels expect it. (Letting every block of the consumer kernel spin on the mailbox line was measured slower than
// NCCL: ~10^4 spinning threads starve the incoming NVLink write.)
template <int K>
__global__ void k_peer_gather(const DfbPeerView pv, DfbSolverState *st, int first, int parity) {
  if (parity >= 0 && st->done[parity]) return;  // converged: the producer did not post either
  double g[K];
  int err = 0;
  peer_gather<K>(pv, g, &err);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) st->red[first + k] = g[k];
    if (err) st->err = 2;
  }
}

struct FlagsHost {
  int done[2];
  int iter;
  int err;
};

static int vec_grid(const dfb_ctx *c, uint64_t num_blk) {
  uint64_t g = (num_blk + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 8;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  if (g > DFB_RED_MAX_BLOCKS) g = DFB_RED_MAX_BLOCKS;
  return (int)g;
}

// shared driver. mode 0 = CG history semantics (ratios, k=0 excluded), 1 = PCG (absolute, k=0 included, trailing entry)
static dfb_status run_pcg(const dfb_bsr *m, const dfb_precond *pc, dfb_vec *r, dfb_vec *x, dfb_vec *q, dfb_vec *p, double tol,
                          uint64_t max_it, double eps_sq, int check_pap, int mode, double *hist_out, uint64_t hist_cap,
                          uint64_t *n_hist) {
  dfb_ctx *c = m->ctx;
  const uint64_t nb = m->pat->num_blk;
  const int n = m->blk_dim;
  const int kind = pc ? pc->kind : DFB_PRECOND_NONE;
  const double *inv = pc ? pc->d_inv : nullptr;
  const bool ilu = kind == DFB_PRECOND_ILU0;
  double *zbuf = ilu ? dfb_ilu_z(pc->ilu) : nullptr;
  if (ilu) DFB_REQUIRE(pc->pat == m->pat, "pcg: the ILU(0) preconditioner belongs to another pattern");
  if (ilu) DFB_REQUIRE(c->nranks <= 1, "pcg: ILU(0) is single-GPU only");
  for (dfb_vec *v : {r, x, q, p}) {
    DFB_REQUIRE(v && v->n_blk == nb && v->blk_dim == n, "(p)cg: vector shape != matrix shape (assert_eq!(r_vec.len(), mat.num_blk))");
    DFB_REQUIRE(v->ctx == c, "(p)cg: vector belongs to another context");
  }
  if (pc) DFB_REQUIRE(pc->num_blk == nb && pc->blk_dim == n, "pcg: preconditioner shape != matrix shape");
  DFB_REQUIRE(max_it < 0x7FFFFFF0ull, "(p)cg: max_it too large");
  // device history buffer
  if (c->hist_cap < max_it + 2) {
    dfb_dev_free(c->d_hist);
    c->d_hist = nullptr;
    DFB_TRY(dfb_dev_alloc_t(&c->d_hist, max_it + 2));
    c->hist_cap = max_it + 2;
  }
  const int G = vec_grid(c, nb);
  const uint64_t nflat = nb * (uint64_t)n;
  const int GF = vec_grid(c, (nflat + 1) / 2);
  DfbSolverState *st = c->d_state;

  // scalar all-reduces: fused into the kernels through the NVLink peer mailboxes when mapped, NCCL otherwise
  const bool peer = c->peer_enabled && c->use_peer_allreduce && c->nranks > 1;
  const bool nccl_red = c->nranks > 1 && !peer;
  const DfbPeerView pv_none = dfb_peer_view(c, 0);
  DfbPeerView pv_off = pv_none;  // consumers read st->red[] (filled by k_peer_gather / NCCL / the local reduction)
  pv_off.nranks = 1;
  DfbPeerView pv_i = peer ? dfb_peer_view(c, ++c->peer_seq) : pv_none;
  if (!peer) pv_i.nranks = 1;
  if (ilu) {
    // z = M^-1 r by the level-scheduled sweeps, then the generic init with z taken as given
    DFB_CUDA(cudaMemcpyAsync(zbuf, r->d, nb * (uint64_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    DFB_TRY(dfb_ilu_apply(c, pc->ilu, pc->pat, n, pc->d_inv, zbuf));
#define INITZ(NV) \
  if (n == NV) DFB_LAUNCH(c, (k_pcg_init<NV, DFB_PRECOND_ILU0>), G, 256, 0, r->d, x->d, p->d, zbuf, nb, st, c->d_partials, c->d_ticket + 2, pv_i);
    INITZ(1) INITZ(2) INITZ(3) INITZ(4)
#undef INITZ
  }
#define INIT(NV, KV) DFB_LAUNCH(c, (k_pcg_init<NV, KV>), G, 256, 0, r->d, x->d, p->d, inv, nb, st, c->d_partials, c->d_ticket + 2, pv_i)
  DISPATCH_ALL(INIT)
#undef INIT
  DFB_CHECK_LAUNCH();
  if (nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[1], 2, c->stream));
  if (peer) DFB_LAUNCH(c, k_peer_gather<2>, 1, 32, 0, pv_i, st, 1, -1);
  DFB_LAUNCH(c, k_pcg_init_finalize, 1, 32, 0, st, c->d_hist, tol, eps_sq, check_pap, pv_off);
  DFB_CHECK_LAUNCH();

  FlagsHost *hf = (FlagsHost *)c->h_pinned;  // two slots
  hf[0].done[0] = hf[0].done[1] = hf[1].done[0] = hf[1].done[1] = 0;
  uint64_t chunk = c->iters_per_sync > 0 ? (uint64_t)c->iters_per_sync : 16;
  uint64_t enq = 0;  // iterations enqueued
  int n_chunks = 0;
  bool finished = false;
  FlagsHost last = {{0, 0}, 0, 0};
  while (!finished) {
    const uint64_t todo = (max_it - enq < chunk) ? (max_it - enq) : chunk;
    for (uint64_t j = 0; j < todo; ++j) {
      const int parity = (int)((enq + j) & 1);
      // profile_spmv = K: bracket every K-th SpMV with events (K = 1: all of them)
      const bool prof = (c->profile_spmv > 0 && (enq + j) % (uint64_t)c->profile_spmv == 0) ? dfb_prof_begin(c) : false;
      DfbPeerView pv_a = pv_none, pv_b = pv_none;  // a: p.Ap (K1 -> K2), b: r.r, r.z (K2 -> K3)
      pv_a.nranks = pv_b.nranks = 1;
      if (peer) {
        pv_a = dfb_peer_view(c, ++c->peer_seq);
        pv_b = dfb_peer_view(c, ++c->peer_seq);
      }
      DFB_TRY(dfb_spmv_full(m, q->d, 0.0, 1.0, p->d, 1, parity, peer ? pv_a.seq : 0));
      if (prof) dfb_prof_end(c);
      if (nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[0], 1, c->stream));
      if (peer) DFB_LAUNCH(c, k_peer_gather<1>, 1, 32, 0, pv_a, st, 0, parity);
      if (kind == DFB_PRECOND_BLOCK_JACOBI) {
#define UPD(NV, KV) DFB_LAUNCH(c, (k_pcg_update<NV, KV>), G, 256, 0, r->d, x->d, q->d, p->d, inv, nb, st, parity, c->d_partials, c->d_ticket + 2, pv_off, pv_b)
        DISPATCH_ALL(UPD)
#undef UPD
      } else if (kind == DFB_PRECOND_JACOBI) {
        DFB_LAUNCH(c, k_pcg_update_flat<DFB_PRECOND_JACOBI>, GF, 256, 0, r->d, x->d, q->d, p->d, inv, nflat, st, parity, c->d_partials, c->d_ticket + 2, pv_off, pv_b);
      } else {
        DFB_LAUNCH(c, k_pcg_update_flat<DFB_PRECOND_NONE>, GF, 256, 0, r->d, x->d, q->d, p->d, inv, nflat, st, parity, c->d_partials, c->d_ticket + 2, pv_off, pv_b);
      }
      if (ilu) {
        // the update kernel ran with M = I (r.r in red[1]); now z = M^-1 r explicitly and r.z -> red[2]
        DFB_CUDA(cudaMemcpyAsync(zbuf, r->d, nflat * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        DFB_TRY(dfb_ilu_apply(c, pc->ilu, pc->pat, n, pc->d_inv, zbuf));
        DFB_TRY(dfb_ilu_dot(c, r->d, zbuf, nflat, &st->red[2]));
      }
      if (nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[1], 2, c->stream));
      if (peer) DFB_LAUNCH(c, k_peer_gather<2>, 1, 32, 0, pv_b, st, 1, parity);
      if (kind == DFB_PRECOND_BLOCK_JACOBI) {
#define DIR(NV, KV) DFB_LAUNCH(c, (k_pcg_direction<NV, KV>), G, 256, 0, r->d, p->d, inv, nb, st, parity, c->d_hist, c->hist_cap, pv_off)
        DISPATCH_ALL(DIR)
#undef DIR
      } else if (kind == DFB_PRECOND_JACOBI) {
        DFB_LAUNCH(c, k_pcg_direction_flat<DFB_PRECOND_JACOBI>, GF, 256, 0, r->d, p->d, inv, nflat, st, parity, c->d_hist, c->hist_cap, pv_off);
      } else if (ilu) {
#define DIRZ(NV) \
  if (n == NV) DFB_LAUNCH(c, (k_pcg_direction<NV, DFB_PRECOND_ILU0>), G, 256, 0, r->d, p->d, zbuf, nb, st, parity, c->d_hist, c->hist_cap, pv_off);
        DIRZ(1) DIRZ(2) DIRZ(3) DIRZ(4)
#undef DIRZ
      } else {
        DFB_LAUNCH(c, k_pcg_direction_flat<DFB_PRECOND_NONE>, GF, 256, 0, r->d, p->d, inv, nflat, st, parity, c->d_hist, c->hist_cap, pv_off);
      }
    }
    DFB_CHECK_LAUNCH();
    enq += todo;
    const int slot = n_chunks & 1;
    DFB_CUDA(cudaMemcpyAsync(&hf[slot], &st->done[0], sizeof(FlagsHost), cudaMemcpyDeviceToHost, c->stream));
    DFB_CUDA(cudaEventRecord(c->ev_flag[slot], c->stream));
    ++n_chunks;
    // look at the flags of the PREVIOUS chunk (one chunk of look-ahead keeps the GPU busy)
    if (n_chunks >= 2) {
      const int prev = (n_chunks - 2) & 1;
      DFB_CUDA(cudaEventSynchronize(c->ev_flag[prev]));
      last = hf[prev];
      if (last.done[0] || last.done[1]) finished = true;
    }
    if (enq >= max_it || todo == 0) {
      DFB_CUDA(cudaEventSynchronize(c->ev_flag[slot]));
      last = hf[slot];
      finished = true;
    }
  }
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  {
    const int slot = (n_chunks - 1) & 1;
    last = hf[slot];
  }
  DFB_TRY(dfb_prof_flush(c, c->profile_spmv > 0 ? ((size_t)last.iter + (size_t)c->profile_spmv - 1) / (size_t)c->profile_spmv : 0));
  if (last.err == 2) return dfb_fail(DFB_ERR_NCCL, "(p)cg: peer mailbox all-reduce timed out (a rank is missing or out of step)");
  if (last.err) return dfb_fail(DFB_ERR_NOT_POSITIVE, "cg: p.Ap < 0 (assert!(pap >= T::zero()))");

  // history: the device stored ||r_k||^2 for k = 0..iter
  const uint64_t iters = (uint64_t)last.iter;
  std::vector<double> rr(iters + 1);
  DFB_CUDA(cudaMemcpyAsync(rr.data(), c->d_hist, (iters + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  const double rr0 = rr[0];
  const bool converged = last.done[0] || last.done[1];
  uint64_t nh = 0;
  auto push = [&](double v) {
    if (hist_out && nh < hist_cap) hist_out[nh] = v;
    ++nh;
  };
  if (mode == 0) {
    // CG: empty history if ||r0||^2 < eps (sparse_square.rs:235); ratios for k>=1
    if (!(rr0 < eps_sq)) {
      const double inv0 = 1.0 / rr0;
      for (uint64_t k = 1; k <= iters; ++k) push(sqrt(rr[k] * inv0));
    }
  } else {
    // PCG: absolute norms incl. k=0 (:291,:321), one trailing entry when max_it is exhausted (:341-345)
    push(sqrt(rr0));
    if (!(rr0 < eps_sq)) {
      for (uint64_t k = 1; k <= iters; ++k) push(sqrt(rr[k]));
      if (!converged) push(sqrt(rr[iters]));
    }
  }
  if (n_hist) *n_hist = nh;
  return DFB_OK;
}

DFB_API dfb_status dfb_cg(const dfb_bsr *m, dfb_vec *r, dfb_vec *u, dfb_vec *ap, dfb_vec *p, double tol, uint64_t max_it,
                          double eps_sq, int32_t check_pap, double *hist_out, uint64_t hist_cap, uint64_t *n_hist) {
  DFB_REQUIRE(m, "dfb_cg: matrix is NULL");
  return run_pcg(m, nullptr, r, u, ap, p, tol, max_it, eps_sq, check_pap, 0, hist_out, hist_cap, n_hist);
}

DFB_API dfb_status dfb_pcg(const dfb_bsr *m, const dfb_precond *pc, dfb_vec *r, dfb_vec *x, dfb_vec *pr, dfb_vec *p, double tol,
                           uint64_t max_it, double eps_sq, double *hist_out, uint64_t hist_cap, uint64_t *n_hist) {
  DFB_REQUIRE(m && pc, "dfb_pcg: NULL argument");
  return run_pcg(m, pc, r, x, pr, p, tol, max_it, eps_sq, 0, 1, hist_out, hist_cap, n_hist);
}
