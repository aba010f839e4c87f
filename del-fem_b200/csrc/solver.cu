// solver.cu — conjugate_gradient (del-fem-cpu/src/sparse_square.rs:218-261, del-fem-ls/src/solver_sparse.rs:5-70) and
// preconditioned_conjugate_gradient (sparse_square.rs:264-347, solver_sparse.rs:73-175) with Jacobi / block-Jacobi
// (sparse_ilu.rs:87-219 on an empty off-diagonal pattern).
//
// One iteration = three kernels, no host round trip and NO per-iteration host arguments:
//   K1  q = A p            fused  p.q                       (bsr.cu; deterministic last-block reduction -> red[0])
//   K2  alpha = rho/p.q;   r -= alpha q;  x += alpha p;     fused  r.r, r.(M^-1 r)  -> red[1], red[2]
//   K3  beta = r.z/rho;    p = M^-1 r + beta p;             thread 0 records ||r||^2, tests convergence, swaps rho
// alpha/beta, the iteration index, its parity and the exchange ids of the multi-GPU steps live in the device-side solver state
// (dfb_internal.h), so every iteration is the SAME three launches: a chunk of `iters_per_sync` iterations is captured once
// into a CUDA graph and replayed; the host only polls a stop flag per chunk (one chunk of look-ahead).  Iterations enqueued past
// convergence / max_it exit at their first instruction, so x, r and the history are exactly those of the reference's early
// return.  Multi-GPU: the two scalar all-reduces are fused into these kernels over NVLink peer memory (peer.cuh: the last
// block of the producer posts, the first block of the consumer gathers and releases a local flag), and the direction vector's
// boundary entries are pushed into the neighbours' landing zones right after K3 (the next K1's boundary tiles wait for the
// neighbours' flags) — no collective launch in the loop; NCCL (all-reduce + send/recv on a second stream) is the fallback.
#include <math.h>
#include <string.h>

#include <vector>

#include "dfb_internal.h"
#include "reduce.cuh"

dfb_status dfb_spmv_full(const dfb_bsr *m, double *y, double beta, double alpha, double *x, int fuse_dot, int in_solver, int post);
bool dfb_spmv_uses_peer_halo(const dfb_bsr *m);
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
                                 const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col, const PackedView pv,
                                 double *__restrict__ inv, int *err) {
  constexpr int NN = N * N;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    // diagonal block: from the packed records when they are the matrix's current layout (pv.base != nullptr), else canonical
    const double *D = pv.base ? pk_diag(pv, i) : find_diag(i, NN, row2val, idx2val, row2idx, idx2col);
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
  PackedView pv = dfb_bsr_packed_view(m);
  if (m->packed_valid && !m->canon_valid && pv.has_dia) {
    // the packed records are the current layout: read the diagonal blocks from them
  } else {
    DFB_TRY(dfb_bsr_canon_read(m));
    pv.base = nullptr;
  }
  dfb_phase_begin(c, DFB_PH_PRECOND);
#define PC_CASE(NV)                                                                                                       \
  case NV:                                                                                                                \
    DFB_LAUNCH(c, k_precond_update<NV>, grid, 128, 0, pc->kind, pc->num_blk, m->pat->convention == 0 ? m->d_row2val : nullptr, \
               m->d_idx2val, m->pat->d_row2idx, m->pat->d_idx2col, pv, pc->d_inv, c->d_errflag + 2);                       \
    break;
  switch (pc->blk_dim) {
    PC_CASE(1)
    PC_CASE(2)
    PC_CASE(3)
    PC_CASE(4)
  }
#undef PC_CASE
  dfb_phase_end(c);
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
// K0: x = 0 ; p = M^-1 r ; red[1] = r.r ; red[2] = r.(M^-1 r).  Posts them to the peers as exchange xseq_next + 1.
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_init(const double *__restrict__ r, double *__restrict__ x, double *__restrict__ p,
                                                  const double *__restrict__ inv, uint64_t num_blk, DfbSolverState *st,
                                                  double *partials, unsigned int *ticket, const DfbPeerView pv) {
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
    if (pv.nranks > 1 && pv.post) peer_post_block<2>(pv, st->xseq_next + 1ull, s_buf);
  }
}

// after the (all-reduced) initial dots: set up the state of this solve (one block). The ring entry 0 holds r0.r0.
__global__ void k_pcg_init_finalize(DfbSolverState *st, double *hist, double tol, double eps_sq, int check_pap, int max_it,
                                    const DfbPeerView pv) {
  double g[2] = {st->red[1], st->red[2]};
  int perr = 0;
  const unsigned long long xb = st->xseq_next, hb = st->hseq_next;
  if (pv.nranks > 1 && pv.post) peer_gather<2>(pv, xb + 1ull, g, &perr);
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double rr0 = g[0];
  st->rr0 = rr0;
  st->inv_rr0 = 1.0 / rr0;
  st->rho[0] = g[1];
  st->rho[1] = 0.0;
  st->tol = tol;
  st->eps_sq = eps_sq;
  st->xbase = xb;
  st->hbase = hb;
  st->done[0] = (rr0 < eps_sq || max_it <= 0) ? 1 : 0;
  st->done[1] = st->done[0];
  st->iter = 0;
  st->iter_dir = 0;
  st->err = perr ? 2 : 0;
  st->check_pap = check_pap;
  st->converged = (rr0 < eps_sq) ? 1 : 0;
  st->max_it = max_it;
  hist[0] = rr0;
}

// after the last chunk: retire the exchange ids this solve used (k completed iterations: 1 + 2k scalar exchanges, 1 + k pushes)
__global__ void k_pcg_end(DfbSolverState *st) {
  st->xseq_next = st->xbase + 2ull + 2ull * (unsigned long long)st->iter;
  st->hseq_next = st->hbase + 2ull + (unsigned long long)st->iter;
}

// what K2 needs before touching the vectors: iteration index, stop flag, (all-reduced) p.q
struct IterHead {
  int it, parity;
  bool stop;
  double pq;
};
__device__ __forceinline__ IterHead update_head(DfbSolverState *st, const DfbPeerView &pv, unsigned int *gticket) {
  IterHead h;
  h.it = st->iter;
  h.parity = h.it & 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) st->iter_dir = h.it;  // the direction kernel reads its iteration index from here
  h.stop = st->done[h.parity] != 0;
  h.pq = 0.0;
  if (h.stop) return h;
  if (pv.nranks > 1) {
    double g[1];
    peer_gather_grid<1>(pv, dfb_xid_pq(st->xbase, h.it), &st->red[0], &st->gflag[0], gticket, &st->err, g);
    h.pq = g[0];
  } else {
    h.pq = st->red[0];
  }
  return h;
}

// K2: alpha = rho / (p.q) ; r -= alpha q ; x += alpha p ; red[1] = r.r ; red[2] = r.(M^-1 r)
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_update(double *__restrict__ r, double *__restrict__ x, const double *__restrict__ q,
                                                    const double *__restrict__ p, const double *__restrict__ inv,
                                                    uint64_t num_blk, DfbSolverState *st, double *partials, unsigned int *ticket,
                                                    unsigned int *gticket, const DfbPeerView pv) {
  __shared__ double s_buf[64];
  const IterHead h = update_head(st, pv, gticket);
  if (h.stop) return;
  const double alpha = st->rho[h.parity] / h.pq;
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
    if (pv.nranks > 1 && pv.post) peer_post_block<2>(pv, dfb_xid_rr(st->xbase, h.it), s_buf);
  }
}

// what K3 needs: iteration index (from iter_dir), stop flag, (all-reduced) r.r / r.z; thread 0 of block 0 does the bookkeeping
struct DirHead {
  bool stop;
  double beta;
};
__device__ __forceinline__ DirHead direction_head(DfbSolverState *st, double *hist, uint64_t hist_ring, const DfbPeerView &pv,
                                                  unsigned int *gticket, bool gather) {
  DirHead h;
  const int it = st->iter_dir, parity = it & 1;
  h.stop = st->done[parity] != 0;
  h.beta = 0.0;
  if (h.stop) {
    // stopped earlier: propagate the flag into the slot the next iteration reads, change nothing else
    if (blockIdx.x == 0 && threadIdx.x == 0) st->done[parity ^ 1] = 1;
    return h;
  }
  double rr = st->red[1], rz = st->red[2];
  if (pv.nranks > 1 && gather) {
    double g[2];
    peer_gather_grid<2>(pv, dfb_xid_rr(st->xbase, it), &st->red[1], &st->gflag[1], gticket, &st->err, g);
    rr = g[0];
    rz = g[1];
  }
  h.beta = rz / st->rho[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int k = it + 1;
    st->iter = k;
    hist[(uint64_t)k % hist_ring] = rr;
    const double conv_ratio = sqrt(rr * st->inv_rr0);
    int stop = 0;
    if (conv_ratio < st->tol) {
      st->converged = 1;
      stop = 1;
    }
    if (st->check_pap && !(st->red[0] >= 0.0)) {
      st->err = 1;
      stop = 1;
    }
    if (k >= st->max_it) stop = 1;
    st->done[parity ^ 1] = stop;
    st->rho[parity ^ 1] = rz;
  }
  return h;
}

// K3: beta = r.z / rho ; p = M^-1 r + beta p
template <int N, int KIND>
__global__ void __launch_bounds__(256) k_pcg_direction(const double *__restrict__ r, double *__restrict__ p,
                                                       const double *__restrict__ inv, uint64_t num_blk, DfbSolverState *st,
                                                       double *hist, uint64_t hist_ring, unsigned int *gticket, const DfbPeerView pv,
                                                       int gather) {
  const DirHead h = direction_head(st, hist, hist_ring, pv, gticket, gather != 0);
  if (h.stop) return;
  const double beta = h.beta;
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
                                                         uint64_t n, DfbSolverState *st, double *partials, unsigned int *ticket,
                                                         unsigned int *gticket, const DfbPeerView pv) {
  __shared__ double s_buf[64];
  const IterHead h = update_head(st, pv, gticket);
  if (h.stop) return;
  const double alpha = st->rho[h.parity] / h.pq;
  double acc[2] = {0.0, 0.0};
  const uint64_t n2 = n >> 1;
  double2 *r2 = reinterpret_cast<double2 *>(r);
  double2 *x2 = reinterpret_cast<double2 *>(x);
  const double2 *q2 = reinterpret_cast<const double2 *>(q);
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  const double2 *i2 = reinterpret_cast<const double2 *>(inv);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 rv = r2[i], xv = x2[i];
    const double2 qv = q2[i], pv2 = p2[i];
    rv.x -= alpha * qv.x;
    rv.y -= alpha * qv.y;
    xv.x += alpha * pv2.x;
    xv.y += alpha * pv2.y;
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
    if (pv.nranks > 1 && pv.post) peer_post_block<2>(pv, dfb_xid_rr(st->xbase, h.it), s_buf);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_pcg_direction_flat(const double *__restrict__ r, double *__restrict__ p,
                                                            const double *__restrict__ inv, uint64_t n, DfbSolverState *st,
                                                            double *hist, uint64_t hist_ring, unsigned int *gticket,
                                                            const DfbPeerView pv) {
  const DirHead h = direction_head(st, hist, hist_ring, pv, gticket, true);
  if (h.stop) return;
  const double beta = h.beta;
  const uint64_t n2 = n >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  double2 *p2 = reinterpret_cast<double2 *>(p);
  const double2 *i2 = reinterpret_cast<const double2 *>(inv);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 z = r2[i], pv2 = p2[i];
    if (KIND == DFB_PRECOND_JACOBI) {
      const double2 iv = i2[i];
      z.x *= iv.x;
      z.y *= iv.y;
    }
    pv2.x = z.x + beta * pv2.x;
    pv2.y = z.y + beta * pv2.y;
    p2[i] = pv2;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint64_t i = n - 1;
    const double z = (KIND == DFB_PRECOND_JACOBI) ? inv[i] * r[i] : r[i];
    p[i] = z + beta * p[i];
  }
}

static int vec_grid(const dfb_ctx *c, uint64_t num_blk) {
  uint64_t g = (num_blk + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 8;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  if (g > DFB_RED_MAX_BLOCKS) g = DFB_RED_MAX_BLOCKS;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------ driver
// everything one iteration's launches depend on; also the key of the cached chunk graph
struct IterPlan {
  const dfb_bsr *m;
  const dfb_precond *pc;
  double *r, *x, *q, *p;
  const double *inv;
  double *zbuf;
  uint64_t nb, nflat;
  int n, kind, G, GF;
  int peer, nccl_red, peer_halo, ilu_levels;
  int64_t spmv_variant, chunk;
  uint64_t hist_ring;
  const void *packed;  // the packed mirror the SpMV nodes were captured with
  const void *hist;
  uint64_t epoch;      // dfb_free_epoch() at capture time (any device free since then invalidates the graph)
};

struct DfbGraphCache {
  IterPlan key;
  cudaGraphExec_t exec = nullptr;
  uint64_t nodes = 0;  // kernel launches one replay stands for
};

void dfb_graph_cache_destroy(dfb_ctx *c) {
  if (!c->graph) return;
  if (c->graph->exec) cudaGraphExecDestroy(c->graph->exec);
  delete c->graph;
  c->graph = nullptr;
}

// enqueue ONE iteration on c->stream (plain launch or under stream capture: identical arguments every time)
static dfb_status enqueue_iteration(dfb_ctx *c, const IterPlan &P, bool profile) {
  DfbSolverState *st = c->d_state;
  const DfbPeerView pv = dfb_peer_view(c, 1);
  DfbPeerView pv_local = pv;   // consumers that read st->red[] as it is (single GPU, or after ncclAllReduce)
  pv_local.nranks = 1;
  const DfbPeerView &pvk = P.peer ? pv : pv_local;
  const int n = P.n, kind = P.kind;
  const bool ilu = kind == DFB_PRECOND_ILU0;
  const bool prof = profile ? dfb_prof_begin(c) : false;
  DFB_TRY(dfb_spmv_full(P.m, P.q, 0.0, 1.0, P.p, 1, 1, P.peer));
  if (prof) dfb_prof_end(c);
  if (P.nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[0], 1, c->stream));
  unsigned int *tk = c->d_ticket + 2, *gt_u = c->d_ticket + 6, *gt_d = c->d_ticket + 7;
  if (kind == DFB_PRECOND_BLOCK_JACOBI) {
#define UPD(NV, KV) DFB_LAUNCH(c, (k_pcg_update<NV, KV>), P.G, 256, 0, P.r, P.x, P.q, P.p, P.inv, P.nb, st, c->d_partials, tk, gt_u, pvk)
    DISPATCH_ALL(UPD)
#undef UPD
  } else if (kind == DFB_PRECOND_JACOBI) {
    DFB_LAUNCH(c, k_pcg_update_flat<DFB_PRECOND_JACOBI>, P.GF, 256, 0, P.r, P.x, P.q, P.p, P.inv, P.nflat, st, c->d_partials, tk, gt_u, pvk);
  } else {
    DFB_LAUNCH(c, k_pcg_update_flat<DFB_PRECOND_NONE>, P.GF, 256, 0, P.r, P.x, P.q, P.p, P.inv, P.nflat, st, c->d_partials, tk, gt_u, pvk);
  }
  if (ilu) {
    // the update kernel ran with M = I (r.r in red[1]); now z = M^-1 r explicitly and r.z -> red[2]
    DFB_CUDA(cudaMemcpyAsync(P.zbuf, P.r, P.nflat * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    DFB_TRY(dfb_ilu_apply(c, P.pc->ilu, P.pc->pat, n, P.pc->d_inv, P.zbuf));
    DFB_TRY(dfb_ilu_dot(c, P.r, P.zbuf, P.nflat, &st->red[2]));
  }
  if (P.nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[1], 2, c->stream));
  if (kind == DFB_PRECOND_BLOCK_JACOBI) {
#define DIR(NV, KV) DFB_LAUNCH(c, (k_pcg_direction<NV, KV>), P.G, 256, 0, P.r, P.p, P.inv, P.nb, st, c->d_hist, P.hist_ring, gt_d, pvk, 1)
    DISPATCH_ALL(DIR)
#undef DIR
  } else if (kind == DFB_PRECOND_JACOBI) {
    DFB_LAUNCH(c, k_pcg_direction_flat<DFB_PRECOND_JACOBI>, P.GF, 256, 0, P.r, P.p, P.inv, P.nflat, st, c->d_hist, P.hist_ring, gt_d, pvk);
  } else if (ilu) {
#define DIRZ(NV) \
  if (n == NV) DFB_LAUNCH(c, (k_pcg_direction<NV, DFB_PRECOND_ILU0>), P.G, 256, 0, P.r, P.p, P.zbuf, P.nb, st, c->d_hist, P.hist_ring, gt_d, pv_local, 0);
    DIRZ(1) DIRZ(2) DIRZ(3) DIRZ(4)
#undef DIRZ
  } else {
    DFB_LAUNCH(c, k_pcg_direction_flat<DFB_PRECOND_NONE>, P.GF, 256, 0, P.r, P.p, P.inv, P.nflat, st, c->d_hist, P.hist_ring, gt_d, pvk);
  }
  // the boundary entries of the new direction go straight into the neighbours' landing zones
  if (P.peer_halo) DFB_TRY(dfb_halo_push(P.m->halo, P.p, n, c->stream));
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// capture `chunk` iterations into a graph (cached per ctx, keyed by every pointer / size baked into the nodes)
static dfb_status chunk_graph(dfb_ctx *c, const IterPlan &P, DfbGraphCache **out) {
  *out = nullptr;
  if (c->graph && c->graph->exec && memcmp(&c->graph->key, &P, sizeof(IterPlan)) == 0) {
    *out = c->graph;
    return DFB_OK;
  }
  dfb_graph_cache_destroy(c);
  const uint64_t launches0 = c->launches;
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGetLastError();
    return DFB_OK;  // caller falls back to plain launches
  }
  dfb_status st = DFB_OK;
  for (int64_t j = 0; j < P.chunk && st == DFB_OK; ++j) st = enqueue_iteration(c, P, false);
  const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  const uint64_t nodes = c->launches - launches0;
  c->launches = launches0;  // capturing launched nothing
  if (st != DFB_OK || e != cudaSuccess || !graph) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    return DFB_OK;
  }
  cudaGraphExec_t exec = nullptr;
  if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    cudaGraphDestroy(graph);
    return DFB_OK;
  }
  cudaGraphDestroy(graph);
  c->graph = new DfbGraphCache();
  c->graph->key = P;
  c->graph->exec = exec;
  c->graph->nodes = nodes;
  *out = c->graph;
  return DFB_OK;
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
  if (m->halo && c->nranks > 1) {
    const uint64_t need = m->halo->recv_ptr.empty() ? 0 : m->halo->recv_ptr.back();
    DFB_REQUIRE(p->n_ghost >= need, "(p)cg: the matrix has a halo of %llu ghost blocks: create p_vec with n_ghost >= that (it has %llu)",
                (unsigned long long)need, (unsigned long long)p->n_ghost);
  }
  if (pc) DFB_REQUIRE(pc->num_blk == nb && pc->blk_dim == n, "pcg: preconditioner shape != matrix shape");
  DFB_REQUIRE(max_it < 0x7FFFFFF0ull, "(p)cg: max_it too large");
  DFB_CUDA(cudaSetDevice(c->device));

  IterPlan P;
  memset(&P, 0, sizeof(P));
  P.m = m;
  P.pc = pc;
  P.r = r->d;
  P.x = x->d;
  P.q = q->d;
  P.p = p->d;
  P.inv = inv;
  P.zbuf = zbuf;
  P.nb = nb;
  P.nflat = nb * (uint64_t)n;
  P.n = n;
  P.kind = kind;
  P.G = vec_grid(c, nb);
  P.GF = vec_grid(c, (P.nflat + 1) / 2);
  // scalar all-reduces: fused into the kernels through the NVLink peer mailboxes when mapped, NCCL otherwise
  P.peer = (c->peer_enabled && c->use_peer_allreduce && c->nranks > 1) ? 1 : 0;
  P.nccl_red = (c->nranks > 1 && !P.peer) ? 1 : 0;
  P.peer_halo = (P.peer && dfb_spmv_uses_peer_halo(m)) ? 1 : 0;
  P.spmv_variant = c->spmv_variant;
  P.ilu_levels = (int)c->ilu_level_launches;
  P.chunk = c->iters_per_sync > 0 ? c->iters_per_sync : 16;
  if (P.chunk > 256) P.chunk = 256;
  // ring of ||r_k||^2: the host drains it after every chunk, so 2 chunks + the plain head iteration always fit
  P.hist_ring = (uint64_t)(2 * P.chunk + 4);
  if (c->hist_cap < P.hist_ring) {
    dfb_dev_free(c->d_hist);
    c->d_hist = nullptr;
    c->hist_cap = 0;
    DFB_TRY(dfb_dev_alloc_t(&c->d_hist, P.hist_ring));
    c->hist_cap = P.hist_ring;
  }
  const size_t slot_bytes = ((sizeof(DfbSolverState) + 15) & ~(size_t)15) + P.hist_ring * sizeof(double);
  if (c->h_poll_slot < slot_bytes) {
    if (c->h_poll) cudaFreeHost(c->h_poll);
    c->h_poll = nullptr;
    c->h_poll_slot = 0;
    DFB_CUDA(cudaMallocHost(&c->h_poll, 2 * slot_bytes));
    c->h_poll_slot = slot_bytes;
  }
  DFB_TRY(dfb_bsr_ensure_packed(m));
  P.packed = m->d_packed;
  P.hist = c->d_hist;
  P.epoch = dfb_free_epoch();
  DfbSolverState *st = c->d_state;
  const DfbPeerView pv = dfb_peer_view(c, P.peer);
  DfbPeerView pv_i = pv;
  if (!P.peer) pv_i.nranks = 1;

  if (ilu) {
    // z = M^-1 r by the level-scheduled sweeps, then the generic init with z taken as given
    DFB_CUDA(cudaMemcpyAsync(zbuf, r->d, P.nflat * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    DFB_TRY(dfb_ilu_apply(c, pc->ilu, pc->pat, n, pc->d_inv, zbuf));
#define INITZ(NV) \
  if (n == NV) DFB_LAUNCH(c, (k_pcg_init<NV, DFB_PRECOND_ILU0>), P.G, 256, 0, r->d, x->d, p->d, zbuf, nb, st, c->d_partials, c->d_ticket + 2, pv_i);
    INITZ(1) INITZ(2) INITZ(3) INITZ(4)
#undef INITZ
  }
#define INIT(NV, KV) DFB_LAUNCH(c, (k_pcg_init<NV, KV>), P.G, 256, 0, r->d, x->d, p->d, inv, nb, st, c->d_partials, c->d_ticket + 2, pv_i)
  DISPATCH_ALL(INIT)
#undef INIT
  DFB_CHECK_LAUNCH();
  if (P.nccl_red) DFB_TRY(dfb_nccl_allreduce_sum(c, &st->red[1], 2, c->stream));
  DFB_LAUNCH(c, k_pcg_init_finalize, 1, 32, 0, st, c->d_hist, tol, eps_sq, check_pap, (int)max_it, pv_i);
  DFB_CHECK_LAUNCH();
  if (P.peer_halo) DFB_TRY(dfb_halo_push(m->halo, p->d, n, c->stream));  // ghosts of the first direction

  // ---- chunks: [one plain iteration (bracketed with events when profiling)] + [graph replay of `chunk` iterations]
  std::vector<double> &rr = c->last_hist;
  rr.assign(1, 0.0);
  // (with one launch per level an ILU application is ~900 launches: far too many nodes per iteration)
  const bool use_graph = c->use_graph != 0 && !(ilu && c->ilu_level_launches);
  DfbGraphCache *gc = nullptr;
  bool graph_tried = false;
  DfbSolverState last;
  memset(&last, 0, sizeof(last));
  uint64_t enq = 0;
  int n_chunks = 0;
  uint64_t drained = 0;  // history entries 1..drained are in rr
  auto drain = [&](int slot) {
    const unsigned char *base = (const unsigned char *)c->h_poll + (size_t)slot * c->h_poll_slot;
    memcpy(&last, base, sizeof(DfbSolverState));
    const double *ring = (const double *)(base + ((sizeof(DfbSolverState) + 15) & ~(size_t)15));
    rr[0] = last.rr0;
    for (uint64_t k = drained + 1; k <= (uint64_t)last.iter; ++k) rr.push_back(ring[k % P.hist_ring]);
    drained = (uint64_t)last.iter;
  };
  bool finished = max_it == 0;
  while (!finished) {
    const bool head = (n_chunks == 0) || c->profile_spmv > 0 || !use_graph;
    if (head) {
      DFB_TRY(enqueue_iteration(c, P, c->profile_spmv > 0));
      enq += 1;
    }
    if (use_graph && !graph_tried) {
      // captured AFTER the first plain iteration (function attributes set, NCCL connections up, mirror packed)
      graph_tried = true;
      DFB_TRY(chunk_graph(c, P, &gc));
    }
    if (use_graph && gc) {
      DFB_CUDA(cudaGraphLaunch(gc->exec, c->stream));
      c->launches += gc->nodes;
      enq += (uint64_t)P.chunk;
    } else {
      for (int64_t j = head ? 1 : 0; j < P.chunk; ++j) DFB_TRY(enqueue_iteration(c, P, false));
      enq += (uint64_t)(P.chunk - (head ? 1 : 0));
    }
    const int slot = n_chunks & 1;
    unsigned char *hs = (unsigned char *)c->h_poll + (size_t)slot * c->h_poll_slot;
    DFB_CUDA(cudaMemcpyAsync(hs, st, sizeof(DfbSolverState), cudaMemcpyDeviceToHost, c->stream));
    DFB_CUDA(cudaMemcpyAsync(hs + ((sizeof(DfbSolverState) + 15) & ~(size_t)15), c->d_hist, P.hist_ring * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    DFB_CUDA(cudaEventRecord(c->ev_flag[slot], c->stream));
    ++n_chunks;
    // look at the state after the PREVIOUS chunk (one chunk of look-ahead keeps the GPU busy)
    if (n_chunks >= 2) {
      const int prev = (n_chunks - 2) & 1;
      DFB_CUDA(cudaEventSynchronize(c->ev_flag[prev]));
      drain(prev);
      if (last.done[0] || last.done[1]) finished = true;
    }
    if (enq >= max_it + (uint64_t)P.chunk) finished = true;  // the device stops at max_it by itself; this only ends the enqueueing
  }
  DFB_LAUNCH(c, k_pcg_end, 1, 1, 0, st);
  DFB_CHECK_LAUNCH();
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  if (n_chunks > 0) {
    if (n_chunks >= 2 && !(last.done[0] || last.done[1])) drain((n_chunks - 2) & 1);
    drain((n_chunks - 1) & 1);
  } else {
    DFB_CUDA(cudaMemcpyAsync(c->h_poll, st, sizeof(DfbSolverState), cudaMemcpyDeviceToHost, c->stream));
    DFB_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(&last, c->h_poll, sizeof(DfbSolverState));
    rr[0] = last.rr0;
  }
  DFB_TRY(dfb_prof_flush(c, c->profile_spmv > 0 ? (size_t)((last.iter + P.chunk) / (P.chunk + 1)) : 0));
  if (last.err == 2) return dfb_fail(DFB_ERR_NCCL, "(p)cg: peer exchange timed out (a rank is missing or out of step)");
  if (last.err) return dfb_fail(DFB_ERR_NOT_POSITIVE, "cg: p.Ap < 0 (assert!(pap >= T::zero()))");

  // history: rr[k] = ||r_k||^2 for k = 0..iter
  const uint64_t iters = (uint64_t)last.iter;
  const double rr0 = rr[0];
  const bool converged = last.converged != 0;
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

// ||r_k||^2 (k = 0..iterations) of the last (P)CG solve on this ctx, kept on the host: lets a caller size its history buffer from
// the returned length instead of from max_it
DFB_API dfb_status dfb_ctx_last_history_sq(dfb_ctx *c, double *out, uint64_t cap, uint64_t *n) {
  DFB_REQUIRE(c && n, "dfb_ctx_last_history_sq: NULL argument");
  *n = c->last_hist.size();
  if (out)
    for (uint64_t k = 0; k < cap && k < c->last_hist.size(); ++k) out[k] = c->last_hist[k];
  return DFB_OK;
}
