// linearsystem.cu — `linearsystem::Solver` of del-fem-ls (del-fem-ls/src/linearsystem.rs:8-112) behind the C ABI, one-to-one:
//   new :33   initialize :48   begin_merge :60   end_merge :67   solve_cg :73   solve_pcg :85   clone :96
// The struct owns what the reference's struct owns (sparse, ilu, r_vec, u_vec, ap_vec, p_vec, conv, conv_ratio_tol,
// max_num_iteration), all device-resident; the fields are reachable as borrowed handles so that the caller merges element
// matrices / applies boundary conditions between begin_merge and end_merge exactly like reference code does on `solver.sparse`
// and `solver.r_vec`.  Generalised from the reference's scalar T to blk_dim 1..4.
#include <math.h>

#include <vector>

#include "dfb_internal.h"

dfb_status dfb_ilu_copy_values(dfb_ctx *c, DfbIlu *dst, const DfbIlu *src, int NN);  // ilu.cu

struct dfb_solver {
  dfb_ctx *ctx = nullptr;
  int blk_dim = 1;
  dfb_pattern *pat = nullptr;
  dfb_bsr *sparse = nullptr;
  dfb_precond *ilu = nullptr;
  int ilu_kind = DFB_PRECOND_ILU0;   // Solver::initialize sets up ILU(0) (:53)
  uint64_t ilu_lev_fill = 0;
  int ilu_decomposed = 0;
  dfb_vec *r_vec = nullptr, *u_vec = nullptr, *ap_vec = nullptr, *p_vec = nullptr;
  std::vector<double> conv;
  double conv_ratio_tol = (double)1.0e-5f;  // 1.0e-5_f32.as_() (:43)
  uint64_t max_num_iteration = 100;         // (:44)
};

static const double EPS_LS = (double)1.0e-20f;  // del-fem-ls/src/solver_sparse.rs:40,108

DFB_API dfb_status dfb_solver_new(dfb_ctx *ctx, int32_t blk_dim, dfb_solver **out) {
  DFB_REQUIRE(ctx && out, "dfb_solver_new: NULL argument");
  DFB_REQUIRE(blk_dim >= 1 && blk_dim <= 4, "dfb_solver_new: blk_dim %d not in 1..4", blk_dim);
  dfb_solver *s = new dfb_solver();
  s->ctx = ctx;
  s->blk_dim = blk_dim;
  *out = s;
  return DFB_OK;
}

static void solver_release(dfb_solver *s) {
  dfb_precond_destroy(s->ilu);
  dfb_bsr_destroy(s->sparse);
  dfb_pattern_destroy(s->pat);
  dfb_vec_destroy(s->r_vec);
  dfb_vec_destroy(s->u_vec);
  dfb_vec_destroy(s->ap_vec);
  dfb_vec_destroy(s->p_vec);
  s->ilu = nullptr;
  s->sparse = nullptr;
  s->pat = nullptr;
  s->r_vec = s->u_vec = s->ap_vec = s->p_vec = nullptr;
  s->ilu_decomposed = 0;
}

DFB_API dfb_status dfb_solver_destroy(dfb_solver *s) {
  if (!s) return DFB_OK;
  solver_release(s);
  delete s;
  return DFB_OK;
}

static dfb_status solver_make_precond(dfb_solver *s) {
  if (s->ilu_kind == DFB_PRECOND_ILU0 && s->ilu_lev_fill > 0) return dfb_precond_create_iluk(s->ctx, s->sparse, s->ilu_lev_fill, &s->ilu);
  return dfb_precond_create(s->ctx, s->ilu_kind, s->sparse, &s->ilu);
}

// Solver::initialize(colind, rowptr) — NB the reference's parameter NAMES are swapped relative to their meaning (:48-51): the FIRST
// array is the row pointer (num_blk + 1 entries), the second the column indices. Positional semantics are kept.
DFB_API dfb_status dfb_solver_initialize(dfb_solver *s, const uint64_t *colind, uint64_t colind_len, const uint64_t *rowptr,
                                         uint64_t rowptr_len) {
  DFB_REQUIRE(s && colind, "dfb_solver_initialize: NULL argument");
  DFB_REQUIRE(colind_len >= 1, "dfb_solver_initialize: the row-pointer array is empty (colind.len() - 1 underflows)");
  const uint64_t nblk = colind_len - 1;  // let nblk = colind.len() - 1 (:51)
  DFB_REQUIRE(colind[nblk] == rowptr_len, "dfb_solver_initialize: row pointer end %llu != number of column indices %llu",
              (unsigned long long)colind[nblk], (unsigned long long)rowptr_len);
  solver_release(s);
  dfb_status st = dfb_pattern_create(s->ctx, colind, rowptr, nblk, 0, &s->pat);  // symbolic_initialization (:49)
  if (st == DFB_OK) st = dfb_bsr_create(s->ctx, s->pat, s->blk_dim, &s->sparse);
  if (st == DFB_OK) st = dfb_vec_create(s->ctx, nblk, 0, s->blk_dim, &s->r_vec);  // r_vec.resize(nblk) (:52)
  if (st == DFB_OK) st = dfb_vec_create(s->ctx, nblk, 0, s->blk_dim, &s->u_vec);
  if (st == DFB_OK) st = dfb_vec_create(s->ctx, nblk, 0, s->blk_dim, &s->ap_vec);
  if (st == DFB_OK) st = dfb_vec_create(s->ctx, nblk, 0, s->blk_dim, &s->p_vec);
  if (st == DFB_OK) st = solver_make_precond(s);  // self.ilu.initialize_ilu0(&self.sparse) (:53)
  if (st != DFB_OK) solver_release(s);
  return st;
}

// the alternatives the reference keeps commented out next to initialize_ilu0 (:50,:54-56) plus the derived Jacobi kinds:
// kind = DFB_PRECOND_ILU0 with lev_fill (0 = ILU(0), large = complete LU), DFB_PRECOND_JACOBI, DFB_PRECOND_BLOCK_JACOBI, DFB_PRECOND_NONE
DFB_API dfb_status dfb_solver_set_preconditioner(dfb_solver *s, int32_t kind, uint64_t lev_fill) {
  DFB_REQUIRE(s, "dfb_solver_set_preconditioner: NULL argument");
  DFB_REQUIRE(kind >= DFB_PRECOND_NONE && kind <= DFB_PRECOND_ILU0, "dfb_solver_set_preconditioner: unknown kind %d", kind);
  s->ilu_kind = kind;
  s->ilu_lev_fill = kind == DFB_PRECOND_ILU0 ? lev_fill : 0;
  if (s->sparse) {
    dfb_precond_destroy(s->ilu);
    s->ilu = nullptr;
    s->ilu_decomposed = 0;
    DFB_TRY(solver_make_precond(s));
  }
  return DFB_OK;
}

DFB_API dfb_status dfb_solver_begin_merge(dfb_solver *s) {
  DFB_REQUIRE(s && s->sparse, "dfb_solver_begin_merge: call dfb_solver_initialize first");
  DFB_TRY(dfb_bsr_set_zero(s->sparse));  // self.sparse.set_zero() (:61)
  return dfb_vec_set_zero(s->r_vec);     // r_vec.resize + set_zero (:62-64)
}

DFB_API dfb_status dfb_solver_end_merge(dfb_solver *s) {
  DFB_REQUIRE(s && s->sparse, "dfb_solver_end_merge: call dfb_solver_initialize first");
  DFB_TRY(dfb_precond_update(s->ilu, s->sparse));  // copy_value + decompose (:69-70)
  s->ilu_decomposed = 1;
  return DFB_OK;
}

static dfb_status solver_run(dfb_solver *s, bool pcg) {
  DFB_REQUIRE(s && s->sparse, "dfb_solver_solve: call dfb_solver_initialize first");
  const uint64_t cap = s->max_num_iteration + 3;
  std::vector<double> hist((size_t)(cap < 4096 ? cap : 4096));
  uint64_t n = 0;
  if (pcg) DFB_TRY(dfb_pcg(s->sparse, s->ilu, s->r_vec, s->u_vec, s->ap_vec, s->p_vec, s->conv_ratio_tol, s->max_num_iteration, EPS_LS,
                           hist.data(), hist.size(), &n));
  else DFB_TRY(dfb_cg(s->sparse, s->r_vec, s->u_vec, s->ap_vec, s->p_vec, s->conv_ratio_tol, s->max_num_iteration, EPS_LS, 1, hist.data(),
                      hist.size(), &n));
  if (n > hist.size()) {
    // longer than the first guess: rebuild from the library's host copy of ||r_k||^2
    uint64_t nsq = 0;
    DFB_TRY(dfb_ctx_last_history_sq(s->ctx, nullptr, 0, &nsq));
    std::vector<double> sq((size_t)nsq);
    DFB_TRY(dfb_ctx_last_history_sq(s->ctx, sq.data(), nsq, &nsq));
    hist.assign((size_t)n, 0.0);
    if (pcg) {
      for (uint64_t k = 0; k < nsq && k < n; ++k) hist[k] = sqrt(sq[k]);
      if (n == nsq + 1 && nsq > 0) hist[n - 1] = sqrt(sq[nsq - 1]);
    } else {
      const double inv0 = nsq > 0 ? 1.0 / sq[0] : 0.0;
      for (uint64_t k = 1; k < nsq && k - 1 < n; ++k) hist[k - 1] = sqrt(sq[k] * inv0);
    }
  }
  hist.resize((size_t)n);
  s->conv.swap(hist);
  return DFB_OK;
}

// solver_sparse::conjugate_gradient (del-fem-ls/src/solver_sparse.rs:5-70: eps 1e-20, assert!(pap >= 0)) on the owned vectors
DFB_API dfb_status dfb_solver_solve_cg(dfb_solver *s) { return solver_run(s, false); }
// solver_sparse::preconditioned_conjugate_gradient (:73-175) with self.ilu
DFB_API dfb_status dfb_solver_solve_pcg(dfb_solver *s) {
  DFB_REQUIRE(s && s->ilu_decomposed, "dfb_solver_solve_pcg: call dfb_solver_end_merge first (the preconditioner has no values)");
  return solver_run(s, true);
}

static dfb_status vec_clone(const dfb_vec *src, dfb_vec **out) {
  DFB_TRY(dfb_vec_create(src->ctx, src->n_blk, src->n_ghost, src->blk_dim, out));
  return dfb_vec_copy(*out, src);
}

// Solver::clone (:96-111): deep copy of every field, EXCEPT max_num_iteration which the reference resets to 100 (:109)
DFB_API dfb_status dfb_solver_clone(const dfb_solver *s, dfb_solver **out) {
  DFB_REQUIRE(s && out, "dfb_solver_clone: NULL argument");
  dfb_solver *t = new dfb_solver();
  t->ctx = s->ctx;
  t->blk_dim = s->blk_dim;
  t->ilu_kind = s->ilu_kind;
  t->ilu_lev_fill = s->ilu_lev_fill;
  t->conv = s->conv;
  t->conv_ratio_tol = s->conv_ratio_tol;
  t->max_num_iteration = 100;
  dfb_status st = DFB_OK;
  if (s->sparse) {
    // the pattern is immutable: the clone shares it (reference: Vec clone of identical content)
    t->pat = s->pat;
    t->pat->refcount += 1;
    st = dfb_bsr_create(t->ctx, t->pat, t->blk_dim, &t->sparse);
    if (st == DFB_OK) st = dfb_bsr_canon_read(s->sparse);
    if (st == DFB_OK) st = dfb_bsr_canon_overwrite(t->sparse);
    if (st == DFB_OK) {
      const uint64_t nn = (uint64_t)s->blk_dim * s->blk_dim;
      cudaError_t e = cudaMemcpyAsync(t->sparse->d_idx2val, s->sparse->d_idx2val, s->pat->nnz * nn * sizeof(double), cudaMemcpyDeviceToDevice, s->ctx->stream);
      if (e == cudaSuccess && s->sparse->d_row2val)
        e = cudaMemcpyAsync(t->sparse->d_row2val, s->sparse->d_row2val, s->pat->num_blk * nn * sizeof(double), cudaMemcpyDeviceToDevice, s->ctx->stream);
      if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_solver_clone: %s", cudaGetErrorString(e));
    }
    if (st == DFB_OK) st = vec_clone(s->r_vec, &t->r_vec);
    if (st == DFB_OK) st = vec_clone(s->u_vec, &t->u_vec);
    if (st == DFB_OK) st = vec_clone(s->ap_vec, &t->ap_vec);
    if (st == DFB_OK) st = vec_clone(s->p_vec, &t->p_vec);
    if (st == DFB_OK) st = solver_make_precond(t);
    if (st == DFB_OK && s->ilu_decomposed) {
      // copy the factor values as they are (the reference clones the decomposed ILU, it does not re-decompose)
      const uint64_t per = s->ilu->kind == DFB_PRECOND_JACOBI ? (uint64_t)s->blk_dim : (uint64_t)s->blk_dim * s->blk_dim;
      if (s->ilu->d_inv) {
        cudaError_t e = cudaMemcpyAsync(t->ilu->d_inv, s->ilu->d_inv, s->ilu->num_blk * per * sizeof(double), cudaMemcpyDeviceToDevice, s->ctx->stream);
        if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_solver_clone: %s", cudaGetErrorString(e));
      }
      if (st == DFB_OK && s->ilu->ilu) st = dfb_ilu_copy_values(s->ctx, t->ilu->ilu, s->ilu->ilu, s->blk_dim * s->blk_dim);
      if (st == DFB_OK) t->ilu_decomposed = 1;
    }
  }
  if (st != DFB_OK) {
    dfb_solver_destroy(t);
    return st;
  }
  *out = t;
  return DFB_OK;
}

// ---- field access (borrowed handles, owned by the solver)
DFB_API dfb_bsr *dfb_solver_sparse(dfb_solver *s) { return s ? s->sparse : nullptr; }
DFB_API dfb_precond *dfb_solver_ilu(dfb_solver *s) { return s ? s->ilu : nullptr; }
DFB_API dfb_vec *dfb_solver_r_vec(dfb_solver *s) { return s ? s->r_vec : nullptr; }
DFB_API dfb_vec *dfb_solver_u_vec(dfb_solver *s) { return s ? s->u_vec : nullptr; }
DFB_API dfb_vec *dfb_solver_ap_vec(dfb_solver *s) { return s ? s->ap_vec : nullptr; }
DFB_API dfb_vec *dfb_solver_p_vec(dfb_solver *s) { return s ? s->p_vec : nullptr; }

DFB_API dfb_status dfb_solver_set_params(dfb_solver *s, double conv_ratio_tol, uint64_t max_num_iteration) {
  DFB_REQUIRE(s, "dfb_solver_set_params: NULL argument");
  s->conv_ratio_tol = conv_ratio_tol;
  s->max_num_iteration = max_num_iteration;
  return DFB_OK;
}

DFB_API dfb_status dfb_solver_get_params(const dfb_solver *s, double *conv_ratio_tol, uint64_t *max_num_iteration) {
  DFB_REQUIRE(s, "dfb_solver_get_params: NULL argument");
  if (conv_ratio_tol) *conv_ratio_tol = s->conv_ratio_tol;
  if (max_num_iteration) *max_num_iteration = s->max_num_iteration;
  return DFB_OK;
}

// `conv` (the history the last solve returned): *n = conv.len() even if it exceeds cap
DFB_API dfb_status dfb_solver_conv(const dfb_solver *s, double *out, uint64_t cap, uint64_t *n) {
  DFB_REQUIRE(s && n, "dfb_solver_conv: NULL argument");
  *n = s->conv.size();
  if (out)
    for (uint64_t k = 0; k < cap && k < s->conv.size(); ++k) out[k] = s->conv[k];
  return DFB_OK;
}
