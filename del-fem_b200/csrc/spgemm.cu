// spgemm.cu — C = M0 * M1 for two scalar "CSR + separate diagonal" matrices, restating del-fem-ls
// sparse_matrix_multiplication.rs: symbolic_multiplication :7-107 (pattern of C in discovery order, diagonal kept apart) and
// mult_square_matrices :109-188 (values).  The reference walks the rows serially with a dense col2flag / col2idx scratch; here
// every row is independent: ONE thread per row repeats the reference's loop order (so the column order of C and — compiled
// --fmad=false — every sum are bit-identical), with a linear search in the row's own (short) column list instead of the dense
// scratch.  Off the hot path (Galerkin products / AMG set-up): simple on purpose, three passes:
//   bound (upper bound of each row's length) -> scan -> symbolic (discover columns into the bounded scratch) -> scan -> compact + numeric.
#include "dfb_internal.h"

// visit the candidate columns of row i in the reference's order: for k in A(i): B(k)..., k ; then B(i)...   (diagonal i skipped)
template <class F>
__device__ __forceinline__ void spgemm_visit(uint32_t i, const uint32_t *__restrict__ a_r2i, const uint32_t *__restrict__ a_col,
                                             const uint32_t *__restrict__ b_r2i, const uint32_t *__restrict__ b_col, F &&f) {
  for (uint32_t ka = a_r2i[i]; ka < a_r2i[i + 1]; ++ka) {
    const uint32_t k = a_col[ka];
    for (uint32_t kb = b_r2i[k]; kb < b_r2i[k + 1]; ++kb) f(b_col[kb]);
    f(k);
  }
  for (uint32_t kb = b_r2i[i]; kb < b_r2i[i + 1]; ++kb) f(b_col[kb]);
}

__global__ void k_spgemm_bound(uint64_t n, const uint32_t *__restrict__ a_r2i, const uint32_t *__restrict__ a_col,
                               const uint32_t *__restrict__ b_r2i, uint32_t *__restrict__ ub, int *overflow) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t s = b_r2i[i + 1] - b_r2i[i];
    for (uint32_t ka = a_r2i[i]; ka < a_r2i[i + 1]; ++ka) {
      const uint32_t k = a_col[ka];
      s += (uint64_t)(b_r2i[k + 1] - b_r2i[k]) + 1;
    }
    if (s > 0x7FFFFFFFull) *overflow = 1;
    ub[i] = (uint32_t)s;
  }
}

__global__ void k_spgemm_symbolic(uint64_t n, const uint32_t *__restrict__ a_r2i, const uint32_t *__restrict__ a_col,
                                  const uint32_t *__restrict__ b_r2i, const uint32_t *__restrict__ b_col,
                                  const uint32_t *__restrict__ ub_ptr, uint32_t *__restrict__ scratch, uint32_t *__restrict__ cnt) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t *cols = scratch + ub_ptr[i];
    uint32_t c = 0;
    spgemm_visit((uint32_t)i, a_r2i, a_col, b_r2i, b_col, [&](uint32_t j) {
      if (j == (uint32_t)i) return;  // c_is_square: the diagonal stays out of the pattern
      for (uint32_t q = 0; q < c; ++q)
        if (cols[q] == j) return;
      cols[c++] = j;
    });
    cnt[i] = c;
  }
}

__global__ void k_spgemm_numeric(uint64_t n, const uint32_t *__restrict__ a_r2i, const uint32_t *__restrict__ a_col,
                                 const double *__restrict__ a_val, const double *__restrict__ a_dia,
                                 const uint32_t *__restrict__ b_r2i, const uint32_t *__restrict__ b_col,
                                 const double *__restrict__ b_val, const double *__restrict__ b_dia,
                                 const uint32_t *__restrict__ ub_ptr, const uint32_t *__restrict__ scratch,
                                 const uint32_t *__restrict__ c_r2i, uint32_t *__restrict__ c_col, double *__restrict__ c_val,
                                 double *__restrict__ c_dia) {
  for (uint64_t ii = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; ii < n; ii += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)ii;
    const uint32_t rb = c_r2i[i], len = c_r2i[i + 1] - rb;
    uint32_t *cols = c_col + rb;
    double *vals = c_val + rb;
    const uint32_t *src = scratch + ub_ptr[i];
    for (uint32_t q = 0; q < len; ++q) {
      cols[q] = src[q];
      vals[q] = 0.0;
    }
    auto slot = [&](uint32_t j) -> uint32_t {
      for (uint32_t q = 0; q < len; ++q)
        if (cols[q] == j) return q;
      return 0xFFFFFFFFu;  // cannot happen: the symbolic pass visited the same pairs
    };
    double dia = 0.0;
    for (uint32_t ka = a_r2i[i]; ka < a_r2i[i + 1]; ++ka) {
      const uint32_t k = a_col[ka];
      const double a = a_val[ka];
      for (uint32_t kb = b_r2i[k]; kb < b_r2i[k + 1]; ++kb) {
        const uint32_t j = b_col[kb];
        if (j == i) dia += a * b_val[kb];
        else vals[slot(j)] += a * b_val[kb];
      }
      if (k == i) dia += a * b_dia[k];
      else vals[slot(k)] += a * b_dia[k];
    }
    const double ad = a_dia[i];
    for (uint32_t kb = b_r2i[i]; kb < b_r2i[i + 1]; ++kb) {
      const uint32_t j = b_col[kb];
      if (j == i) dia += ad * b_val[kb];
      else vals[slot(j)] += ad * b_val[kb];
    }
    dia += ad * b_dia[i];
    c_dia[i] = dia;
  }
}

DFB_API dfb_status dfb_bsr_mult_square_matrices(const dfb_bsr *m0, const dfb_bsr *m1, dfb_pattern **out_pattern, dfb_bsr **out) {
  DFB_REQUIRE(m0 && m1 && out_pattern && out, "dfb_bsr_mult_square_matrices: NULL argument");
  DFB_REQUIRE(m0->ctx == m1->ctx, "dfb_bsr_mult_square_matrices: matrices belong to different contexts");
  DFB_REQUIRE(m0->blk_dim == 1 && m1->blk_dim == 1, "dfb_bsr_mult_square_matrices: scalar matrices only (del-fem-ls Matrix<T>)");
  DFB_REQUIRE(m0->pat->convention == 0 && m1->pat->convention == 0, "dfb_bsr_mult_square_matrices: needs the separate-diagonal convention");
  DFB_REQUIRE(m0->pat->num_blk == m1->pat->num_blk, "dfb_bsr_mult_square_matrices: assert_eq!(num_blk, m1.num_blk)");
  DFB_REQUIRE(!m0->halo && !m1->halo, "dfb_bsr_mult_square_matrices: not available for distributed matrices");
  dfb_ctx *c = m0->ctx;
  DFB_CUDA(cudaSetDevice(c->device));
  const uint64_t n = m0->pat->num_blk;
  const int G = c->sm_count * 8;
  uint32_t *d_ub = nullptr, *d_cnt = nullptr, *d_scratch = nullptr;
  dfb_pattern *pat = nullptr;
  dfb_status st = dfb_dev_alloc_t(&d_ub, n + 2);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_cnt, n + 2);
  uint64_t total_ub = 0, nnz_c = 0;
  if (st == DFB_OK) {
    cudaMemsetAsync(d_ub, 0, (n + 2) * sizeof(uint32_t), c->stream);
    cudaMemsetAsync(d_cnt, 0, (n + 2) * sizeof(uint32_t), c->stream);
    cudaMemsetAsync(c->d_errflag + 3, 0, sizeof(int), c->stream);
    DFB_LAUNCH(c, k_spgemm_bound, G, 256, 0, n, m0->pat->d_row2idx, m0->pat->d_idx2col, m1->pat->d_row2idx, d_ub, c->d_errflag + 3);
    st = dfb_check_error_flag(c, 3, DFB_ERR_INDEX_OVERFLOW, "dfb_bsr_mult_square_matrices: a row of the product is too long for u32 indices");
  }
  if (st == DFB_OK) st = dfb_exclusive_scan_u32(c, d_ub, n, &total_ub);
  if (st == DFB_OK && total_ub >= 0xFFFFFFF0ull) st = dfb_fail(DFB_ERR_INDEX_OVERFLOW, "dfb_bsr_mult_square_matrices: product too large for u32 indices");
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_scratch, total_ub + 2);
  if (st == DFB_OK) {
    DFB_LAUNCH(c, k_spgemm_symbolic, G, 128, 0, n, m0->pat->d_row2idx, m0->pat->d_idx2col, m1->pat->d_row2idx, m1->pat->d_idx2col, d_ub,
               d_scratch, d_cnt);
    if (cudaGetLastError() != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_bsr_mult_square_matrices: symbolic launch failed");
  }
  if (st == DFB_OK) st = dfb_exclusive_scan_u32(c, d_cnt, n, &nnz_c);   // d_cnt becomes C's row2idx
  if (st == DFB_OK) {
    pat = new dfb_pattern();
    pat->ctx = c;
    pat->num_blk = n;
    pat->nnz = nnz_c;
    pat->convention = 0;
    pat->refcount = 1;
    pat->d_row2idx = d_cnt;
    d_cnt = nullptr;
    pat->d_idx2col = nullptr;
    st = dfb_dev_alloc_t(&pat->d_idx2col, nnz_c + 8);
  }
  dfb_bsr *cm = nullptr;
  if (st == DFB_OK) st = dfb_bsr_create(c, pat, 1, &cm);
  if (st == DFB_OK) st = dfb_bsr_canon_overwrite(cm);
  if (st == DFB_OK) st = dfb_bsr_canon_read(m0);
  if (st == DFB_OK) st = dfb_bsr_canon_read(m1);
  if (st == DFB_OK) {
    DFB_LAUNCH(c, k_spgemm_numeric, G, 128, 0, n, m0->pat->d_row2idx, m0->pat->d_idx2col, m0->d_idx2val, m0->d_row2val, m1->pat->d_row2idx,
               m1->pat->d_idx2col, m1->d_idx2val, m1->d_row2val, d_ub, d_scratch, pat->d_row2idx, pat->d_idx2col, cm->d_idx2val,
               cm->d_row2val);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_bsr_mult_square_matrices: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(c->stream);
  dfb_dev_free(d_ub);
  dfb_dev_free(d_cnt);
  dfb_dev_free(d_scratch);
  if (st != DFB_OK) {
    if (cm) dfb_bsr_destroy(cm);
    if (pat) dfb_pattern_destroy(pat);
    return st;
  }
  *out_pattern = pat;
  *out = cm;
  return DFB_OK;
}
