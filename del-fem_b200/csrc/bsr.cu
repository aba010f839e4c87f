// bsr.cu — sparse_square::Matrix<[T;NN]> on the device (del-fem-cpu/src/sparse_square.rs:75-214, :349-447;
// del-fem-ls/src/sparse_square.rs:6-164): storage, set_zero, BC (set_fixed_bc :3-55) and mult_vec (:419-447).
//
// SpMV design (HBM-bound: 76 B per off-diagonal 3x3 block, 0.24 flop/B — no tensor cores):
//   variant 0   one thread per block-row, direct global loads (simple; reference point and fallback)
//   variant 4   4 lanes per row, 8-row tiles, two-stage TMA ring per warp over the canonical arrays (no extra memory)
//   variant 20  [default] the same compute phase over a tile-packed mirror of the matrix: ONE TMA bulk copy per tile
// Persistent warps (grid = #SM) walk the tiles with a fixed stride, so the optional fused dot product (p . Ap for CG) is
// reduced over a fixed grid in a fixed order: bit-reproducible.  (The superseded experiments — warp-tile cp.async / TMA staging,
// 2/8 lanes per row, 3-stage rings, row-pointer copies and prefetch — are recorded in profiles/README.md, not compiled.)
#include "dfb_internal.h"
#include "reduce.cuh"

// ------------------------------------------------------------------------------------------------ create / layouts / IO
static inline bool bsr_has_dia(const dfb_bsr *m) { return m->pat->convention == 0; }

DFB_API dfb_status dfb_bsr_create(dfb_ctx *ctx, dfb_pattern *pattern, int32_t blk_dim, dfb_bsr **out) {
  DFB_REQUIRE(ctx && pattern && out, "dfb_bsr_create: NULL argument");
  DFB_REQUIRE(blk_dim >= 1 && blk_dim <= 4, "dfb_bsr_create: blk_dim %d not in 1..4", blk_dim);
  DFB_REQUIRE(pattern->ctx == ctx, "dfb_bsr_create: pattern belongs to another context");
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_bsr *m = new dfb_bsr();
  m->ctx = ctx;
  m->pat = pattern;
  pattern->refcount += 1;
  m->blk_dim = blk_dim;
  m->int_begin = 0;
  m->int_end = pattern->num_blk;
  // no values are allocated yet: the matrix is "all zero" until something writes it (see dfb_internal.h)
  *out = m;
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_destroy(dfb_bsr *m) {
  if (!m) return DFB_OK;
  cudaStreamSynchronize(m->ctx->stream);
  dfb_dev_free(m->d_idx2val);
  dfb_dev_free(m->d_row2val);
  dfb_dev_free(m->d_flag);
  dfb_dev_free(m->d_packed);
  dfb_dev_free(m->d_tile_off);
  dfb_pattern_destroy(m->pat);
  delete m;
  return DFB_OK;
}

DFB_API dfb_pattern *dfb_bsr_pattern(const dfb_bsr *m) { return m ? m->pat : nullptr; }

// ---- the tile-packed layout: per tile of 8 block-rows one contiguous 16-byte-aligned record
//     [ u32 rowptr_rel[9] + pad : 48 B | u32 cols[nb] (pad 16) | diag 8 x NN f64 | vals nb x NN f64 (pad 16) ]
// so the SpMV fetches a tile with a single cp.async.bulk and everything its compute phase needs arrives with it.
__host__ __device__ inline uint32_t pk_tile_units(uint32_t nb, int NN, int has_dia, int rpt) {
  const uint32_t bytes = (uint32_t)dfb_pk_hdr_bytes(rpt) + ((nb * 4u + 15u) & ~15u) + (has_dia ? (uint32_t)(rpt * NN * 8) : 0u) + ((nb * NN * 8u + 15u) & ~15u);
  return bytes >> 4;
}

__global__ void k_pack_sizes(const uint32_t *__restrict__ row2idx, uint64_t n_rows, uint64_t n_tiles, int NN, int has_dia, int rpt,
                             uint32_t *__restrict__ tile_units, unsigned int *max_units) {
  unsigned int mx = 0;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r0 = t * rpt, r1 = min(r0 + (uint64_t)rpt, n_rows);
    const uint32_t u = pk_tile_units(row2idx[r1] - row2idx[r0], NN, has_dia, rpt);
    tile_units[t] = u;
    mx = max(mx, u);
  }
  for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0) atomicMax(max_units, mx);
}

// one warp per tile. STRUCT: header + column indices (pattern only, written once); VALUES: pack the canonical arrays into the
// record; UNPACK: the reverse
enum { PK_STRUCT = 1, PK_VALUES = 2, PK_UNPACK = 4 };
template <int MODE>
__global__ void __launch_bounds__(256) k_pack_tiles(const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col,
                                                    double *__restrict__ idx2val, double *__restrict__ row2val, uint64_t n_rows,
                                                    uint64_t n_tiles, int NN, int has_dia, int rpt, const uint32_t *__restrict__ tile_off,
                                                    unsigned char *__restrict__ packed) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t hdr_bytes = (uint32_t)dfb_pk_hdr_bytes(rpt);
  for (uint64_t t = warp; t < n_tiles; t += nwarp) {
    const uint64_t r0 = t * rpt, r1 = min(r0 + (uint64_t)rpt, n_rows);
    const uint32_t k0 = row2idx[r0], nb = row2idx[r1] - k0;
    unsigned char *rec = packed + (uint64_t)tile_off[t] * 16;
    const uint32_t ncol_pad = ((nb * 4u + 15u) & ~15u) / 4u;
    if (MODE & PK_STRUCT) {
      uint32_t *hdr = reinterpret_cast<uint32_t *>(rec);
      if (lane <= (unsigned)rpt) hdr[lane] = (r0 + lane <= r1) ? row2idx[r0 + lane] - k0 : nb;
      if (lane > (unsigned)rpt && lane < hdr_bytes / 4) hdr[lane] = 0u;
      uint32_t *cols = reinterpret_cast<uint32_t *>(rec + hdr_bytes);
      for (uint32_t k = lane; k < ncol_pad; k += 32) cols[k] = (k < nb) ? idx2col[k0 + k] : 0u;
    }
    double *dia = reinterpret_cast<double *>(rec + hdr_bytes + ncol_pad * 4u);
    double *vals = dia + (has_dia ? rpt * NN : 0);
    const uint32_t nv = nb * NN, nv_pad = (((nb * NN * 8u) + 15u) & ~15u) / 8u;
    const uint32_t nd = (uint32_t)(r1 - r0) * NN;
    if (MODE & PK_VALUES) {
      if (has_dia)
        for (uint32_t q = lane; q < (uint32_t)(rpt * NN); q += 32) dia[q] = (q < nd) ? row2val[r0 * NN + q] : 0.0;
      const double *src = idx2val + (uint64_t)k0 * NN;
      for (uint32_t q = lane; q < nv_pad; q += 32) vals[q] = (q < nv) ? src[q] : 0.0;
    }
    if (MODE & PK_UNPACK) {
      if (has_dia)
        for (uint32_t q = lane; q < nd; q += 32) row2val[r0 * NN + q] = dia[q];
      double *dst = idx2val + (uint64_t)k0 * NN;
      for (uint32_t q = lane; q < nv; q += 32) dst[q] = vals[q];
    }
  }
}

static uint32_t packed_buf_bytes(int N, int rpt);  // the SpMV's per-stage shared-memory buffer (defined with the kernel)

// allocate the packed records and write their pattern-only part (once per matrix); all value bytes start as zero
static dfb_status bsr_packed_structure(dfb_bsr *m) {
  if (m->packed_ok >= 0) return DFB_OK;
  dfb_ctx *c = m->ctx;
  const uint64_t n = m->pat->num_blk;
  const int NN = m->blk_dim * m->blk_dim, has_dia = bsr_has_dia(m);
  // sparse rows (spring networks: 6 blocks per row) get 16-row tiles: the SpMV is bound by bulk-copy operations per SM, not bytes,
  // when a tile is only ~4 KB (measured: 2048^2 cloth 3.96 TB/s with 8-row tiles)
  m->pk_rpt = (c->spmv_tile_rows == 8 || c->spmv_tile_rows == 16) ? (int)c->spmv_tile_rows : ((n > 0 && m->pat->nnz <= 8 * n) ? 16 : 8);
  const int rpt = m->pk_rpt;
  const uint64_t n_tiles = (n + rpt - 1) / rpt;
  DFB_TRY(dfb_dev_alloc_t(&m->d_tile_off, n_tiles + 2));
  unsigned int *d_max = (unsigned int *)(c->d_scratch + 18);
  DFB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned int), c->stream));
  DFB_LAUNCH(c, k_pack_sizes, c->sm_count * 8, 256, 0, m->pat->d_row2idx, n, n_tiles, NN, has_dia, rpt, m->d_tile_off, d_max);
  DFB_CHECK_LAUNCH();
  unsigned int h_max = 0;
  DFB_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  uint64_t total_units = 0;
  DFB_TRY(dfb_exclusive_scan_u32(c, m->d_tile_off, n_tiles, &total_units));  // synchronises: h_max is valid afterwards
  m->n_tiles = n_tiles;
  m->packed_cap = total_units * 16 + 256;
  DFB_TRY(dfb_dev_alloc((void **)&m->d_packed, m->packed_cap));
  DFB_CUDA(cudaMemsetAsync(m->d_packed, 0, m->packed_cap, c->stream));
  if (n_tiles > 0) {
    DFB_LAUNCH(c, k_pack_tiles<PK_STRUCT>, c->sm_count * 8, 256, 0, m->pat->d_row2idx, m->pat->d_idx2col, (double *)nullptr, (double *)nullptr, n,
               n_tiles, NN, has_dia, rpt, m->d_tile_off, m->d_packed);
    DFB_CHECK_LAUNCH();
  }
  m->packed_ok = ((uint64_t)h_max * 16 <= packed_buf_bytes(m->blk_dim, rpt)) ? 1 : 0;
  return DFB_OK;
}

static dfb_status bsr_canon_alloc(dfb_bsr *m, bool zero) {
  const uint64_t nn = (uint64_t)m->blk_dim * m->blk_dim;
  const bool fresh = !m->d_idx2val;
  if (!m->d_idx2val) DFB_TRY(dfb_dev_alloc_t(&m->d_idx2val, m->pat->nnz * nn + 64));
  if (bsr_has_dia(m) && !m->d_row2val) DFB_TRY(dfb_dev_alloc_t(&m->d_row2val, m->pat->num_blk * nn + 64));
  if (zero || fresh) {
    DFB_CUDA(cudaMemsetAsync(m->d_idx2val, 0, (m->pat->nnz * nn + 64) * sizeof(double), m->ctx->stream));
    if (m->d_row2val) DFB_CUDA(cudaMemsetAsync(m->d_row2val, 0, (m->pat->num_blk * nn + 64) * sizeof(double), m->ctx->stream));
  }
  return DFB_OK;
}

bool dfb_bsr_packed_native(const dfb_bsr *m) {
  if (m->ctx->spmv_variant != 20) return false;
  dfb_bsr *mm = const_cast<dfb_bsr *>(m);
  if (mm->packed_ok < 0 && bsr_packed_structure(mm) != DFB_OK) return false;
  return mm->packed_ok == 1;
}

PackedView dfb_bsr_packed_view(const dfb_bsr *m) {
  PackedView v;
  v.base = m->d_packed;
  v.tile_off = m->d_tile_off;
  v.NN = m->blk_dim * m->blk_dim;
  v.has_dia = bsr_has_dia(m) ? 1 : 0;
  v.rpt = m->pk_rpt;
  v.hdr = dfb_pk_hdr_bytes(m->pk_rpt);
  return v;
}

dfb_status dfb_bsr_canon_read(const dfb_bsr *cm) {
  dfb_bsr *m = const_cast<dfb_bsr *>(cm);
  if (m->canon_valid) return DFB_OK;
  dfb_ctx *c = m->ctx;
  if (!m->packed_valid) {  // all-zero matrix
    DFB_TRY(bsr_canon_alloc(m, true));
  } else {
    DFB_TRY(bsr_canon_alloc(m, false));
    if (m->n_tiles > 0) {
      DFB_LAUNCH(c, k_pack_tiles<PK_UNPACK>, c->sm_count * 8, 256, 0, m->pat->d_row2idx, m->pat->d_idx2col, m->d_idx2val, m->d_row2val,
                 m->pat->num_blk, m->n_tiles, m->blk_dim * m->blk_dim, bsr_has_dia(m) ? 1 : 0, m->pk_rpt, m->d_tile_off, m->d_packed);
      DFB_CHECK_LAUNCH();
    }
  }
  m->canon_valid = 1;
  return DFB_OK;
}

dfb_status dfb_bsr_canon_write(dfb_bsr *m) {
  DFB_TRY(dfb_bsr_canon_read(m));
  m->packed_valid = 0;
  return DFB_OK;
}

dfb_status dfb_bsr_canon_overwrite(dfb_bsr *m) {
  DFB_TRY(bsr_canon_alloc(m, false));
  m->canon_valid = 1;
  m->packed_valid = 0;
  return DFB_OK;
}

dfb_status dfb_bsr_packed_overwrite(dfb_bsr *m) {
  DFB_TRY(bsr_packed_structure(m));
  m->packed_valid = 1;
  m->canon_valid = 0;
  return DFB_OK;
}

// packed records hold the current values (packing the canonical arrays / zero-filling on demand)
dfb_status dfb_bsr_ensure_packed(const dfb_bsr *cm) {
  dfb_bsr *m = const_cast<dfb_bsr *>(cm);
  if (m->packed_valid) return DFB_OK;
  dfb_ctx *c = m->ctx;
  DFB_TRY(bsr_packed_structure(m));
  if (!m->canon_valid) {
    // all-zero matrix: clear the value part of every record (structure is kept)
    DFB_TRY(bsr_canon_alloc(m, true));
    m->canon_valid = 1;
  }
  if (m->n_tiles > 0) {
    dfb_phase_begin(c, DFB_PH_PACK);
    DFB_LAUNCH(c, k_pack_tiles<PK_VALUES>, c->sm_count * 8, 256, 0, m->pat->d_row2idx, m->pat->d_idx2col, m->d_idx2val, m->d_row2val,
               m->pat->num_blk, m->n_tiles, m->blk_dim * m->blk_dim, bsr_has_dia(m) ? 1 : 0, m->pk_rpt, m->d_tile_off, m->d_packed);
    dfb_phase_end(c);
    DFB_CHECK_LAUNCH();
  }
  m->packed_valid = 1;
  return DFB_OK;
}

dfb_status dfb_bsr_packed_modify(dfb_bsr *m) {
  DFB_TRY(dfb_bsr_ensure_packed(m));
  m->canon_valid = 0;
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_pack(dfb_bsr *m) {
  DFB_REQUIRE(m, "dfb_bsr_pack: NULL argument");
  if (m->ctx->spmv_variant != 20) return DFB_OK;
  return dfb_bsr_ensure_packed(m);
}

DFB_API dfb_status dfb_bsr_set_zero(dfb_bsr *m) {
  DFB_REQUIRE(m, "dfb_bsr_set_zero: NULL argument");
  // lazily: whoever needs the values next materialises the zeros in the layout it works on
  m->packed_valid = 0;
  m->canon_valid = 0;
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_upload(dfb_bsr *m, const double *row2val, const double *idx2val) {
  DFB_REQUIRE(m, "dfb_bsr_upload: NULL argument");
  if (row2val) DFB_REQUIRE(bsr_has_dia(m), "dfb_bsr_upload: matrix has no separate diagonal (convention 1)");
  if (idx2val && (row2val || !bsr_has_dia(m))) DFB_TRY(dfb_bsr_canon_overwrite(m));
  else DFB_TRY(dfb_bsr_canon_write(m));  // partial upload: the other array keeps its values
  const uint64_t nn = (uint64_t)m->blk_dim * m->blk_dim;
  if (idx2val) DFB_CUDA(cudaMemcpyAsync(m->d_idx2val, idx2val, m->pat->nnz * nn * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  if (row2val) DFB_CUDA(cudaMemcpyAsync(m->d_row2val, row2val, m->pat->num_blk * nn * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  DFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_download(const dfb_bsr *m, double *row2val, double *idx2val) {
  DFB_REQUIRE(m, "dfb_bsr_download: NULL argument");
  if (row2val) DFB_REQUIRE(bsr_has_dia(m), "dfb_bsr_download: matrix has no separate diagonal (convention 1)");
  DFB_TRY(dfb_bsr_canon_read(m));
  const uint64_t nn = (uint64_t)m->blk_dim * m->blk_dim;
  if (idx2val) DFB_CUDA(cudaMemcpyAsync(idx2val, m->d_idx2val, m->pat->nnz * nn * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  if (row2val) DFB_CUDA(cudaMemcpyAsync(row2val, m->d_row2val, m->pat->num_blk * nn * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  DFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ BC
// set_fixed_bc (sparse_square.rs:3-55). The reference's three passes touch disjoint data per block, so they collapse
// into one row-parallel pass: thread per block-row handles its diagonal block, then every off-diagonal block of the
// row (row zeroing by its own flags, column zeroing by the column block's flags). Only flagged entries are written.
template <int N>
__global__ void k_set_fixed_bc(double val_dia, const int32_t *__restrict__ flag, double *__restrict__ row2val,
                               double *__restrict__ idx2val, const uint32_t *__restrict__ row2idx,
                               const uint32_t *__restrict__ idx2col, uint64_t num_blk, int convention) {
  constexpr int NN = N * N;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    int fi[N];
    bool any_i = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      fi[d] = flag[i * N + d];
      any_i |= (fi[d] != 0);
    }
    if (convention == 0 && any_i) {
      double *D = row2val + i * NN;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        if (fi[d] == 0) continue;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          D[d + N * j] = 0.0;
          D[j + N * d] = 0.0;
        }
        D[d + N * d] = val_dia;
      }
    }
    for (uint32_t k = row2idx[i]; k < row2idx[i + 1]; ++k) {
      const uint32_t c = idx2col[k];
      double *V = idx2val + (uint64_t)k * NN;
      if (convention == 1 && c == (uint32_t)i) {
        // diagonal block stored inside the pattern (blkcsr): same treatment as the separate diagonal
#pragma unroll
        for (int d = 0; d < N; ++d) {
          if (fi[d] == 0) continue;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            V[d + N * j] = 0.0;
            V[j + N * d] = 0.0;
          }
          V[d + N * d] = val_dia;
        }
        continue;
      }
#pragma unroll
      for (int d = 0; d < N; ++d) {
        if (fi[d] != 0) {
#pragma unroll
          for (int j = 0; j < N; ++j) V[d + N * j] = 0.0;
        }
        if (flag[(uint64_t)c * N + d] != 0) {
#pragma unroll
          for (int a = 0; a < N; ++a) V[a + N * d] = 0.0;
        }
      }
    }
  }
}

// the same pass over the packed records (native layout of spmv_variant 20): thread per block-row
template <int N>
__global__ void k_set_fixed_bc_packed(double val_dia, const int32_t *__restrict__ flag, const PackedView pv, uint64_t num_blk) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x) {
    int fi[N];
    bool any_i = false;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      fi[d] = flag[i * N + d];
      any_i |= (fi[d] != 0);
    }
    if (pv.has_dia && any_i) {
      double *D = pk_diag(pv, i);
#pragma unroll
      for (int d = 0; d < N; ++d) {
        if (fi[d] == 0) continue;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          D[d + N * j] = 0.0;
          D[j + N * d] = 0.0;
        }
        D[d + N * d] = val_dia;
      }
    }
    const uint32_t *cols;
    uint32_t n;
    double *vals = pk_row(pv, i, &cols, &n);
    for (uint32_t k = 0; k < n; ++k) {
      const uint32_t c = cols[k];
      double *V = vals + (uint64_t)k * (N * N);
      if (!pv.has_dia && c == (uint32_t)i) {
#pragma unroll
        for (int d = 0; d < N; ++d) {
          if (fi[d] == 0) continue;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            V[d + N * j] = 0.0;
            V[j + N * d] = 0.0;
          }
          V[d + N * d] = val_dia;
        }
        continue;
      }
#pragma unroll
      for (int d = 0; d < N; ++d) {
        if (fi[d] != 0) {
#pragma unroll
          for (int j = 0; j < N; ++j) V[d + N * j] = 0.0;
        }
        if (flag[(uint64_t)c * N + d] != 0) {
#pragma unroll
          for (int a = 0; a < N; ++a) V[a + N * d] = 0.0;
        }
      }
    }
  }
}

static dfb_status launch_set_fixed_bc(dfb_bsr *m, double val_dia, const int32_t *d_flag) {
  dfb_ctx *c = m->ctx;
  const int grid = c->sm_count * 8;
  if ((m->packed_valid || !m->canon_valid) && dfb_bsr_packed_native(m)) {
    DFB_TRY(dfb_bsr_packed_modify(m));
    const PackedView pv = dfb_bsr_packed_view(m);
    dfb_phase_begin(c, DFB_PH_BC);
    switch (m->blk_dim) {
      case 1: DFB_LAUNCH(c, k_set_fixed_bc_packed<1>, grid, 128, 0, val_dia, d_flag, pv, m->pat->num_blk); break;
      case 2: DFB_LAUNCH(c, k_set_fixed_bc_packed<2>, grid, 128, 0, val_dia, d_flag, pv, m->pat->num_blk); break;
      case 3: DFB_LAUNCH(c, k_set_fixed_bc_packed<3>, grid, 128, 0, val_dia, d_flag, pv, m->pat->num_blk); break;
      case 4: DFB_LAUNCH(c, k_set_fixed_bc_packed<4>, grid, 128, 0, val_dia, d_flag, pv, m->pat->num_blk); break;
    }
    dfb_phase_end(c);
    DFB_CHECK_LAUNCH();
    return DFB_OK;
  }
  DFB_TRY(dfb_bsr_canon_write(m));
  dfb_phase_begin(c, DFB_PH_BC);
  switch (m->blk_dim) {
    case 1: DFB_LAUNCH(c, k_set_fixed_bc<1>, grid, 128, 0, val_dia, d_flag, m->d_row2val, m->d_idx2val, m->pat->d_row2idx, m->pat->d_idx2col, m->pat->num_blk, m->pat->convention); break;
    case 2: DFB_LAUNCH(c, k_set_fixed_bc<2>, grid, 128, 0, val_dia, d_flag, m->d_row2val, m->d_idx2val, m->pat->d_row2idx, m->pat->d_idx2col, m->pat->num_blk, m->pat->convention); break;
    case 3: DFB_LAUNCH(c, k_set_fixed_bc<3>, grid, 128, 0, val_dia, d_flag, m->d_row2val, m->d_idx2val, m->pat->d_row2idx, m->pat->d_idx2col, m->pat->num_blk, m->pat->convention); break;
    case 4: DFB_LAUNCH(c, k_set_fixed_bc<4>, grid, 128, 0, val_dia, d_flag, m->d_row2val, m->d_idx2val, m->pat->d_row2idx, m->pat->d_idx2col, m->pat->num_blk, m->pat->convention); break;
  }
  dfb_phase_end(c);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

static uint64_t bsr_n_all(const dfb_bsr *m) {
  return m->pat->num_blk + (m->halo ? (m->halo->recv_ptr.empty() ? 0 : m->halo->recv_ptr.back()) : 0);
}

DFB_API dfb_status dfb_bsr_set_fixed_dof(dfb_bsr *m, double val_dia, const int32_t *flag_host) {
  DFB_REQUIRE(m && flag_host, "dfb_bsr_set_fixed_dof: NULL argument");
  dfb_ctx *c = m->ctx;
  const uint64_t need = bsr_n_all(m) * m->blk_dim;
  if (m->flag_cap < need) {
    dfb_dev_free(m->d_flag);
    m->d_flag = nullptr;
    DFB_TRY(dfb_dev_alloc_t(&m->d_flag, need));
    m->flag_cap = need;
  }
  DFB_CUDA(cudaMemcpyAsync(m->d_flag, flag_host, need * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  DFB_TRY(launch_set_fixed_bc(m, val_dia, m->d_flag));
  DFB_CUDA(cudaStreamSynchronize(c->stream));  // flag_host is borrowed only for the duration of the call
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_set_fixed_dof_dev(dfb_bsr *m, double val_dia, const dfb_bcflag *f) {
  DFB_REQUIRE(m && f, "dfb_bsr_set_fixed_dof_dev: NULL argument");
  DFB_REQUIRE(f->blk_dim == m->blk_dim && f->n_blk_all >= bsr_n_all(m), "dfb_bsr_set_fixed_dof_dev: assert_eq!(bc_flag.len(), row2val.len())");
  return launch_set_fixed_bc(m, val_dia, f->d);
}

// row2val[i][d + N d] += scale * v[i][d]
__global__ void k_add_to_diagonal(double *__restrict__ row2val, int N, double scale, const double *__restrict__ v, uint64_t n_dof) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_dof; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = q / N;
    const int d = (int)(q - i * N);
    row2val[i * N * N + d + N * d] += scale * v[q];
  }
}
__global__ void k_add_to_diagonal_packed(const PackedView pv, int N, double scale, const double *__restrict__ v, uint64_t n_dof) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_dof; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = q / N;
    const int d = (int)(q - i * N);
    pk_diag(pv, i)[d + N * d] += scale * v[q];
  }
}

DFB_API dfb_status dfb_bsr_add_to_diagonal(dfb_bsr *m, double scale, const dfb_vec *v) {
  DFB_REQUIRE(m && v, "dfb_bsr_add_to_diagonal: NULL argument");
  DFB_REQUIRE(bsr_has_dia(m), "dfb_bsr_add_to_diagonal: convention 1 matrices have no separate diagonal");
  DFB_REQUIRE(v->n_blk == m->pat->num_blk && v->blk_dim == m->blk_dim, "dfb_bsr_add_to_diagonal: shape mismatch");
  if ((m->packed_valid || !m->canon_valid) && dfb_bsr_packed_native(m)) {
    DFB_TRY(dfb_bsr_packed_modify(m));
    DFB_LAUNCH(m->ctx, k_add_to_diagonal_packed, m->ctx->sm_count * 8, 256, 0, dfb_bsr_packed_view(m), m->blk_dim, scale, v->d, v->len());
  } else {
    DFB_TRY(dfb_bsr_canon_write(m));
    DFB_LAUNCH(m->ctx, k_add_to_diagonal, m->ctx->sm_count * 8, 256, 0, m->d_row2val, m->blk_dim, scale, v->d, v->len());
  }
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// values of m += alpha * values of other (same pattern object): sums matrices assembled from different element families
__global__ void k_axpy_values(double *__restrict__ y, double alpha, const double *__restrict__ x, uint64_t n) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) y[q] += alpha * x[q];
}
// packed form: the two matrices share the pattern, hence the record layout; one warp per tile walks the value part of the record
__global__ void __launch_bounds__(256) k_axpy_packed(const PackedView y, const PackedView x, double alpha, uint64_t n_tiles) {
  const unsigned lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = warp; t < n_tiles; t += nwarp) {
    const uint32_t nb = reinterpret_cast<const uint32_t *>(pk_record(y, t))[y.rpt];
    const uint32_t n = ((y.has_dia ? y.rpt : 0) + nb) * y.NN;
    double *yv = pk_tile_values(y, t);
    const double *xv = pk_tile_values(x, t);
    for (uint32_t q = lane; q < n; q += 32) yv[q] += alpha * xv[q];
  }
}

DFB_API dfb_status dfb_bsr_add_scaled(dfb_bsr *m, double alpha, const dfb_bsr *other) {
  DFB_REQUIRE(m && other, "dfb_bsr_add_scaled: NULL argument");
  DFB_REQUIRE(m->pat == other->pat && m->blk_dim == other->blk_dim, "dfb_bsr_add_scaled: matrices must share pattern and block size");
  const uint64_t NN = (uint64_t)m->blk_dim * m->blk_dim;
  if (other->packed_valid && (m->packed_valid || !m->canon_valid) && dfb_bsr_packed_native(m)) {
    DFB_TRY(dfb_bsr_packed_modify(m));
    if (m->n_tiles) DFB_LAUNCH(m->ctx, k_axpy_packed, m->ctx->sm_count * 8, 256, 0, dfb_bsr_packed_view(m), dfb_bsr_packed_view(other), alpha, m->n_tiles);
    DFB_CHECK_LAUNCH();
    return DFB_OK;
  }
  DFB_TRY(dfb_bsr_canon_write(m));
  DFB_TRY(dfb_bsr_canon_read(other));
  const uint64_t n_off = m->pat->nnz * NN;
  if (n_off) DFB_LAUNCH(m->ctx, k_axpy_values, m->ctx->sm_count * 8, 256, 0, m->d_idx2val, alpha, other->d_idx2val, n_off);
  if (bsr_has_dia(m)) DFB_LAUNCH(m->ctx, k_axpy_values, m->ctx->sm_count * 8, 256, 0, m->d_row2val, alpha, other->d_row2val, m->pat->num_blk * NN);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// single-element merge (API-parity shim; merge_for_array_blk / ls merge / csrdia / blkcsr). One thread.
__global__ void k_merge_one(const double *__restrict__ emat, const uint32_t *__restrict__ node2vtx, int n_node, int NN,
                            int flavour, int convention, const uint32_t *__restrict__ row2idx,
                            const uint32_t *__restrict__ idx2col, uint64_t num_blk, double *row2val, double *idx2val, int *err) {
  for (int in = 0; in < n_node; ++in) {
    const uint32_t r = node2vtx[in];
    if (r >= num_blk) {
      *err = 1;
      return;
    }
    for (int jn = 0; jn < n_node; ++jn) {
      const uint32_t cvt = node2vtx[jn];
      const bool is_dia = convention == 0 && (flavour == 0 ? (in == jn) : (r == cvt));
      double *dst = nullptr;
      if (is_dia) dst = row2val + (uint64_t)r * NN;
      else {
        // last occurrence wins, like the reference's col2idx scatter
        for (uint32_t k = row2idx[r]; k < row2idx[r + 1]; ++k)
          if (idx2col[k] == cvt) dst = idx2val + (uint64_t)k * NN;
        if (!dst) {
          *err = 1;
          return;
        }
      }
      const double *e = emat + (in * n_node + jn) * NN;
      for (int q = 0; q < NN; ++q) dst[q] += e[q];
    }
  }
}

DFB_API dfb_status dfb_bsr_merge_one(dfb_bsr *m, const double *emat, const uint64_t *node2vtx, int32_t n_node, int32_t flavour) {
  DFB_REQUIRE(m && emat && node2vtx, "dfb_bsr_merge_one: NULL argument");
  DFB_REQUIRE(n_node >= 1 && n_node <= 8, "dfb_bsr_merge_one: n_node %d not in 1..8", n_node);
  DFB_TRY(dfb_bsr_canon_write(m));
  dfb_ctx *c = m->ctx;
  const int NN = m->blk_dim * m->blk_dim;
  double *d_e = nullptr;
  uint32_t *d_n = nullptr;
  uint32_t h_n[8];
  for (int i = 0; i < n_node; ++i) {
    if (node2vtx[i] >= 0xFFFFFFFFull) return dfb_fail(DFB_ERR_INDEX_OVERFLOW, "dfb_bsr_merge_one: vertex index overflow");
    h_n[i] = (uint32_t)node2vtx[i];
  }
  DFB_TRY(dfb_dev_alloc_t(&d_e, (size_t)n_node * n_node * NN));
  DFB_TRY(dfb_dev_alloc_t(&d_n, 8));
  DFB_CUDA(cudaMemcpyAsync(d_e, emat, (size_t)n_node * n_node * NN * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  DFB_CUDA(cudaMemcpyAsync(d_n, h_n, n_node * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  DFB_LAUNCH(c, k_merge_one, 1, 1, 0, d_e, d_n, n_node, NN, flavour, m->pat->convention, m->pat->d_row2idx, m->pat->d_idx2col,
             m->pat->num_blk, m->d_row2val, m->d_idx2val, c->d_errflag + 1);
  DFB_CHECK_LAUNCH();
  dfb_status st = dfb_check_error_flag(c, 1, DFB_ERR_PATTERN_MISS, "dfb_bsr_merge_one: (row, col) of the element is not in the pattern");
  dfb_dev_free(d_e);
  dfb_dev_free(d_n);
  return st;
}

// ------------------------------------------------------------------------------------------------ SpMV
struct SpmvArgs {
  const uint32_t *row2idx;
  const uint32_t *idx2col;
  const double *idx2val;
  const double *row2val;  // nullptr for convention 1
  const double *x;
  double *y;
  double alpha, beta;
  uint32_t row_begin, row_end;
  int fuse_dot;        // 0 none, 1 red[0] = sum x.y, 2 red[0] += sum x.y
  DfbSolverState *st;  // nullptr outside the solvers; inside, the launch reads iteration / stop flag / exchange ids from it
  double *partials;
  unsigned int *ticket;
  DfbPeerView pv;      // pv.post (and nranks > 1): the launch that completes p.Ap posts it to every peer's mailbox
  DfbHaloWait hw;      // hw.xg[0] != nullptr: ghost columns come from the peer landing zone (variant 20 inside the solvers)
};

template <int N>
__device__ __forceinline__ void block_times_vec(double (&s)[N], const double *__restrict__ v, const double (&xv)[N]) {
#pragma unroll
  for (int b = 0; b < N; ++b)
#pragma unroll
    for (int a = 0; a < N; ++a) s[a] += v[a + N * b] * xv[b];
}

template <int N>
__device__ __forceinline__ void load_x(double (&xv)[N], const double *__restrict__ x, uint32_t col) {
  const double *p = x + (uint64_t)col * N;
#pragma unroll
  for (int a = 0; a < N; ++a) xv[a] = __ldg(p + a);
}

template <int N>
__device__ __forceinline__ void load_x_coherent(double (&xv)[N], const double *x, uint32_t col) {
  const double *p = x + (uint64_t)col * N;
#pragma unroll
  for (int a = 0; a < N; ++a) xv[a] = __ldcg(p + a);
}

// off-diagonal part of one row; vals/cols may point to shared or global memory
template <int N>
__device__ __forceinline__ void row_offdiag(double (&s)[N], const double *vals, const uint32_t *cols, uint32_t kb, uint32_t ke,
                                            const double *__restrict__ x) {
  constexpr int NN = N * N;
#pragma unroll 2
  for (uint32_t k = kb; k < ke; ++k) {
    double xv[N];
    load_x<N>(xv, x, cols[k]);
    block_times_vec<N>(s, vals + (uint64_t)k * NN, xv);
  }
}

template <int N>
__device__ __forceinline__ double row_finish(const SpmvArgs &a, uint32_t r, double (&s)[N], const double *diag_blk) {
  double xo[N];
  load_x<N>(xo, a.x, r);
  if (diag_blk) block_times_vec<N>(s, diag_blk, xo);
  double d = 0.0;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    double yv = a.alpha * s[q];
    if (a.beta != 0.0) yv += a.beta * a.y[(uint64_t)r * N + q];
    a.y[(uint64_t)r * N + q] = yv;
    d += xo[q] * yv;
  }
  return d;
}

__device__ __forceinline__ void spmv_publish_dot(const SpmvArgs &a, double d, double *s_buf) {
  double acc[1] = {d};
  if (grid_sum_last_block<1>(acc, a.partials, a.ticket, s_buf)) {  // true for every thread of the last block
    if (threadIdx.x == 0) {
      if (a.fuse_dot == 2) a.st->red[0] += acc[0];
      else a.st->red[0] = acc[0];
      s_buf[0] = a.st->red[0];
    }
    if (a.pv.nranks > 1 && a.pv.post) peer_post_block<1>(a.pv, dfb_xid_pq(a.st->xbase, a.st->iter), s_buf);
  }
}

// ---- variant 0: thread per row
template <int N>
__global__ void __launch_bounds__(128) k_spmv_row(const SpmvArgs a) {
  __shared__ double s_buf[32];
  if (a.st && a.st->done[a.st->iter & 1]) return;
  double dot = 0.0;
  for (uint64_t r = (uint64_t)a.row_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.row_end;
       r += (uint64_t)gridDim.x * blockDim.x) {
    double s[N];
#pragma unroll
    for (int q = 0; q < N; ++q) s[q] = 0.0;
    row_offdiag<N>(s, a.idx2val, a.idx2col, a.row2idx[r], a.row2idx[r + 1], a.x);
    dot += row_finish<N>(a, (uint32_t)r, s, a.row2val ? a.row2val + r * N * N : nullptr);
  }
  if (a.fuse_dot) spmv_publish_dot(a, dot, s_buf);
}

// ---- shared-memory / TMA helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// one-lane TMA bulk copy of the 16-byte-aligned superset of [src, src+bytes) (the caller guarantees 16 bytes of slack on both
// sides of the source); the data starts (uintptr_t)src & 15 bytes into the staged region
__device__ __forceinline__ void stage_bulk(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar_smem,
                                           unsigned long long pol) {
  const uintptr_t s = (uintptr_t)src;
  const uint32_t mis = (uint32_t)(s & 15);
  const char *g = (const char *)(s - mis);
  const uint32_t total = (mis + bytes + 15u) & ~15u;
  if (total > 0) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
        "l"(g), "r"(total), "r"(bar_smem), "l"(pol)
        : "memory");
  }
}

__device__ __forceinline__ uint32_t aligned_total(const void *src, uint32_t bytes) {
  const uint32_t mis = (uint32_t)((uintptr_t)src & 15);
  return (mis + bytes + 15u) & ~15u;
}

// ---- variant 4: 4 lanes per row, tiles of 8 rows, two-stage TMA ring per warp over the CANONICAL arrays (no extra memory).
// While a warp computes tile i out of one buffer, the bulk copies of tile i+1 are already in flight into the other; the row's
// blocks are split over 4 lanes (one gather round instead of ~14 dependent ones) and summed with shuffles. Three bulk copies
// per tile (values, column indices, diagonal blocks); the row pointers are plain loads issued before the barrier wait.
template <int N>
struct RingSmem {
  static constexpr int NN = N * N;
  static constexpr int LPR = 4, NST = 2;
  static constexpr int RPT = 32 / LPR;   // rows per tile
  static constexpr int CAP = RPT * 16;   // blocks per tile held in shared memory (longer tiles take the direct path)
  static constexpr int VAL_BYTES = ((CAP * NN * 8 + 16 + 15) / 16) * 16;
  static constexpr int COL_BYTES = ((CAP * 4 + 16 + 15) / 16) * 16;
  static constexpr int DIA_BYTES = ((RPT * NN * 8 + 16 + 15) / 16) * 16;
  static constexpr int BUF_BYTES = VAL_BYTES + COL_BYTES + DIA_BYTES;
  static constexpr int BAR_BYTES = ((NST * 8 + 15) / 16) * 16;
  static constexpr int WARP_BYTES = NST * BUF_BYTES + BAR_BYTES;
  static constexpr int WARPS_RAW = (220 * 1024) / WARP_BYTES;
  static constexpr int WARPS = WARPS_RAW > 24 ? 24 : (WARPS_RAW < 1 ? 1 : WARPS_RAW);  // warps per CTA (one CTA per SM)
};

// one row (or its share) of the ring kernel; vals/cols/dia point to the staged tile (shared) or to global memory
template <int N, int LPR, bool HALO = false>
__device__ __forceinline__ double ring_row(const SpmvArgs &a, const double *vals, const uint32_t *cols, const double *dia,
                                           uint32_t kb, uint32_t ke, uint32_t r, bool valid, unsigned sub,
                                           const double *xg = nullptr) {
  constexpr int NN = N * N;
  double s[N];
#pragma unroll
  for (int q = 0; q < N; ++q) s[q] = 0.0;
  // this lane's share of the row: blocks kb+sub, kb+sub+LPR, ... four at a time (index loads, then gathers, then FMAs)
  for (uint32_t k = kb + sub; k < ke; k += 4 * LPR) {
    uint32_t c[4];
    double xv[4][N];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t kk = k + u * LPR;
      c[u] = (kk < ke) ? cols[kk] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c[u] != 0xFFFFFFFFu) {
        // ghost columns live in the landing zone the neighbours write WHILE this kernel runs: read them with L2-coherent loads
        // (ld.global.cg), never through the non-coherent path; owned x is read-only for the kernel's lifetime (__ldg)
        if (HALO && c[u] >= a.hw.n_own) load_x_coherent<N>(xv[u], xg, c[u]);
        else load_x<N>(xv[u], a.x, c[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (c[u] != 0xFFFFFFFFu) block_times_vec<N>(s, vals + (uint64_t)(k + u * LPR) * NN, xv[u]);
    }
  }
  // diagonal block: lane `sub` contributes column `sub` (, sub+LPR, ...)
  if (valid && dia) {
    for (int cc = sub; cc < N; cc += LPR) {
      const double xo = __ldg(a.x + (uint64_t)r * N + cc);
#pragma unroll
      for (int q = 0; q < N; ++q) s[q] += dia[q + N * cc] * xo;
    }
  }
#pragma unroll
  for (int off = 1; off < LPR; off <<= 1) {
#pragma unroll
    for (int q = 0; q < N; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], off);
  }
  double d = 0.0;
  if (valid && sub == 0) {
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double yv = a.alpha * s[q];
      if (a.beta != 0.0) yv += a.beta * a.y[(uint64_t)r * N + q];
      a.y[(uint64_t)r * N + q] = yv;
      d += __ldg(a.x + (uint64_t)r * N + q) * yv;
    }
  }
  return d;
}

template <int N>
__global__ void __launch_bounds__(RingSmem<N>::WARPS * 32) k_spmv_ring(const SpmvArgs a) {
  using S = RingSmem<N>;
  constexpr int NN = N * N, RPT = S::RPT, CAP = S::CAP, LPR = S::LPR, NST = S::NST;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_buf[32];
  if (a.st && a.st->done[a.st->iter & 1]) return;

  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned sub = lane % LPR, rl = lane / LPR;
  unsigned char *wbase = smem_raw + (size_t)wid * S::WARP_BYTES;
  const uint32_t bar0 = smem_u32(wbase + NST * S::BUF_BYTES);
  const unsigned long long pol = policy_evict_first();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NST; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * j) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint32_t n_rows = a.row_end - a.row_begin;
  const uint32_t n_tiles = (n_rows + RPT - 1) / RPT;
  const uint32_t wg = blockIdx.x * nw + wid, ws = gridDim.x * nw;

  // lane 0 issues the three bulk copies of tile t into buffer b (nothing for oversized tiles)
  auto issue = [&](uint32_t t, uint32_t b) {
    if (lane != 0 || t >= n_tiles) return;
    const uint32_t r0 = a.row_begin + t * RPT;
    const uint32_t r1 = min(r0 + (uint32_t)RPT, a.row_end);
    const uint32_t k0 = __ldg(a.row2idx + r0), nb = __ldg(a.row2idx + r1) - k0;
    if (nb > (uint32_t)CAP) return;
    unsigned char *buf = wbase + (size_t)b * S::BUF_BYTES;
    const uint32_t bar = bar0 + 8 * b;
    const double *g_val = a.idx2val + (uint64_t)k0 * NN;
    const uint32_t *g_col = a.idx2col + k0;
    const double *g_dia = a.row2val ? a.row2val + (uint64_t)r0 * NN : nullptr;
    const uint32_t vb = nb * NN * 8u, cb = nb * 4u, db = (r1 - r0) * NN * 8u;
    const uint32_t tx = aligned_total(g_val, vb) + aligned_total(g_col, cb) + (g_dia ? aligned_total(g_dia, db) : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
    stage_bulk(smem_u32(buf), g_val, vb, bar, pol);
    stage_bulk(smem_u32(buf + S::VAL_BYTES), g_col, cb, bar, pol);
    if (g_dia) stage_bulk(smem_u32(buf + S::VAL_BYTES + S::COL_BYTES), g_dia, db, bar, pol);
  };

#pragma unroll
  for (int j = 0; j < NST; ++j) issue(wg + j * ws, j);
  uint32_t phase_bits = 0u;
  double dot = 0.0;
  uint32_t b = 0;
  for (uint32_t t = wg; t < n_tiles; t += ws) {
    const uint32_t r0 = a.row_begin + t * RPT;
    const uint32_t r1 = min(r0 + (uint32_t)RPT, a.row_end);
    const uint32_t k0 = __ldg(a.row2idx + r0), k1 = __ldg(a.row2idx + r1);
    const uint32_t nb = k1 - k0;
    const uint32_t r = r0 + rl;
    const bool valid = r < r1;
    uint32_t kb = k0, ke = k0;
    if (valid) {  // requested before the barrier wait so that their latency overlaps it
      kb = a.row2idx[r];
      ke = a.row2idx[r + 1];
    }
    if (nb <= (uint32_t)CAP) {
      unsigned char *buf = wbase + (size_t)b * S::BUF_BYTES;
      const uint32_t bar = bar0 + 8 * b;
      const uint32_t ph = (phase_bits >> b) & 1u;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "RING_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra RING_DONE;\n"
          "bra RING_WAIT;\n"
          "RING_DONE:\n"
          "}\n" ::"r"(bar),
          "r"(ph)
          : "memory");
      phase_bits ^= (1u << b);
      const double *g_val = a.idx2val + (uint64_t)k0 * NN;
      const uint32_t *g_col = a.idx2col + k0;
      const double *sv = reinterpret_cast<const double *>(buf + ((uintptr_t)g_val & 15));
      const uint32_t *sc = reinterpret_cast<const uint32_t *>(buf + S::VAL_BYTES + ((uintptr_t)g_col & 15));
      const double *sd = nullptr;
      if (a.row2val) {
        const double *g_dia = a.row2val + (uint64_t)r0 * NN;
        sd = reinterpret_cast<const double *>(buf + S::VAL_BYTES + S::COL_BYTES + ((uintptr_t)g_dia & 15)) + rl * NN;
      }
      dot += ring_row<N, LPR>(a, sv, sc, sd, kb - k0, ke - k0, r, valid, sub);
    } else {
      dot += ring_row<N, LPR>(a, a.idx2val, a.idx2col, a.row2val ? a.row2val + (uint64_t)r * NN : nullptr, kb, ke, r, valid, sub);
    }
    __syncwarp();  // every lane is done with buffer b
    issue(t + NST * ws, b);
    b = (b + 1 == NST) ? 0u : b + 1;
  }
  if (a.fuse_dot) spmv_publish_dot(a, dot, s_buf);
}

template <int N>
static dfb_status spmv_launch_ring(dfb_ctx *c, const SpmvArgs &a, cudaStream_t strm) {
  using S = RingSmem<N>;
  const int warps = S::WARPS;
  const size_t smem = (size_t)warps * S::WARP_BYTES;
  // the dynamic shared-memory opt-in is a per-device function attribute: set it on every launch (cheap) instead of caching a
  // process-global flag that a second context on another device would wrongly inherit
  DFB_CUDA(cudaFuncSetAttribute(k_spmv_ring<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint64_t n_rows = a.row_end - a.row_begin;
  const uint64_t n_tiles = (n_rows + S::RPT - 1) / S::RPT;
  uint64_t g = (n_tiles + warps - 1) / warps;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count) g = (uint64_t)c->sm_count;
  DFB_LAUNCH_ON(c, strm, (k_spmv_ring<N>), (unsigned)g, warps * 32, smem, a);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ---- variant 20: tile-packed mirror, ONE TMA bulk copy per 8-row tile.
// Measured on the ring kernels: every additional bulk-copy operation per tile costs ~10 % (the per-SM TMA unit is busy per
// operation, not only per byte), and row-pointer loads from DRAM sit on every warp's critical path. The packed mirror stores,
// per tile of 8 block-rows, one contiguous 16-byte-aligned record
//     [ u32 rowptr_rel[9] + pad : 48 B | u32 cols[nb] (pad 16) | diag 8 x NN f64 | vals nb x NN f64 (pad 16) ]
// so a tile is fetched by a single cp.async.bulk and everything the compute phase needs (relative row pointers included)
// arrives with it. With this variant the packed records are the matrix's native storage (dfb_internal.h): assembly, BC and the
// Jacobi preconditioners write / read them directly, the canonical arrays are only materialised for entry points that need them.
template <int N, int RPT>
struct PackedCfg {
  static constexpr int NN = N * N;
  static constexpr int LPR = 32 / RPT;   // lanes per row: 4 (8-row tiles) or 2 (16-row tiles)
  static constexpr int HDR = ((RPT + 1) * 4 + 15) / 16 * 16;
  static constexpr int CAP = 128;        // blocks per tile held in shared memory (longer tiles take the direct path)
  static constexpr int BUF_BYTES = HDR + ((CAP * 4 + 15) / 16) * 16 + RPT * NN * 8 + ((CAP * NN * 8 + 15) / 16) * 16;
  static constexpr int WARP_BYTES = 2 * BUF_BYTES + 16;
  static constexpr int WARPS_RAW = (220 * 1024) / WARP_BYTES;
  static constexpr int WARPS = WARPS_RAW > 24 ? 24 : (WARPS_RAW < 1 ? 1 : WARPS_RAW);
};

static uint32_t packed_buf_bytes(int N, int rpt) {
  if (rpt == 16) {
    switch (N) {
      case 1: return PackedCfg<1, 16>::BUF_BYTES;
      case 2: return PackedCfg<2, 16>::BUF_BYTES;
      case 3: return PackedCfg<3, 16>::BUF_BYTES;
      default: return PackedCfg<4, 16>::BUF_BYTES;
    }
  }
  switch (N) {
    case 1: return PackedCfg<1, 8>::BUF_BYTES;
    case 2: return PackedCfg<2, 8>::BUF_BYTES;
    case 3: return PackedCfg<3, 8>::BUF_BYTES;
    default: return PackedCfg<4, 8>::BUF_BYTES;
  }
}

// HALO = false is the single-domain kernel; HALO = true adds the peer-halo logic (ghost columns from the landing zone, interior
// tiles first, bounded wait before the first boundary tile; only inside the solvers: needs a.st). The two are separate
// instantiations because the tile mapping in the hot loop cost the single-domain kernel when it was compiled in unconditionally.
template <int N, bool HALO, int RPT>
__global__ void __launch_bounds__(PackedCfg<N, RPT>::WARPS * 32) k_spmv_packed(const SpmvArgs a, const unsigned char *__restrict__ packed,
                                                                              const uint32_t *__restrict__ tile_off, int has_dia) {
  using S = PackedCfg<N, RPT>;
  constexpr int NN = N * N, LPR = S::LPR, PK_HDR = S::HDR;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_buf[32];
  if (a.st && a.st->done[a.st->iter & 1]) return;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned sub = lane % LPR, rl = lane / LPR;
  unsigned char *wbase = smem_raw + (size_t)wid * S::WARP_BYTES;
  const uint32_t bar0 = smem_u32(wbase + 2 * S::BUF_BYTES);
  const unsigned long long pol = policy_evict_first();
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t t_begin = a.row_begin / RPT, t_end = (a.row_end + RPT - 1) / RPT;
  const uint32_t wg = blockIdx.x * nw + wid, ws = gridDim.x * nw;
  // Tiles are visited in a virtual order v = t_begin + wg, + ws, ...: with a peer halo the interior tiles come first and the
  // tiles that read ghost columns ([t_begin, t_lo) and [t_hi, t_end)) last, so the neighbours' pushes have normally landed by
  // the time a warp reaches them; without a halo the mapping is the identity.
  const uint32_t t_lo = HALO ? min(max(a.hw.tile_lo_end, t_begin), t_end) : t_begin;
  const uint32_t t_hi = HALO ? min(max(a.hw.tile_hi_begin, t_lo), t_end) : t_end;
  const uint32_t n_int = t_hi - t_lo;
  auto tile_of = [&](uint32_t v) -> uint32_t {   // v and the result are absolute tile indices in [t_begin, t_end)
    if (!HALO || v >= t_hi) return v;
    const uint32_t k = v - t_begin;
    return k < n_int ? t_lo + k : t_begin + (k - n_int);
  };
  bool waited = !HALO;
  // exchange id of the ghost values this launch multiplies, and the landing-zone parity buffer that holds them
  const unsigned long long hseq = HALO ? dfb_hid(a.st->hbase, a.st->iter) : 0ull;
  const double *xg = HALO ? a.hw.xg[hseq & 1ull] : nullptr;

  struct Off {
    uint32_t o0, o1;
  };
  auto offs = [&](uint32_t v) -> Off {
    Off o = {0u, 0u};
    if (v < t_end) {
      const uint32_t t = tile_of(v);
      o.o0 = __ldg(tile_off + t);
      o.o1 = __ldg(tile_off + t + 1);
    }
    return o;
  };
  auto issue = [&](uint32_t t, uint32_t b, Off o) {
    if (lane != 0 || t >= t_end) return;
    const uint32_t bytes = (o.o1 - o.o0) << 4;
    if (bytes > (uint32_t)S::BUF_BYTES) return;
    const uint32_t bar = bar0 + 8 * b;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(wbase + (size_t)b * S::BUF_BYTES)),
                 "l"(packed + ((uint64_t)o.o0 << 4)), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
  };

  Off q0 = offs(t_begin + wg), q1 = offs(t_begin + wg + ws), q2 = offs(t_begin + wg + 2 * ws);
  issue(t_begin + wg, 0, q0);
  issue(t_begin + wg + ws, 1, q1);
  uint32_t phase0 = 0u, phase1 = 0u, b = 0u;
  double dot = 0.0;
  for (uint32_t v = t_begin + wg; v < t_end; v += ws) {
    const Off qn = offs(v + 3 * ws);  // requested now, used at the end of the next iteration
    const uint32_t t = tile_of(v);
    // boundary tiles (the only ones that read ghost columns) take their own copy of the row code, so that the interior tiles run
    // exactly the single-domain instruction stream (the address select per gather measured 6-14 % when every tile paid for it)
    const bool bnd = HALO && (t < t_lo || t >= t_hi);
    if (HALO && bnd && !waited) {
      // first boundary tile of this warp: the ghost values must have landed (all neighbours, see peer.cuh)
      int ok = 1;
      if (lane == 0) ok = halo_wait(a.hw, hseq) ? 1 : 0;
      ok = __shfl_sync(0xffffffffu, ok, 0);
      __syncwarp();  // memory ordering: lane 0's acquire covers the whole warp's ghost loads
      if (!ok && lane == 0) a.st->err = 2;
      waited = true;
    }
    const uint32_t r = t * RPT + rl;
    const bool valid = r >= a.row_begin && r < a.row_end;
    const uint32_t bytes = (q0.o1 - q0.o0) << 4;
    if (bytes <= (uint32_t)S::BUF_BYTES) {
      unsigned char *buf = wbase + (size_t)b * S::BUF_BYTES;
      const uint32_t bar = bar0 + 8 * b;
      const uint32_t ph = b ? phase1 : phase0;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "PK_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra PK_DONE;\n"
          "bra PK_WAIT;\n"
          "PK_DONE:\n"
          "}\n" ::"r"(bar),
          "r"(ph)
          : "memory");
      if (b) phase1 ^= 1u;
      else phase0 ^= 1u;
      const uint32_t *hdr = reinterpret_cast<const uint32_t *>(buf);
      const uint32_t nb = hdr[RPT];
      const uint32_t kb = hdr[rl], ke = hdr[rl + 1];
      const uint32_t *sc = reinterpret_cast<const uint32_t *>(buf + PK_HDR);
      const double *sd0 = reinterpret_cast<const double *>(buf + PK_HDR + ((nb * 4u + 15u) & ~15u));
      const double *sv = sd0 + (has_dia ? RPT * NN : 0);
      const double *sd = has_dia ? sd0 + rl * NN : nullptr;
      if (HALO && bnd) dot += ring_row<N, LPR, true>(a, sv, sc, sd, valid ? kb : 0u, valid ? ke : 0u, r, valid, sub, xg);
      else dot += ring_row<N, LPR, false>(a, sv, sc, sd, valid ? kb : 0u, valid ? ke : 0u, r, valid, sub);
    } else {
      // oversized tile: straight from the canonical arrays
      const uint32_t kb = valid ? a.row2idx[r] : 0u, ke = valid ? a.row2idx[r + 1] : 0u;
      const double *gd = (valid && a.row2val) ? a.row2val + (uint64_t)r * NN : nullptr;
      if (HALO && bnd) dot += ring_row<N, LPR, true>(a, a.idx2val, a.idx2col, gd, kb, ke, r, valid, sub, xg);
      else dot += ring_row<N, LPR, false>(a, a.idx2val, a.idx2col, gd, kb, ke, r, valid, sub);
    }
    __syncwarp();
    issue(v + 2 * ws, b, q2);
    q0 = q1;
    q1 = q2;
    q2 = qn;
    b ^= 1u;
  }
  if (a.fuse_dot) spmv_publish_dot(a, dot, s_buf);
}

template <int N, int RPT>
static dfb_status spmv_launch_packed(dfb_ctx *c, const SpmvArgs &a, const dfb_bsr *m, cudaStream_t strm) {
  using S = PackedCfg<N, RPT>;
  const int warps = S::WARPS;
  const size_t smem = (size_t)warps * S::WARP_BYTES;
  const uint64_t t_begin = a.row_begin / RPT, t_end = ((uint64_t)a.row_end + RPT - 1) / RPT;
  uint64_t g = (t_end - t_begin + warps - 1) / warps;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count) g = (uint64_t)c->sm_count;
  // (the dynamic shared-memory opt-in is a per-device function attribute: set on every launch, cheap)
  if (a.hw.xg[0] != nullptr) {
    DFB_CUDA(cudaFuncSetAttribute(k_spmv_packed<N, true, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFB_LAUNCH_ON(c, strm, (k_spmv_packed<N, true, RPT>), (unsigned)g, warps * 32, smem, a, m->d_packed, m->d_tile_off, bsr_has_dia(m) ? 1 : 0);
  } else {
    DFB_CUDA(cudaFuncSetAttribute(k_spmv_packed<N, false, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFB_LAUNCH_ON(c, strm, (k_spmv_packed<N, false, RPT>), (unsigned)g, warps * 32, smem, a, m->d_packed, m->d_tile_off, bsr_has_dia(m) ? 1 : 0);
  }
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

template <int N>
static dfb_status spmv_dispatch(dfb_ctx *c, const SpmvArgs &a, cudaStream_t strm) {
  if (c->spmv_variant == 4) return spmv_launch_ring<N>(c, a, strm);
  if (c->spmv_variant != 0) return dfb_fail(DFB_ERR_BAD_ARG, "spmv_variant %lld: only 0 (row per thread), 4 (TMA ring) and 20 (packed) exist", (long long)c->spmv_variant);
  const uint64_t n_rows = a.row_end - a.row_begin;
  if (n_rows == 0 && !a.fuse_dot) return DFB_OK;
  uint64_t g = (n_rows + 127) / 128;
  if (g < 1) g = 1;
  const uint64_t cap = (uint64_t)c->sm_count * 16;
  if (g > cap) g = cap;
  if (g > DFB_RED_MAX_BLOCKS) g = DFB_RED_MAX_BLOCKS;
  DFB_LAUNCH_ON(c, strm, k_spmv_row<N>, (unsigned)g, 128, 0, a);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// internal launcher shared with the solvers
dfb_status dfb_spmv_launch(const dfb_bsr *m, double *y, double beta, double alpha, const double *x, uint64_t row_begin,
                           uint64_t row_end, int fuse_dot, int in_solver, cudaStream_t strm, int post, const DfbHaloWait *hw) {
  dfb_ctx *c = m->ctx;
  SpmvArgs a;
  a.row2idx = m->pat->d_row2idx;
  a.idx2col = m->pat->d_idx2col;
  if (c->spmv_variant == 20) DFB_TRY(dfb_bsr_ensure_packed(m));
  // variants 0 / 4, and the oversized tiles of variant 20, multiply straight from the canonical arrays
  if (c->spmv_variant != 20 || m->packed_ok != 1) DFB_TRY(dfb_bsr_canon_read(m));
  a.idx2val = m->d_idx2val;
  a.row2val = bsr_has_dia(m) ? m->d_row2val : nullptr;
  a.x = x;
  a.y = y;
  a.alpha = alpha;
  a.beta = beta;
  a.row_begin = (uint32_t)row_begin;
  a.row_end = (uint32_t)row_end;
  a.fuse_dot = fuse_dot;
  a.st = in_solver ? c->d_state : nullptr;  // standalone mult_vec: no solver state
  if ((fuse_dot || hw) && !a.st) return dfb_fail(DFB_ERR_BAD_ARG, "spmv: fused dot / peer halo need the solver state");
  a.partials = c->d_partials;
  a.ticket = c->d_ticket + 1;
  a.pv = dfb_peer_view(c, post);
  if (hw) {
    a.hw = *hw;
  } else {
    a.hw.xg[0] = a.hw.xg[1] = nullptr;
    a.hw.n_own = 0xFFFFFFFFu;
    a.hw.wait_mask = 0;
    a.hw.halo_flag = nullptr;
    a.hw.tile_lo_end = 0;
    a.hw.tile_hi_begin = 0xFFFFFFFFu;
  }
  if (c->spmv_variant == 20) {
    if (m->pk_rpt == 16) {
      switch (m->blk_dim) {
        case 1: return spmv_launch_packed<1, 16>(c, a, m, strm);
        case 2: return spmv_launch_packed<2, 16>(c, a, m, strm);
        case 3: return spmv_launch_packed<3, 16>(c, a, m, strm);
        case 4: return spmv_launch_packed<4, 16>(c, a, m, strm);
      }
    }
    switch (m->blk_dim) {
      case 1: return spmv_launch_packed<1, 8>(c, a, m, strm);
      case 2: return spmv_launch_packed<2, 8>(c, a, m, strm);
      case 3: return spmv_launch_packed<3, 8>(c, a, m, strm);
      case 4: return spmv_launch_packed<4, 8>(c, a, m, strm);
    }
  }
  switch (m->blk_dim) {
    case 1: return spmv_dispatch<1>(c, a, strm);
    case 2: return spmv_dispatch<2>(c, a, strm);
    case 3: return spmv_dispatch<3>(c, a, strm);
    case 4: return spmv_dispatch<4>(c, a, strm);
  }
  return dfb_fail(DFB_ERR_UNSUPPORTED, "spmv: blk_dim %d", m->blk_dim);
}

static uint64_t halo_recv_total(const dfb_bsr *m) { return (m->halo && !m->halo->recv_ptr.empty()) ? m->halo->recv_ptr.back() : 0; }

// true when (P)CG exchanges the ghosts of its direction vector by the NVLink peer push (one SpMV launch over all rows, boundary
// tiles wait for the neighbours' flags) instead of ncclSend/ncclRecv on the second stream
bool dfb_spmv_uses_peer_halo(const dfb_bsr *m) {
  const dfb_ctx *c = m->ctx;
  return m->halo && c->nranks > 1 && c->spmv_variant == 20 && c->peer_enabled && c->use_peer_allreduce &&
         dfb_halo_peer_usable(m->halo, m->blk_dim);
}

// y <- alpha*A*x + beta*y  with halo exchange of x when a halo is attached.
//   in_solver + peer halo: the producer of x has already pushed its boundary entries (dfb_halo_push); ONE launch over all rows
//   otherwise: pack + ncclSend/ncclRecv on the comm stream, overlapped with the interior rows; boundary rows follow the event
dfb_status dfb_spmv_full(const dfb_bsr *m, double *y, double beta, double alpha, double *x, int fuse_dot, int in_solver, int post) {
  dfb_ctx *c = m->ctx;
  const uint64_t n = m->pat->num_blk;
  if (!m->halo || c->nranks <= 1) return dfb_spmv_launch(m, y, beta, alpha, x, 0, n, fuse_dot, in_solver, c->stream, post);
  if (in_solver && dfb_spmv_uses_peer_halo(m)) {
    DfbHaloWait hw;
    dfb_halo_wait_view(m->halo, m->blk_dim, n, &hw);
    hw.tile_lo_end = (uint32_t)((m->int_begin + m->pk_rpt - 1) / m->pk_rpt);
    hw.tile_hi_begin = (uint32_t)(m->int_end / m->pk_rpt);
    return dfb_spmv_launch(m, y, beta, alpha, x, 0, n, fuse_dot, in_solver, c->stream, post, &hw);
  }
  // comm stream: wait until x is final, exchange ghosts
  DFB_CUDA(cudaEventRecord(c->ev_a, c->stream));
  DFB_CUDA(cudaStreamWaitEvent(c->stream_comm, c->ev_a, 0));
  DFB_TRY(dfb_halo_exchange_on(m->halo, x, n, m->blk_dim, c->stream_comm));
  DFB_CUDA(cudaEventRecord(c->ev_b, c->stream_comm));
  // interior rows do not read ghosts; the launch that completes the fused dot product also posts it to the peers
  const bool has_lo = m->int_begin > 0, has_hi = m->int_end < n, has_int = m->int_end > m->int_begin;
  int mode = fuse_dot;
  if (has_int) {
    DFB_TRY(dfb_spmv_launch(m, y, beta, alpha, x, m->int_begin, m->int_end, mode, in_solver, c->stream, (!has_lo && !has_hi) ? post : 0));
    if (mode == 1) mode = 2;
  }
  DFB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_b, 0));
  if (has_lo) {
    DFB_TRY(dfb_spmv_launch(m, y, beta, alpha, x, 0, m->int_begin, mode, in_solver, c->stream, !has_hi ? post : 0));
    if (mode == 1) mode = 2;
  }
  if (has_hi) {
    DFB_TRY(dfb_spmv_launch(m, y, beta, alpha, x, m->int_end, n, mode, in_solver, c->stream, post));
    if (mode == 1) mode = 2;
  }
  if (!has_int && !has_lo && !has_hi && fuse_dot) DFB_TRY(dfb_spmv_launch(m, y, beta, alpha, x, 0, 0, fuse_dot, in_solver, c->stream, post));
  return DFB_OK;
}

DFB_API dfb_status dfb_bsr_mult_vec(const dfb_bsr *m, dfb_vec *y, double beta, double alpha, dfb_vec *x) {
  DFB_REQUIRE(m && y && x, "dfb_bsr_mult_vec: NULL argument");
  DFB_REQUIRE(y->n_blk == m->pat->num_blk && x->n_blk == m->pat->num_blk, "dfb_bsr_mult_vec: y/x length != num_blk (assert_eq!(y_vec.len(), self.num_blk))");
  DFB_REQUIRE(y->blk_dim == m->blk_dim && x->blk_dim == m->blk_dim, "dfb_bsr_mult_vec: vector block dim != matrix block dim");
  DFB_REQUIRE(x->d != y->d, "dfb_bsr_mult_vec: x and y alias");
  if (m->halo && m->ctx->nranks > 1)
    DFB_REQUIRE(x->n_ghost >= halo_recv_total(m), "dfb_bsr_mult_vec: the matrix has a halo of %llu ghost blocks but x was created with n_ghost = %llu",
                (unsigned long long)halo_recv_total(m), (unsigned long long)x->n_ghost);
  return dfb_spmv_full(m, y->d, beta, alpha, x->d, 0, 0, 0);
}

// del-fem-ls mult_mat (del-fem-ls/src/sparse_square.rs:134-164): a SCALAR matrix applied to num_dim interleaved components,
// y[i][d] <- beta y[i][d] + sum_j (alpha a_ij) x[j][d] + (alpha d_i) x[i][d], terms in the reference's order (off-diagonals in
// storage order, diagonal last). One thread per (row, component); off the hot path (ARAP-style uses), kept simple.
__global__ void k_mult_mat_scalar(const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col, const double *__restrict__ idx2val,
                                  const double *__restrict__ row2val, uint64_t num_rows, int nd, double beta, double alpha,
                                  const double *__restrict__ x, double *__restrict__ y) {
  const uint64_t total = num_rows * (uint64_t)nd;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = q / nd;
    const int d = (int)(q - i * nd);
    double acc = y[q] * beta;
    for (uint32_t k = row2idx[i]; k < row2idx[i + 1]; ++k) acc += alpha * idx2val[k] * x[(uint64_t)idx2col[k] * nd + d];
    if (row2val) acc += alpha * row2val[i] * x[q];
    y[q] = acc;
  }
}

DFB_API dfb_status dfb_bsr_mult_mat(const dfb_bsr *m, dfb_vec *y, double beta, double alpha, const dfb_vec *x) {
  DFB_REQUIRE(m && y && x, "dfb_bsr_mult_mat: NULL argument");
  DFB_REQUIRE(m->blk_dim == 1, "dfb_bsr_mult_mat: the matrix must be scalar (blk_dim 1), got %d", m->blk_dim);
  DFB_REQUIRE(y->n_blk == m->pat->num_blk && x->n_blk == m->pat->num_blk && x->blk_dim == y->blk_dim,
              "dfb_bsr_mult_mat: assert_eq!(y_mat.len(), x_mat.len()) / num_dim * num_row");
  DFB_REQUIRE(x->d != y->d, "dfb_bsr_mult_mat: x and y alias");
  DFB_REQUIRE(!m->halo, "dfb_bsr_mult_mat: not available for distributed matrices");
  DFB_TRY(dfb_bsr_canon_read(m));
  dfb_ctx *c = m->ctx;
  const uint64_t total = m->pat->num_blk * (uint64_t)x->blk_dim;
  uint64_t g = (total + 255) / 256;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
  DFB_LAUNCH(c, k_mult_mat_scalar, (unsigned)g, 256, 0, m->pat->d_row2idx, m->pat->d_idx2col, m->d_idx2val, m->d_row2val, m->pat->num_blk,
             x->blk_dim, beta, alpha, x->d, y->d);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

__global__ void k_max_u32(const uint32_t *__restrict__ a, uint64_t n, unsigned int *out) {
  unsigned int m = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, a[i]);
  for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

DFB_API dfb_status dfb_bsr_set_halo(dfb_bsr *m, dfb_halo *h, uint64_t int_begin, uint64_t int_end) {
  DFB_REQUIRE(m, "dfb_bsr_set_halo: NULL argument");
  DFB_REQUIRE(int_begin <= int_end && int_end <= m->pat->num_blk, "dfb_bsr_set_halo: bad interior range");
  if (h && m->pat->nnz > 0) {
    // every column the local rows reference must be an owned row or one of the ghosts the halo receives
    dfb_ctx *c = m->ctx;
    unsigned int *d_max = (unsigned int *)(c->d_scratch + 16);
    DFB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned int), c->stream));
    DFB_LAUNCH(c, k_max_u32, c->sm_count * 4, 256, 0, m->pat->d_idx2col, m->pat->nnz, d_max);
    DFB_CHECK_LAUNCH();
    unsigned int h_max = 0;
    DFB_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    DFB_CUDA(cudaStreamSynchronize(c->stream));
    const uint64_t n_all = m->pat->num_blk + (h->recv_ptr.empty() ? 0 : h->recv_ptr.back());
    DFB_REQUIRE((uint64_t)h_max < n_all, "dfb_bsr_set_halo: the pattern references column %u but the matrix has only %llu owned rows + %llu ghosts",
                h_max, (unsigned long long)m->pat->num_blk, (unsigned long long)(n_all - m->pat->num_blk));
  }
  m->halo = h;
  m->int_begin = int_begin;
  m->int_end = int_end;
  return DFB_OK;
}
