// assemble_tile.cu — fused row-tile assembly: element kernel + merge in ONE launch, no global scratch.     (compiled --fmad=false)
//
// Reference being replaced: the serial element loop + scatter-add of sparse_square.rs:180-214 / merge.rs:3-88, whose accumulation
// order into a given nonzero is "ascending element index".  The slot-centric gather of assemble.cu reproduces that order but is
// bound by L1 sector throughput (ncu, profiles/r1_assemble_recs_ncu.txt: every contribution fetches two scattered 32-byte node
// records from global memory, 2.6x the algorithmic DRAM traffic, 72-byte-strided stores 4x amplified).  Here
//   * a CTA owns a tile of AT_ROWS consecutive block-rows; the plan lists the tile's UNIQUE incident elements once;
//   * phase 1: one thread per listed element evaluates the per-element record (geometry / constitutive data) into SHARED memory
//     — elements are re-evaluated by the ~3 tiles that touch them, which costs flops instead of a 128 B/tet scratch round trip;
//   * phase 2: one thread per nonzero slot of the tile walks its contribution list (16-bit codes: tile-local element, i, j) in
//     ascending element order, forms block (i,j) from the shared-memory record and accumulates in registers: same arithmetic, same
//     order => same bits as the serial loop; the thread of a row's diagonal slot also gathers the row's gradient (energy models);
//   * phase 3: the tile's values leave through shared memory as contiguous, coalesced stores — straight into the tile-packed SpMV
//     records when those are the matrix's native layout (no pack pass, no second copy), else into the canonical arrays.
// Measured on S10 (19.8 M tets, profiles/r2_assemble_tile_*): 4.5 ms (1.2 TB/s algorithmic), DRAM traffic == algorithmic bytes;
// bound by shared-memory wavefronts of the record reads (random element per lane) with the fp64 pipe at ~30 %: the bit-exact
// formulas cost ~74 non-fused fp64 instructions per contribution, 1.6 ms of fp64 issue alone. Variants that were built, measured
// SLOWER and removed (gpurun_out logs summarised in profiles/README.md): staging the tile's lists in shared memory (+2..12 %),
// 4 lanes per diagonal slot with ordered shuffles (+27..50 %), staging vertex coordinates through a per-tile vertex list (+0..7 %),
// 16-double node records read with 16-byte loads (every lane on one bank: +45 %), tiles of 24 / 32 rows (+29 % / +70 %).
#include "dfb_internal.h"
#include "elements.cuh"
#include "reduce.cuh"

static const int AT_MAX_THREADS = 384;
static const int AT_BUILD_THREADS = 256;
static const int AT_PAIR_CAP = 2048;  // (row, element) pairs per tile the builder can sort in shared memory
static const int AT_ELEM_CAP = 4095;  // tile-local element index must fit 12 bits of the 16-bit contribution code
static const int AT_META = 8;         // u32 per tile: k0, k1, ob, oe, pb, pe, n_elem, pad
static const int AT_PKT = 10;         // packed records per assembly tile + 1 (a tile of up to 64 rows over 8-row records)

// ------------------------------------------------------------------------------------------------ tile plan
__device__ __forceinline__ void bitonic_sort_smem(uint32_t *keys, uint32_t P) {
  for (uint32_t k = 2; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const uint32_t a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            keys[i] = b;
            keys[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// sorted keys[0..n) -> unique values in out[0..total) (block-wide; cnt = scratch of blockDim.x entries)
__device__ __forceinline__ uint32_t unique_sorted_smem(const uint32_t *keys, uint32_t n, uint32_t P, uint32_t *out, uint32_t *cnt) {
  const unsigned tid = threadIdx.x, T = blockDim.x;
  const uint32_t C = (P + T - 1) / T;
  const uint32_t cb = min(tid * C, n), ce = min(cb + C, n);
  uint32_t local = 0;
  for (uint32_t i = cb; i < ce; ++i) local += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  cnt[tid] = local;
  __syncthreads();
  for (uint32_t off = 1; off < T; off <<= 1) {  // inclusive Hillis-Steele scan of the per-thread counts
    const uint32_t v = (tid >= off) ? cnt[tid - off] : 0u;
    __syncthreads();
    cnt[tid] += v;
    __syncthreads();
  }
  const uint32_t total = cnt[T - 1];
  uint32_t w = cnt[tid] - local;
  for (uint32_t i = cb; i < ce; ++i)
    if (i == 0 || keys[i] != keys[i - 1]) out[w++] = keys[i];
  __syncthreads();
  return total;
}

__device__ __forceinline__ uint32_t lower_bound_smem(const uint32_t *a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// one CTA per tile: the tile's sorted unique element list, its meta record and the 16-bit recoding of its contribution lists
__global__ void __launch_bounds__(AT_BUILD_THREADS) k_tile_build(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib,
                                                                 uint64_t nnz, uint64_t num_rows, int nn, const uint32_t *__restrict__ row2idx,
                                                                 uint32_t *__restrict__ tile_elem, uint32_t *__restrict__ tile_meta,
                                                                 uint16_t *__restrict__ code16, uint32_t *stats, uint64_t n_tiles,
                                                                 uint32_t AT_ROWS) {
  __shared__ uint32_t keys[AT_PAIR_CAP];
  __shared__ uint32_t uniq[AT_PAIR_CAP];
  __shared__ uint32_t cnt[AT_BUILD_THREADS];
  const uint32_t pair_base = nz2ptr[nnz];
  const uint32_t nn2 = (uint32_t)(nn * nn);
  const unsigned tid = threadIdx.x;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t r0 = tile * AT_ROWS, r1 = min(r0 + (uint64_t)AT_ROWS, num_rows);
    const uint32_t pb = nz2ptr[nnz + r0], pe = nz2ptr[nnz + r1], np = pe - pb;
    const uint32_t k0 = row2idx[r0], k1 = row2idx[r1];
    if (np > (uint32_t)AT_PAIR_CAP) {
      if (tid == 0) atomicExch(&stats[2], 1u);
      continue;
    }
    uint32_t P = 1;
    while (P < np) P <<= 1;
    for (uint32_t i = tid; i < P; i += AT_BUILD_THREADS) keys[i] = (i < np) ? contrib[pb + i] / nn2 : 0xFFFFFFFFu;
    __syncthreads();
    bitonic_sort_smem(keys, P);
    const uint32_t n_le = unique_sorted_smem(keys, np, P, uniq, cnt);
    const uint32_t eoff = pb - pair_base, ob = nz2ptr[k0], oe = nz2ptr[k1];
    if (tid == 0) {
      uint32_t *mt = tile_meta + tile * AT_META;
      mt[0] = k0; mt[1] = k1; mt[2] = ob; mt[3] = oe; mt[4] = pb; mt[5] = pe; mt[6] = n_le; mt[7] = 0;
      atomicMax(&stats[0], n_le);
      atomicMax(&stats[1], (k1 - k0) + (uint32_t)(r1 - r0));
      if (n_le > (uint32_t)AT_ELEM_CAP) atomicExch(&stats[2], 1u);
    }
    for (uint32_t i = tid; i < n_le; i += AT_BUILD_THREADS) tile_elem[eoff + i] = uniq[i];
    // recode the contributions of the tile's slots: off-diagonal slots [k0, k1), then the diagonal slots of rows [r0, r1)
    const uint32_t total = (oe - ob) + np;
    for (uint32_t q = tid; q < total; q += AT_BUILD_THREADS) {
      const uint32_t c = (q < oe - ob) ? ob + q : pb + (q - (oe - ob));
      const uint32_t code = contrib[c];
      const uint32_t e = code / nn2, ij = code - e * nn2;
      const uint32_t i = ij / (uint32_t)nn, j = ij - i * (uint32_t)nn;
      const uint32_t lo = lower_bound_smem(uniq, n_le, e);
      if (lo >= n_le || uniq[lo] != e) atomicExch(&stats[2], 1u);  // a contribution whose (element, row) is not a pair of the row
      code16[c] = (uint16_t)((lo << 4) | (i << 2) | j);
    }
    __syncthreads();
  }
}

static void plan_free_tiles(dfb_plan *plan) {
  dfb_dev_free(plan->d_tile_elem);
  dfb_dev_free(plan->d_tile_meta);
  dfb_dev_free(plan->d_code16);
  plan->d_tile_elem = plan->d_tile_meta = nullptr;
  plan->d_code16 = nullptr;
}
void dfb_plan_free_tiles(dfb_plan *plan) { plan_free_tiles(plan); }

dfb_status dfb_plan_build_tiles(dfb_plan *plan) {
  if (plan->tile_state != 0) return DFB_OK;
  plan->tile_state = -1;
  dfb_ctx *c = plan->ctx;
  if (plan->pat->convention != 0 || plan->flavour != 0 || plan->num_node > 4 || plan->pat->num_blk == 0) return DFB_OK;
  const uint64_t nnz = plan->pat->nnz, nr = plan->pat->num_blk;
  uint32_t ends[2] = {0, 0};
  DFB_CUDA(cudaMemcpyAsync(&ends[0], plan->d_nz2ptr + nnz, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaMemcpyAsync(&ends[1], plan->d_nz2ptr + nnz + nr, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  const uint64_t npairs = ends[1] - ends[0];
  // 16 rows per tile (two of the SpMV's 8-row tiles): measured best for tets (~240 slots) and for spring edges (~112 slots);
  // 8 for very dense rows so that a tile's slots still fit one CTA
  uint32_t AT_ROWS = ((double)nnz / (double)nr > 28.0) ? 8 : 16;
  if (c->assemble_tile_rows > 0) AT_ROWS = (uint32_t)((c->assemble_tile_rows + 7) / 8 * 8);
  if (AT_ROWS > 64) AT_ROWS = 64;
  const uint64_t n_tiles = (nr + AT_ROWS - 1) / AT_ROWS;
  uint32_t *d_stats = nullptr;
  dfb_status st = dfb_dev_alloc_t(&plan->d_tile_elem, npairs + 2);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&plan->d_tile_meta, (n_tiles + 1) * AT_META);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&plan->d_code16, plan->num_contrib + 8);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_stats, 4);
  uint32_t h_stats[4] = {0, 0, 1, 0};
  if (st == DFB_OK) {
    cudaMemsetAsync(d_stats, 0, 4 * sizeof(uint32_t), c->stream);
    cudaMemsetAsync(plan->d_tile_meta, 0, (n_tiles + 1) * AT_META * sizeof(uint32_t), c->stream);
    const uint64_t g = n_tiles < (uint64_t)c->sm_count * 8 ? n_tiles : (uint64_t)c->sm_count * 8;
    DFB_LAUNCH(c, k_tile_build, (unsigned)g, AT_BUILD_THREADS, 0, plan->d_nz2ptr, plan->d_contrib, nnz, nr, plan->num_node, plan->pat->d_row2idx,
               plan->d_tile_elem, plan->d_tile_meta, plan->d_code16, d_stats, n_tiles, AT_ROWS);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_plan_build_tiles: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(c->stream);
  dfb_dev_free(d_stats);
  if (st != DFB_OK || h_stats[2] != 0) {
    // not applicable (very high valence / degenerate lists): the slot-centric gather of assemble.cu stays in charge
    plan_free_tiles(plan);
    return st;
  }
  plan->tile_max_elem = h_stats[0];
  plan->tile_max_slots = h_stats[1];
  plan->tile_rows = AT_ROWS;
  plan->n_atiles = n_tiles;
  plan->pair_base = ends[0];
  plan->tile_state = 1;
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ models
// A model provides: NNODE, N, NDIM, REC (doubles per element record), HAS_GRAD, NEEDS_DISP, Args (kernel parameters),
//   make(rec, px, pu, args)           -> evaluates the record of one element into shared memory; px / pu = per-node pointers to the
//                                        coordinates / displacements (staged in shared memory, or global); returns the element's energy
//   block(rec, i, j, args, blk)       -> block (i,j) of the element matrix, column-major N x N
//   grad(rec, i, args, g)             -> gradient of node i (energy models)

// linear kinds: the record is ElemGeo<KIND> itself and block() IS the reference formula (ElemGeo<KIND>::block). The record stride
// is an ODD number of doubles (tet: 13): threads of a warp read the records of different elements at the same offset, and an odd
// stride spreads those reads over the shared-memory banks. (ncu: 16-double node records — two 16-byte loads per node — put every
// lane on the same bank: 1.33 G shared-memory wavefronts, L1 81 % busy, 6.6 ms; the 13-double records: 0.75 G.)
template <int KIND>
struct ModelLinear {
  using G = ElemGeo<KIND>;
  static constexpr int NNODE = G::NNODE, N = G::N, NDIM = G::NDIM, HAS_GRAD = 0, NEEDS_DISP = 0;
  static constexpr int REC = (int)((sizeof(G) + 7) / 8) | 1;
  struct Args {
    double p0, p1;
  };
  __device__ static double make(double *rec, const double *const *p, const double *const *, const Args &a) {
    G g;
    g.init(p, a.p0, a.p1);
    *reinterpret_cast<G *>(rec) = g;
    return 0.0;
  }
  __device__ static void block(const double *rec, int i, int j, const Args &a, double *blk) {
    reinterpret_cast<const G *>(rec)->block(i, j, a.p0, a.p1, blk);
  }
  __device__ static void grad(const double *, int, const Args &, double *) {}
};

// spring3::wdwddw_squared_length_difference (del-fem-cpu/src/spring3.rs:1-28): same operations as k_spring3 (assemble.cu)
struct ModelSpring3 {
  static constexpr int NNODE = 2, N = 3, NDIM = 3, REC = 7, HAS_GRAD = 1, NEEDS_DISP = 1;
  struct Args {
    double stiffness;
  };
  // rec: v[3], t0, t1, gp (= c k / l)
  __device__ static double make(double *rec, const double *const *px, const double *const *pu, const Args &a) {
    const double *x0 = px[0], *x1 = px[1], *u0 = pu[0], *u1 = pu[1];
    double r[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      r[k] = x0[k] - x1[k];
      v[k] = (x0[k] + u0[k]) - (x1[k] + u1[k]);
    }
    const double l0 = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double c = l0 - l;
    rec[0] = v[0];
    rec[1] = v[1];
    rec[2] = v[2];
    rec[3] = a.stiffness * l0 / (l * l * l);
    rec[4] = a.stiffness * (l - l0) / l;
    rec[5] = c * a.stiffness / l;
    return 0.5 * a.stiffness * c * c;
  }
  __device__ static void block(const double *rec, int i, int j, const Args &, double *blk) {
    const double t0 = rec[3], t1 = rec[4];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double m = (rec[a] * rec[b]) * t0 + ((a == b) ? t1 : 0.0);
        blk[a + 3 * b] = (i == j) ? m : -m;
      }
  }
  __device__ static void grad(const double *rec, int i, const Args &, double *g) {
    const double gp = rec[5];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = rec[k] * (i == 0 ? -gp : gp);
  }
};

// DERIVED tet neo-Hookean: the reference's chain rule (solid_incompressive_hex.rs:86-148) with 4 nodes; the same
// operations as k_neohookean_tet (assemble.cu). Record: g[4][3], z = F [3][3], dwrdc[6], ddwrddc[6][6], detwei.
struct ModelNeoHookeanTet {
  static constexpr int NNODE = 4, N = 3, NDIM = 3, REC = 65, HAS_GRAD = 1, NEEDS_DISP = 1;
  struct Args {
    double myu, lambda;
  };
  __device__ static double make(double *rec, const double *const *px, const double *const *pu, const Args &a) {
    const int IJ[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {1, 2}, {2, 0}};  // dudx.rs:36
    double X[4][3], U[4][3];
    for (int n = 0; n < 4; ++n)
      for (int d = 0; d < 3; ++d) {
        X[n][d] = px[n][d];
        U[n][d] = pu[n][d];
      }
    double g[4][3];
    const double vol = tet_vol_grad(g, X[0], X[1], X[2], X[3]);
    double dudx[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;
        for (int n = 0; n < 4; ++n) t += U[n][i] * g[n][j];
        dudx[i][j] = t;
      }
    double C[3][3], Ci[3][3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        double t = dudx[i][j] + dudx[j][i];
        for (int k = 0; k < 3; ++k) t += dudx[k][i] * dudx[k][j];
        C[i][j] = t;
      }
      C[i][i] += 1.0;
    }
    const double d = C[0][0] * (C[1][1] * C[2][2] - C[1][2] * C[2][1]) - C[0][1] * (C[1][0] * C[2][2] - C[1][2] * C[2][0]) +
                     C[0][2] * (C[1][0] * C[2][1] - C[1][1] * C[2][0]);
    const double t = 1.0 / d;
    Ci[0][0] = t * (C[1][1] * C[2][2] - C[1][2] * C[2][1]);
    Ci[0][1] = t * (C[0][2] * C[2][1] - C[0][1] * C[2][2]);
    Ci[0][2] = t * (C[0][1] * C[1][2] - C[0][2] * C[1][1]);
    Ci[1][0] = t * (C[1][2] * C[2][0] - C[1][0] * C[2][2]);
    Ci[1][1] = t * (C[0][0] * C[2][2] - C[0][2] * C[2][0]);
    Ci[1][2] = t * (C[0][2] * C[1][0] - C[0][0] * C[1][2]);
    Ci[2][0] = t * (C[1][0] * C[2][1] - C[1][1] * C[2][0]);
    Ci[2][1] = t * (C[0][1] * C[2][0] - C[0][0] * C[2][1]);
    Ci[2][2] = t * (C[0][0] * C[1][1] - C[0][1] * C[1][0]);
    const double lnj = 0.5 * log(d);
    const double wr = 0.5 * a.myu * (C[0][0] + C[1][1] + C[2][2] - 3.0) - a.myu * lnj + 0.5 * a.lambda * lnj * lnj;
    const double sco = 0.5 * a.lambda * lnj - 0.5 * a.myu;
    const double tco = 0.5 * a.myu - 0.5 * a.lambda * lnj;
    double *rg = rec, *rz = rec + 12, *rdw = rec + 21, *rddw = rec + 27;
    for (int n = 0; n < 4; ++n)
      for (int k = 0; k < 3; ++k) rg[n * 3 + k] = g[n][k];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) rz[i * 3 + j] = dudx[i][j] + ((i == j) ? 1.0 : 0.0);  // def_grad_tensor dudx.rs:1-14
    for (int p = 0; p < 6; ++p) {
      const int i = IJ[p][0], j = IJ[p][1];
      rdw[p] = ((i == j) ? 0.5 * a.myu : 0.0) + sco * Ci[i][j];
      for (int q = 0; q < 6; ++q) {
        const int k = IJ[q][0], l = IJ[q][1];
        const double v1 = 0.25 * a.lambda * Ci[i][j] * Ci[k][l];
        const double v2 = 0.5 * tco * Ci[i][l] * Ci[j][k];
        const double v3 = 0.5 * tco * Ci[i][k] * Ci[j][l];
        rddw[p * 6 + q] = v1 + v2 + v3;
      }
    }
    rec[63] = vol;
    return 0.0 + wr * vol;
  }
  __device__ __forceinline__ static void dcdu(const double *rec, int n, double (&dc)[3][6]) {
    const double *g = rec + n * 3, *z = rec + 12;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      dc[d][0] = g[0] * z[d * 3 + 0];
      dc[d][1] = g[1] * z[d * 3 + 1];
      dc[d][2] = g[2] * z[d * 3 + 2];
      dc[d][3] = g[0] * z[d * 3 + 1] + g[1] * z[d * 3 + 0];
      dc[d][4] = g[1] * z[d * 3 + 2] + g[2] * z[d * 3 + 1];
      dc[d][5] = g[2] * z[d * 3 + 0] + g[0] * z[d * 3 + 2];
    }
  }
  __device__ static void block(const double *rec, int ino, int jno, const Args &, double *out) {
    const double *gi = rec + ino * 3, *gj = rec + jno * 3, *dwrdc = rec + 21, *ddw = rec + 27;
    const double detwei = rec[63];
    double di[3][6], dj[3][6];
    dcdu(rec, ino, di);
    dcdu(rec, jno, dj);
    double blk[9];  // reference convention a*3+b
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double dtmp1 = 0.0;
        for (int gg = 0; gg < 6; ++gg)
          for (int hh = 0; hh < 6; ++hh) dtmp1 += 2.0 * ddw[gg * 6 + hh] * di[a][gg] * dj[b][hh];
        blk[a * 3 + b] = 0.0 + 2.0 * detwei * dtmp1;
      }
    double dtmp2 = 0.0;
    dtmp2 += dwrdc[0] * gi[0] * gj[0];
    dtmp2 += dwrdc[1] * gi[1] * gj[1];
    dtmp2 += dwrdc[2] * gi[2] * gj[2];
    dtmp2 += dwrdc[3] * (gi[0] * gj[1] + gi[1] * gj[0]);
    dtmp2 += dwrdc[4] * (gi[1] * gj[2] + gi[2] * gj[1]);
    dtmp2 += dwrdc[5] * (gi[2] * gj[0] + gi[0] * gj[2]);
    for (int a = 0; a < 3; ++a) blk[a * 3 + a] += 2.0 * detwei * dtmp2;
    for (int a = 0; a < 3; ++a)  // transpose into the column-major BSR convention
      for (int b = 0; b < 3; ++b) out[a + 3 * b] = blk[a * 3 + b];
  }
  __device__ static void grad(const double *rec, int n, const Args &, double *g) {
    const double *dwrdc = rec + 21;
    const double detwei = rec[63];
    double dn[3][6];
    dcdu(rec, n, dn);
    for (int d = 0; d < 3; ++d) {
      double dtmp1 = 0.0;
      for (int gg = 0; gg < 6; ++gg) dtmp1 += dn[d][gg] * dwrdc[gg];
      g[d] = 0.0 + 2.0 * detwei * dtmp1;
    }
  }
};

// ------------------------------------------------------------------------------------------------ the kernel
struct TileOut {      // where the tile's values go
  double *idx2val;    // canonical arrays (used when packed.base == nullptr)
  double *row2val;
  PackedView packed;  // the SpMV's native records
};
struct TilePlanView {
  const uint32_t *nz2ptr;
  const uint16_t *code16;
  const uint32_t *tile_elem, *tile_meta;
  const uint32_t *row2idx;
  uint32_t pair_base, rows, max_elem, max_slots;
  uint64_t nnz, num_rows, n_tiles;
};

template <class Model>
__global__ void __launch_bounds__(AT_MAX_THREADS) k_assemble_tile(const typename Model::Args args, const TilePlanView P,
                                                                  const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                                  const double *__restrict__ disp, const TileOut out,
                                                                  double *__restrict__ dw, double *__restrict__ w_elem) {
  constexpr int NN = Model::N * Model::N, N = Model::N, NDIM = Model::NDIM, REC = Model::REC, NNODE = Model::NNODE;
  extern __shared__ __align__(16) double at_smem[];
  double *srec = at_smem;                                                   // [n_le][REC]
  double *sout = srec + (size_t)P.max_elem * REC;                           // [n_off + n_rows][NN]
  uint32_t *smeta = reinterpret_cast<uint32_t *>(sout + (size_t)P.max_slots * NN);   // [AT_META] + packed-store info [AT_PK]
  const unsigned tid = threadIdx.x, T = blockDim.x;
  uint64_t tile = blockIdx.x;
  uint32_t meta_next = (tile < P.n_tiles && tid < AT_META) ? P.tile_meta[tile * AT_META + tid] : 0u;
  for (; tile < P.n_tiles; tile += gridDim.x) {
    if (tid < AT_META) smeta[tid] = meta_next;
    __syncthreads();
    {  // prefetch the next tile's meta record (consumed at the top of the next iteration)
      const uint64_t nt = tile + gridDim.x;
      if (nt < P.n_tiles && tid < AT_META) meta_next = P.tile_meta[nt * AT_META + tid];
    }
    const uint64_t r0 = tile * P.rows, r1 = min(r0 + (uint64_t)P.rows, P.num_rows);
    // what phase 3 needs to find its packed records — record offsets and the row pointers at the packed-tile boundaries — is
    // requested now (ncu: the dependent chain tile offset -> record header sat in front of the stores, 9 % of the warp samples)
    if (out.packed.base && tid >= 32 && tid < 32 + 2 * AT_PKT) {
      const uint32_t j = (tid - 32) % AT_PKT;
      const uint64_t prt = (uint64_t)out.packed.rpt, row = r0 + j * prt;
      if (tid < 32 + AT_PKT) smeta[AT_META + j] = row < r1 ? out.packed.tile_off[row / prt] : 0u;
      else smeta[AT_META + AT_PKT + j] = P.row2idx[min(row, P.num_rows)];
    }
    const uint32_t k0 = smeta[0], k1 = smeta[1], pb = smeta[4], n_le = smeta[6];
    const uint32_t n_off = k1 - k0, n_rows = (uint32_t)(r1 - r0), eoff = pb - P.pair_base;
    // ---- phase 1: per-element records into shared memory (one thread per listed element)
    for (uint32_t le = tid; le < n_le; le += T) {
      const uint32_t e = P.tile_elem[eoff + le];
      const double *px[NNODE], *pu[NNODE];
      uint32_t vmin = 0xFFFFFFFFu;
#pragma unroll
      for (int n = 0; n < NNODE; ++n) {
        const uint32_t v = e2v[(uint64_t)e * NNODE + n];
        px[n] = xyz + (uint64_t)v * NDIM;
        pu[n] = Model::NEEDS_DISP ? disp + (uint64_t)v * NDIM : nullptr;
        vmin = min(vmin, v);
      }
      const double w = Model::make(srec + (size_t)le * REC, px, pu, args);
      // the element's energy is recorded once: by the tile that holds its lowest-numbered vertex
      if (Model::HAS_GRAD && w_elem && vmin / P.rows == tile) w_elem[e] = w;
    }
    // the first slot's list bounds and first code do not depend on the records: requested before the barrier
    uint32_t pre_cb = 0, pre_ce = 0, pre_code = 0;
    if (tid < n_off + n_rows) {
      const uint64_t slot0 = tid >= n_off ? P.nnz + r0 + (tid - n_off) : (uint64_t)k0 + tid;
      pre_cb = P.nz2ptr[slot0];
      pre_ce = P.nz2ptr[slot0 + 1];
      pre_code = pre_cb < pre_ce ? P.code16[pre_cb] : 0u;
    }
    __syncthreads();
    // ---- phase 2: one thread per slot (off-diagonal slots [k0, k1), then the diagonal slots of rows [r0, r1)): walk the list in
    // ascending element order (= the serial loop's accumulation order, same bits), accumulate in registers
    for (uint32_t s = tid; s < n_off + n_rows; s += T) {
      const bool is_dia = s >= n_off;
      const uint64_t slot = is_dia ? P.nnz + r0 + (s - n_off) : (uint64_t)k0 + s;
      double acc[NN], gacc[N];
#pragma unroll
      for (int q = 0; q < NN; ++q) acc[q] = 0.0;
#pragma unroll
      for (int d = 0; d < N; ++d) gacc[d] = 0.0;
      const bool first = s == tid;
      const uint32_t cb = first ? pre_cb : P.nz2ptr[slot], ce = first ? pre_ce : P.nz2ptr[slot + 1];
      // the code of the NEXT contribution is fetched while this one is evaluated (ncu: the load sat at the head of every
      // iteration's dependent chain, 6 % of the warp samples; -2.6 % on the S10 assembly)
      uint32_t code_next = first ? pre_code : (cb < ce ? P.code16[cb] : 0u);
      for (uint32_t c = cb; c < ce; ++c) {
        const uint32_t code = code_next;
        code_next = c + 1 < ce ? P.code16[c + 1] : 0u;
        const double *rec = srec + (size_t)(code >> 4) * REC;
        const int i = (int)((code >> 2) & 3u), j = (int)(code & 3u);
        double blk[NN];
        Model::block(rec, i, j, args, blk);
#pragma unroll
        for (int q = 0; q < NN; ++q) acc[q] += blk[q];
        if (Model::HAS_GRAD && dw && is_dia) {
          double g[N];
          Model::grad(rec, i, args, g);
#pragma unroll
          for (int d = 0; d < N; ++d) gacc[d] += g[d];
        }
      }
      double *dst = sout + (size_t)s * NN;
#pragma unroll
      for (int q = 0; q < NN; ++q) dst[q] = acc[q];
      if (Model::HAS_GRAD && dw && is_dia) {
#pragma unroll
        for (int d = 0; d < N; ++d) dw[(r0 + (s - n_off)) * N + d] = gacc[d];
      }
    }
    __syncthreads();
    // ---- phase 3: contiguous, coalesced stores
    if (out.packed.base) {
      const uint64_t prt = (uint64_t)out.packed.rpt;   // 8 or 16: the assembly tile (16 rows) is a whole number of packed tiles
      uint32_t j = 0;
      for (uint64_t pt = r0 / prt; pt * prt < r1; ++pt, ++j) {
        const uint64_t pr0 = pt * prt, pr1 = min(pr0 + prt, P.num_rows);
        const uint32_t pk0 = smeta[AT_META + AT_PKT + j], nb = smeta[AT_META + AT_PKT + j + 1] - pk0;
        // [diagonal blocks | the tile's off-diagonal blocks] of packed record pt (the header and the columns precede them)
        double *vals = reinterpret_cast<double *>(out.packed.base + ((uint64_t)smeta[AT_META + j] << 4) + out.packed.hdr + ((nb * 4u + 15u) & ~15u));
        const uint32_t nd = (uint32_t)(pr1 - pr0) * NN;
        const double *sd = sout + (size_t)(n_off + (pr0 - r0)) * NN;
        for (uint32_t q = tid; q < nd; q += T) vals[q] = sd[q];
        const double *sv = sout + (size_t)(pk0 - k0) * NN;
        double *ov = vals + prt * NN;
        for (uint32_t q = tid; q < nb * NN; q += T) ov[q] = sv[q];
      }
    } else {
      double *ov = out.idx2val + (uint64_t)k0 * NN;
      for (uint32_t q = tid; q < n_off * NN; q += T) ov[q] = sout[q];
      double *od = out.row2val + r0 * NN;
      const double *sd = sout + (size_t)n_off * NN;
      for (uint32_t q = tid; q < n_rows * NN; q += T) od[q] = sd[q];
    }
    // (the __syncthreads at the top of the next iteration orders these reads of sout / smeta before the next tile's writes)
  }
}

template <class Model>
static dfb_status launch_tile(dfb_plan *plan, const typename Model::Args &args, const dfb_mesh *mesh, const dfb_vec *disp, dfb_bsr *m,
                              double *dw, double *w_elem) {
  dfb_ctx *c = plan->ctx;
  constexpr int NN = Model::N * Model::N;
  int threads = (int)((plan->tile_max_slots + 31) / 32 * 32);  // one thread per slot of the largest tile
  if (c->assemble_tile_threads > 0) threads = (int)c->assemble_tile_threads;
  if (threads < 64) threads = 64;
  if (threads > AT_MAX_THREADS) threads = AT_MAX_THREADS;
  const size_t smem = ((size_t)plan->tile_max_elem * Model::REC + (size_t)plan->tile_max_slots * NN) * sizeof(double) + (AT_META + 2 * AT_PKT) * sizeof(uint32_t) + 16;
  if (smem > 200 * 1024) return DFB_ERR_UNSUPPORTED;  // caller falls back (no error string: not a failure)
  TileOut out;
  if (dfb_bsr_packed_native(m)) {
    DFB_TRY(dfb_bsr_packed_overwrite(m));
    out.packed = dfb_bsr_packed_view(m);
    out.idx2val = out.row2val = nullptr;
  } else {
    DFB_TRY(dfb_bsr_canon_overwrite(m));
    out.packed.base = nullptr;
    out.packed.tile_off = nullptr;
    out.packed.NN = NN;
    out.packed.has_dia = 1;
    out.packed.rpt = 8;
    out.packed.hdr = 48;
    out.idx2val = m->d_idx2val;
    out.row2val = m->d_row2val;
  }
  TilePlanView P;
  P.nz2ptr = plan->d_nz2ptr;
  P.code16 = plan->d_code16;
  P.tile_elem = plan->d_tile_elem;
  P.tile_meta = plan->d_tile_meta;
  P.row2idx = plan->pat->d_row2idx;
  P.pair_base = plan->pair_base;
  P.rows = plan->tile_rows;
  P.max_elem = plan->tile_max_elem;
  P.max_slots = plan->tile_max_slots;
  P.nnz = plan->pat->nnz;
  P.num_rows = plan->pat->num_blk;
  P.n_tiles = plan->n_atiles;
  DFB_CUDA(cudaFuncSetAttribute(k_assemble_tile<Model>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_assemble_tile<Model>, threads, smem));
  if (per_sm < 1) per_sm = 1;
  uint64_t g = plan->n_atiles;
  if (g > (uint64_t)c->sm_count * per_sm) g = (uint64_t)c->sm_count * per_sm;
  dfb_phase_begin(c, DFB_PH_TILE);
  DFB_LAUNCH(c, k_assemble_tile<Model>, (unsigned)g, threads, smem, args, P, mesh->d_elem2vtx, mesh->d_xyz, disp ? disp->d : nullptr, out, dw,
             w_elem);
  dfb_phase_end(c);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// linear kinds: returns DFB_ERR_UNSUPPORTED (without touching m) when the tile path does not apply
dfb_status dfb_assemble_tile_linear(dfb_plan *plan, int kind, const dfb_mesh *mesh, double p0, double p1, dfb_bsr *m) {
  DFB_TRY(dfb_plan_build_tiles(plan));
  if (plan->tile_state != 1) return DFB_ERR_UNSUPPORTED;
#define LIN_CASE(K)                                    \
  case K: {                                            \
    ModelLinear<K>::Args a;                            \
    a.p0 = p0;                                         \
    a.p1 = p1;                                         \
    return launch_tile<ModelLinear<K>>(plan, a, mesh, nullptr, m, nullptr, nullptr); \
  }
  switch (kind) {
    LIN_CASE(DFB_ELEM_LAPLACE_TRI2)
    LIN_CASE(DFB_ELEM_COT_LAPLACE_TRI3)
    LIN_CASE(DFB_ELEM_SOLID_LINEAR_TRI2)
    LIN_CASE(DFB_ELEM_LAPLACE_TET)
    LIN_CASE(DFB_ELEM_SOLID_LINEAR_TET)
    LIN_CASE(DFB_ELEM_MASS_TET)
    default: return DFB_ERR_UNSUPPORTED;
  }
#undef LIN_CASE
}

// energy models (spring3, neo-Hookean tet): Hessian -> m, gradient -> dw, per-element energies -> w_elem
dfb_status dfb_assemble_tile_energy(dfb_plan *plan, int kind, const dfb_mesh *mesh, const dfb_vec *disp, double p0, double p1, dfb_bsr *m,
                                    double *dw, double *w_elem) {
  DFB_TRY(dfb_plan_build_tiles(plan));
  if (plan->tile_state != 1) return DFB_ERR_UNSUPPORTED;
  if (kind == DFB_ELEM_SPRING3) {
    ModelSpring3::Args a;
    a.stiffness = p0;
    return launch_tile<ModelSpring3>(plan, a, mesh, disp, m, dw, w_elem);
  }
  if (kind == DFB_ELEM_NEOHOOKEAN_TET && plan->ctx->assemble_tile_neo) {
    ModelNeoHookeanTet::Args a;
    a.myu = p0;
    a.lambda = p1;
    return launch_tile<ModelNeoHookeanTet>(plan, a, mesh, disp, m, dw, w_elem);
  }
  return DFB_ERR_UNSUPPORTED;
}
