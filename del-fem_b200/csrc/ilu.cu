// ilu.cu — ILU(0) preconditioner of the reference on the device (compiled --fmad=false so that the factorisation rounds like
// the reference's).  Restates del-fem-cpu/src/sparse_ilu.rs:
//   initialize_ilu0 :28-56   symbolic phase: per-row sorted copy of the pattern, row2idx_dia = first strictly-upper entry
//   copy_value      :87-125  scatter A's off-diagonals through a column map, copy the diagonal blocks
//   decompose       :127-181 IKJ incomplete factorisation; stores D_i^-1 and scales the strictly-upper part of row i by it
//   solve_preconditioning_vec :183-219  forward sweep (L, then multiply by the stored inverse), backward sweep (U)
// The reference runs all of it serially. Here rows are grouped into dependency LEVELS (level_f(i) = 1 + max level_f(k) over the
// strictly-lower columns k of row i; level_b likewise over the strictly-upper columns, from the last row up): rows of one level
// are independent, each row is processed by ONE thread with exactly the reference's loop order, so the factors and the sweeps
// are deterministic and reproduce the serial arithmetic. One launch per level (S10 Kuhn cube: 448 forward + 448 backward
// levels), which makes this preconditioner latency-bound on a GPU — it exists for fidelity with the reference's own PCG
// (ILU(0) is what `solid_linear_tri2.rs:239-256` and `linearsystem::Solver` use), not for speed; Jacobi is the fast path.
#include <algorithm>
#include <map>
#include <set>
#include <vector>

#include "dfb_internal.h"
#include "reduce.cuh"

struct DfbIlu {
  uint32_t *d_r2i = nullptr;      // row pointers of the ILU pattern [n+1] (== A's for ILU(0), larger rows with fill for ILU(k))
  uint64_t nnz = 0;               // entries of the ILU pattern
  uint32_t *d_col = nullptr;      // sorted columns [nnz]
  uint32_t *d_a2ilu = nullptr;    // position of A's k-th entry inside the ILU arrays [nnz of A]
  uint32_t *d_dia = nullptr;      // row2idx_dia [n]
  uint32_t *d_rows_f = nullptr;   // rows ordered by forward level [n]
  uint32_t *d_rows_b = nullptr;   // rows ordered by backward level [n]
  std::vector<uint32_t> lev_f, lev_b;  // level pointers into d_rows_f / d_rows_b (host)
  double *d_val = nullptr;        // factor values [nnz*NN]
  double *d_z = nullptr;          // explicit preconditioned residual for PCG [n*N]
  uint32_t *d_epoch = nullptr;    // [8] grid-barrier counter of the persistent sweeps at [4]
  uint32_t *d_pos = nullptr;      // [n] position of a row in the forward level order (inverse of d_rows_f)
  double *d_yp = nullptr;         // [n*N] the vector the persistent sweeps exchange, stored in forward level order
  struct Sweep {                  // persistent sweep of one direction
    std::vector<uint32_t> lev_rec;   // records of level l: [lev_rec[l], lev_rec[l+1])
    uint2 *d_groups = nullptr;       // per record {first position in the level-ordered row list, rows}
    uint32_t n_rec = 0;
    unsigned char *d_rec = nullptr;  // sweep-ordered factor copy (IluRec), rebuilt by every update
    uint64_t rec_bytes = 0;
    uint2 *d_items = nullptr;        // work items for one launch shape (ilu_items)
    uint32_t n_items = 0, pass = 0;
  } fwd, bwd;
};

template <int N>
__device__ __forceinline__ void mat_mul(double *c, const double *a, const double *b) {  // matn_col_major::mult_mat_col_major
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = s + a[i + N * k] * b[k + N * j];
      c[i + N * j] = s;
    }
}

template <int N>
__device__ __forceinline__ void mat_vec(double *y, const double *a, const double *x) {  // matn_col_major::mult_vec
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s = s + a[i + N * j] * x[j];
    y[i] = s;
  }
}

template <int N>
__device__ __forceinline__ bool mat_inverse(double *inv, const double *a) {  // try_inverse: Gauss-Jordan, partial pivoting
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
      if (f == 0.0) continue;
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

// copy_value: ilu.idx2val[a2ilu[k]] = a.idx2val[k]; ilu.row2val = a.row2val
__global__ void k_ilu_copy(const double *__restrict__ a_val, const double *__restrict__ a_dia, const uint32_t *__restrict__ a2ilu,
                           uint64_t nnz, uint64_t n, int NN, double *__restrict__ val, double *__restrict__ dia) {
  const uint64_t total = (nnz + n) * NN;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t blk = q / NN;
    const int e = (int)(q - blk * NN);
    if (blk < nnz) val[(uint64_t)a2ilu[blk] * NN + e] = a_val[q];
    else dia[(blk - nnz) * NN + e] = a_dia[(blk - nnz) * NN + e];
  }
}

// decompose, one level: rows[first..last)
template <int N>
__global__ void k_ilu_decompose_level(const uint32_t *__restrict__ rows, uint32_t first, uint32_t last, const uint32_t *__restrict__ r2i,
                                      const uint32_t *__restrict__ col, const uint32_t *__restrict__ diap, double *__restrict__ val,
                                      double *__restrict__ dia, int *err) {
  constexpr int NN = N * N;
  for (uint32_t q = first + blockIdx.x * blockDim.x + threadIdx.x; q < last; q += gridDim.x * blockDim.x) {
    const uint32_t i = rows[q];
    const uint32_t rb = r2i[i], re = r2i[i + 1], rd = diap[i];
    for (uint32_t ik = rb; ik < rd; ++ik) {
      const uint32_t k = col[ik];
      double ikv[NN];
#pragma unroll
      for (int e = 0; e < NN; ++e) ikv[e] = val[(uint64_t)ik * NN + e];
      for (uint32_t kj = diap[k]; kj < r2i[k + 1]; ++kj) {
        const uint32_t j = col[kj];
        double ikkj[NN];
        mat_mul<N>(ikkj, ikv, val + (uint64_t)kj * NN);
        if (j != i) {
          // col2idx[j] of the reference: position of column j in (sorted) row i, if present
          uint32_t lo = rb, hi = re;
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (col[mid] < j) lo = mid + 1;
            else hi = mid;
          }
          if (lo == re || col[lo] != j) continue;
#pragma unroll
          for (int e = 0; e < NN; ++e) val[(uint64_t)lo * NN + e] -= ikkj[e];
        } else {
#pragma unroll
          for (int e = 0; e < NN; ++e) dia[(uint64_t)i * NN + e] -= ikkj[e];
        }
      }
    }
    double inv[NN], d[NN];
#pragma unroll
    for (int e = 0; e < NN; ++e) d[e] = dia[(uint64_t)i * NN + e];
    if (!mat_inverse<N>(inv, d)) {
      atomicExch(err, 1);
      continue;
    }
#pragma unroll
    for (int e = 0; e < NN; ++e) dia[(uint64_t)i * NN + e] = inv[e];
    for (uint32_t ij = rd; ij < re; ++ij) {
      double t[NN];
      mat_mul<N>(t, inv, val + (uint64_t)ij * NN);
#pragma unroll
      for (int e = 0; e < NN; ++e) val[(uint64_t)ij * NN + e] = t[e];
    }
  }
}

// forward sweep, one level: t = v_i - sum_{j<i} L_ij v_j ; v_i = D_i^-1 t
template <int N>
__global__ void k_ilu_forward_level(const uint32_t *__restrict__ rows, uint32_t first, uint32_t last, const uint32_t *__restrict__ r2i,
                                    const uint32_t *__restrict__ col, const uint32_t *__restrict__ diap, const double *__restrict__ val,
                                    const double *__restrict__ dia, double *__restrict__ vec) {
  constexpr int NN = N * N;
  for (uint32_t q = first + blockIdx.x * blockDim.x + threadIdx.x; q < last; q += gridDim.x * blockDim.x) {
    const uint32_t i = rows[q];
    double t[N], v[N];
#pragma unroll
    for (int d = 0; d < N; ++d) t[d] = vec[(uint64_t)i * N + d];
    for (uint32_t k = r2i[i]; k < diap[i]; ++k) {
      mat_vec<N>(v, val + (uint64_t)k * NN, vec + (uint64_t)col[k] * N);
#pragma unroll
      for (int d = 0; d < N; ++d) t[d] -= v[d];
    }
    mat_vec<N>(v, dia + (uint64_t)i * NN, t);
#pragma unroll
    for (int d = 0; d < N; ++d) vec[(uint64_t)i * N + d] = v[d];
  }
}

// backward sweep, one level: v_i -= sum_{j>i} U_ij v_j
template <int N>
__global__ void k_ilu_backward_level(const uint32_t *__restrict__ rows, uint32_t first, uint32_t last, const uint32_t *__restrict__ r2i,
                                     const uint32_t *__restrict__ col, const uint32_t *__restrict__ diap, const double *__restrict__ val,
                                     double *__restrict__ vec) {
  constexpr int NN = N * N;
  for (uint32_t q = first + blockIdx.x * blockDim.x + threadIdx.x; q < last; q += gridDim.x * blockDim.x) {
    const uint32_t i = rows[q];
    double t[N], v[N];
#pragma unroll
    for (int d = 0; d < N; ++d) t[d] = vec[(uint64_t)i * N + d];
    for (uint32_t k = diap[i]; k < r2i[i + 1]; ++k) {
      mat_vec<N>(v, val + (uint64_t)k * NN, vec + (uint64_t)col[k] * N);
#pragma unroll
      for (int d = 0; d < N; ++d) t[d] -= v[d];
    }
#pragma unroll
    for (int d = 0; d < N; ++d) vec[(uint64_t)i * N + d] = t[d];
  }
}

// ---- level sweeps inside ONE persistent launch (S10: 2 launches per application instead of 448 + 448): one CTA per SM, a
// grid-wide barrier between levels.
// Sweep-ordered factor copy: after every decompose the rows' strictly-lower (forward) / strictly-upper (backward) entries are
// re-laid in the order the sweep visits them, in records of RPW rows (= the rows one warp handles at a time, cut at level
// boundaries):   [ uint4 {row, entries here, first / end of the entries beyond KS in the canonical arrays} x RPW | position [RPW]
//                | neighbour positions [KS][RPW] | blocks [KS][N][RPW*N] (component a of row r at r*N + a) | D^-1 [N][RPW*N] (forward) ]
// and the vector the levels exchange (yp) is stored in the same FORWARD LEVEL ORDER ("position"): the rows of a record are
// consecutive there, and so — on a structured mesh — are their k-th neighbours in the earlier levels, so a warp's gather of a
// neighbour touches 2-3 memory lines instead of 8-16. The forward sweep reads the right-hand side from the row-ordered vector
// and writes yp; the backward sweep works on yp and writes the result back in row order.
// A warp fetches its record with ONE bulk copy (TMA, evict-first: the factor streams through L2 once per sweep and must not
// push out the vector the levels exchange through L2), two items ahead, into its own double buffer in shared memory. The
// canonical row-major blocks cost 8 memory wavefronts per load instruction — one per row — and the sweeps were bound by
// exactly that (330 + 200 wavefronts per warp and level: ncu + in-kernel clocks, DFB_ILU_TRACE).
// A block-row is shared by N lanes — lane a carries component a of t = v_i - sum_j L_ij v_j through the reference's loop
// (sparse_ilu.rs:183-219: ascending entries, mult_vec's left fold over the block's columns), the neighbours' values and the
// D^-1 product move between the lanes by shuffle — so every rounding is the one of the one-thread-per-row kernels above
// (bit-identical, tested).
// Work items (one pass of the grid over part of a level) are built on the host for the launch shape and staged in shared memory.
// Measured (tools/precond_compare.py, profiles/r2_precond_compare.txt): S1 1.23 ms per PCG iteration (round 1, one launch per
// level: 6.4), S10 4.1 (41.7). What remains per level at S1 (in-kernel clocks): ~1.5 us grid barrier (tools/micro/gridbar.cu) +
// ~0.5 us until the neighbours' values arrive + ~0.3 us arithmetic.
// Built, measured slower and removed: per-row ready flags polled by the dependent rows (sync-free, Liu et al.), also
// with the exchanged values as their own flags on this kernel (3.2 ms at S1); one thread per row with register prefetch of the
// canonical entries (3.8 ms); the whole sweep on ONE 16-CTA thread-block cluster with the hardware cluster barrier (1.4 ms at
// S1, 8.6 ms at S10); tree barriers (group counters -> root -> release flags: 3.7-8.3 us).
template <int N> struct IluLanes {   // lanes per row (N = 3: the 4th lane of a quad idles) and rows per warp
  static constexpr int LPR = (N == 3) ? 4 : N, RPW = 32 / LPR;
};
static const int ILU_KS = 8;                   // entries per row held in the sweep-ordered records
static const uint32_t ILU_NONE = 0xFFFFFFFFu;
template <int N, bool FORWARD> struct IluRec {
  static constexpr int RPW = IluLanes<N>::RPW, RN = RPW * N;
  static constexpr int OFF_POS = RPW * 16;
  static constexpr int OFF_COL = OFF_POS + RPW * 4;
  static constexpr int OFF_W = OFF_COL + ILU_KS * RPW * 4;
  static constexpr int OFF_DIA = OFF_W + ILU_KS * N * RN * 8;
  static constexpr int BYTES = (OFF_DIA + (FORWARD ? N * RN * 8 : 0) + 127) / 128 * 128;
};
static uint64_t ilu_rec_bytes(int N, bool forward) {
  switch (N) {
    case 1: return forward ? IluRec<1, true>::BYTES : IluRec<1, false>::BYTES;
    case 2: return forward ? IluRec<2, true>::BYTES : IluRec<2, false>::BYTES;
    case 3: return forward ? IluRec<3, true>::BYTES : IluRec<3, false>::BYTES;
    default: return forward ? IluRec<4, true>::BYTES : IluRec<4, false>::BYTES;
  }
}

// one warp per record: gathers the rows' entries from the canonical factor arrays
template <int N, bool FORWARD>
__global__ void __launch_bounds__(256) k_ilu_slice(const uint2 *__restrict__ groups, uint32_t n_groups, const uint32_t *__restrict__ rows,
                                                   const uint32_t *__restrict__ r2i, const uint32_t *__restrict__ diap,
                                                   const uint32_t *__restrict__ col, const double *__restrict__ val,
                                                   const double *__restrict__ dia, const uint32_t *__restrict__ pos,
                                                   unsigned char *__restrict__ rec) {
  using R = IluRec<N, FORWARD>;
  constexpr int NN = N * N, LPR = IluLanes<N>::LPR, RPW = R::RPW, RN = R::RN, KS = ILU_KS;
  const unsigned lane = threadIdx.x & 31u, sub = lane % LPR, r = lane / LPR;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = warp; g < n_groups; g += n_warps) {
    const uint2 gr = groups[g];   // {first position in the level-ordered row list, rows}
    unsigned char *p = rec + (uint64_t)g * R::BYTES;
    const bool live = r < gr.y;
    const uint32_t i = live ? rows[gr.x + r] : ILU_NONE;
    uint32_t kb = 0, ke = 0;
    if (live) {
      kb = FORWARD ? r2i[i] : diap[i];
      ke = FORWARD ? diap[i] : r2i[i + 1];
    }
    const uint32_t cnt = min(ke - kb, (uint32_t)KS);
    if (sub == 0) {
      reinterpret_cast<uint4 *>(p)[r] = make_uint4(i, cnt, kb + cnt, ke);
      reinterpret_cast<uint32_t *>(p + R::OFF_POS)[r] = live ? pos[i] : 0u;
    }
    // neighbours are addressed by their POSITION in the forward level order: the sweeps exchange a vector stored in that order
    uint32_t *pc = reinterpret_cast<uint32_t *>(p + R::OFF_COL);
    for (uint32_t k = sub; k < (uint32_t)KS; k += LPR) pc[k * RPW + r] = k < cnt ? pos[col[kb + k]] : 0u;
    if (sub < (unsigned)N) {
      double *pw = reinterpret_cast<double *>(p + R::OFF_W);
      for (uint32_t k = 0; k < (uint32_t)KS; ++k)
#pragma unroll
        for (int d = 0; d < N; ++d) pw[(k * N + d) * RN + r * N + sub] = k < cnt ? val[(uint64_t)(kb + k) * NN + sub + N * d] : 0.0;
      if (FORWARD) {
        double *pd = reinterpret_cast<double *>(p + R::OFF_DIA);
#pragma unroll
        for (int d = 0; d < N; ++d) pd[d * RN + r * N + sub] = live ? dia[(uint64_t)i * NN + sub + N * d] : 0.0;
      }
    }
  }
}

// grid barrier: arrival = red.release (the CTA's stores, ordered before it by bar.sync, are performed first), one polling thread
// per CTA with ld.acquire. tools/micro/gridbar.cu (profiles/r2_gridbar.txt): 1.52 us per barrier incl. one dependent store/load
// pair on 148 CTAs; fence + atomicAdd + relaxed poll + fence 1.87, a flag per CTA polled by a warp 3.0-3.9, grid.sync() 1.47;
// spreading the arrivals over several counters changes nothing.
__device__ __forceinline__ void ilu_grid_barrier(unsigned int *ctr, unsigned int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned int want = epoch * gridDim.x;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < want);
  }
  __syncthreads();
}
__global__ void k_ilu_barrier_reset(unsigned int *ctr) { *ctr = 0u; }

static const int ILU_ITEM_SMEM = 2048;                  // work items staged in shared memory (longer lists: read from global)
static const uint32_t ILU_ITEM_BARRIER = 0x80000000u;   // flag in an item's record count: a level boundary precedes the item

#ifdef DFB_ILU_TRACE   // debugging aid (make EXTRA=-DDFB_ILU_TRACE): SM clocks of CTA 0 / warp 0 around the phases of every item
__device__ long long g_ilu_trace[8 * 1024];
#define ILU_TRACE(slot) if (FORWARD && cta == 0 && threadIdx.x == 0 && k < 1024) g_ilu_trace[k * 8 + (slot)] = clock64()
#else
#define ILU_TRACE(slot)
#endif

template <int N, bool FORWARD> struct IluSweepSmem {   // dynamic shared memory per warp: two record buffers + their mbarriers
  static constexpr int WARP_BYTES = 2 * IluRec<N, FORWARD>::BYTES + 128;
};

template <int N, bool FORWARD, int THREADS>
__global__ void __launch_bounds__(THREADS) k_ilu_sweep(const unsigned char *__restrict__ rec, const uint2 *__restrict__ items, uint32_t n_items,
                                                       const uint32_t *__restrict__ col, const double *__restrict__ val,
                                                       const uint32_t *__restrict__ pos, double *vec, double *yp, unsigned int *bar) {
  using R = IluRec<N, FORWARD>;
  constexpr int NN = N * N, LPR = IluLanes<N>::LPR, RPW = R::RPW, RN = R::RN, KS = ILU_KS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint2 s_item[ILU_ITEM_SMEM];
  const unsigned n_cta = gridDim.x, cta = blockIdx.x;
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const unsigned sub = lane % LPR, base = lane - sub, r = lane / LPR;
  const bool act = sub < (unsigned)N;
  const unsigned a_ = act ? sub : 0u;           // component of the row this lane carries
  const uint32_t wslot = wid * n_cta + cta;     // consecutive records of a level go to different CTAs
  unsigned char *wbase = smem_raw + (size_t)wid * IluSweepSmem<N, FORWARD>::WARP_BYTES;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(wbase + 2 * R::BYTES);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the factor records stream through L2 once per sweep: evict-first, so that they do not push out the vector the levels
  // exchange through L2 (its lines are touched at one level only and would be gone by the time they are needed)
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const bool in_smem = n_items <= (uint32_t)ILU_ITEM_SMEM;
  if (in_smem)
    for (uint32_t k = threadIdx.x; k < n_items; k += THREADS) s_item[k] = items[k];
  __syncthreads();
  auto item = [&](uint32_t k) -> uint2 {   // {first record, records | barrier flag}
    if (k >= n_items) return make_uint2(0u, 0u);
    return in_smem ? s_item[k] : items[k];
  };
  auto record = [&](uint32_t k) -> const unsigned char * {   // this warp's record of item k (warp-uniform), or null
    const uint2 it = item(k);
    return wslot < (it.y & ~ILU_ITEM_BARRIER) ? rec + (uint64_t)(it.x + wslot) * R::BYTES : nullptr;
  };
  // the record of item k into buffer b: ONE bulk copy (TMA) per warp, no registers and no load-pipe slots spent on the stream
  auto issue = [&](uint32_t k, uint32_t b) {
    const unsigned char *p = record(k);
    if (lane != 0 || !p) return;
    const uint32_t mb = bar0 + 8 * b;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)R::BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(wbase + (size_t)b * R::BYTES)),
                 "l"(p), "r"((uint32_t)R::BYTES), "r"(mb), "l"(pol)
                 : "memory");
  };
  // the rows of one record (in shared memory); executed by whole warps (shuffles)
  auto do_rows = [&](const unsigned char *buf, uint32_t k) {
    (void)k;
    const uint4 h = reinterpret_cast<const uint4 *>(buf)[r];   // {row, entries here, first / end of the entries beyond KS}
    const bool live = h.x != ILU_NONE;
    const uint32_t i = live ? h.x : 0u, cnt = h.y;
    const uint32_t own = reinterpret_cast<const uint32_t *>(buf + R::OFF_POS)[r];   // the row's position in the forward level order
    const uint32_t *pc = reinterpret_cast<const uint32_t *>(buf + R::OFF_COL);
    // forward: the right-hand side comes from vec (row order); backward: the forward result from yp (level order)
    double t = !live ? 0.0 : FORWARD ? __ldcg(vec + (uint64_t)i * N + a_) : __ldcg(yp + (uint64_t)own * N + a_);
    double xv[KS];
#pragma unroll
    for (int q = 0; q < KS; ++q)   // lane a reads component a of the neighbour (written by other SMs in earlier levels: L2)
      xv[q] = ((uint32_t)q < cnt && act) ? __ldcg(yp + (uint64_t)pc[q * RPW + r] * N + a_) : 0.0;
    ILU_TRACE(1);
    const double *pw = reinterpret_cast<const double *>(buf + R::OFF_W);
#pragma unroll
    for (int q = 0; q < KS; ++q) {
      double acc = 0.0;
#pragma unroll
      for (int d = 0; d < N; ++d) acc = acc + pw[(q * N + d) * RN + r * N + a_] * __shfl_sync(0xFFFFFFFFu, xv[q], base + d);   // the row's lanes exchange
      if ((uint32_t)q < cnt) t -= acc;
    }
    for (uint32_t q = h.z; q < h.w; ++q) {   // rows longer than the record (ILU(k) fill): the canonical arrays
      const uint32_t c = pos[col[q]];
      double acc = 0.0;
#pragma unroll
      for (int d = 0; d < N; ++d) acc = acc + val[(uint64_t)q * NN + a_ + N * d] * __ldcg(yp + (uint64_t)c * N + d);
      t -= acc;
    }
    __syncwarp();
    if (FORWARD) {
      const double *pd = reinterpret_cast<const double *>(buf + R::OFF_DIA);
      double acc = 0.0;
#pragma unroll
      for (int d = 0; d < N; ++d) acc = acc + pd[d * RN + r * N + a_] * __shfl_sync(0xFFFFFFFFu, t, base + d);
      t = acc;
    }
    if (live && act) {
      yp[(uint64_t)own * N + a_] = t;
      if (!FORWARD) vec[(uint64_t)i * N + a_] = t;   // the result, back in row order
    }
#ifdef DFB_ILU_TRACE
    if (t == 1.2345e300) g_ilu_trace[8191] = 1;   // the clock read behind the rows must wait for their result
#endif
  };

  if (n_items == 0) return;
  __syncwarp();
  issue(0, 0);
  issue(1, 1);
  uint32_t phase[2] = {0u, 0u};
  unsigned int n_bar = 0;
  for (uint32_t k = 0;; ++k) {
    const uint32_t b = k & 1u;
    ILU_TRACE(0);
    if (record(k)) {
      const uint32_t mb = bar0 + 8 * b;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "ILU_WAIT:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
          "@p bra ILU_DONE;\n"
          "bra ILU_WAIT;\n"
          "ILU_DONE:\n"
          "}\n" ::"r"(mb),
          "r"(phase[b])
          : "memory");
      phase[b] ^= 1u;
      ILU_TRACE(4);
      do_rows(wbase + (size_t)b * R::BYTES, k);
      __syncwarp();   // every lane has read the buffer before the next copy into it is issued
    }
    ILU_TRACE(2);
    issue(k + 2, b);
    if (k + 1 >= n_items) break;
    if (item(k + 1).y & ILU_ITEM_BARRIER) ilu_grid_barrier(bar, ++n_bar);
    ILU_TRACE(3);
  }
}

// red_out[0] = a . b  (deterministic)
__global__ void __launch_bounds__(256) k_ilu_dot(const double *__restrict__ a, const double *__restrict__ b, uint64_t n, double *partials,
                                                 unsigned int *ticket, double *out) {
  __shared__ double s_buf[32];
  double acc[1] = {0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc[0] += a[i] * b[i];
  if (grid_sum_last_block<1>(acc, partials, ticket, s_buf) && threadIdx.x == 0) out[0] = acc[0];
}

// ------------------------------------------------------------------------------------------------ host side
void dfb_ilu_destroy(DfbIlu *s) {
  if (!s) return;
  dfb_dev_free(s->d_r2i);
  dfb_dev_free(s->d_col);
  dfb_dev_free(s->d_a2ilu);
  dfb_dev_free(s->d_dia);
  dfb_dev_free(s->d_rows_f);
  dfb_dev_free(s->d_rows_b);
  dfb_dev_free(s->d_val);
  dfb_dev_free(s->d_z);
  dfb_dev_free(s->d_epoch);
  dfb_dev_free(s->d_pos);
  dfb_dev_free(s->d_yp);
  for (DfbIlu::Sweep *w : {&s->fwd, &s->bwd}) {
    dfb_dev_free(w->d_groups);
    dfb_dev_free(w->d_rec);
    dfb_dev_free(w->d_items);
  }
  delete s;
}

// initialize_ilu0 (:28-56) / initialize_iluk (:58-64, symbolic_iluk :221-312) + level schedule
dfb_status dfb_ilu_create(dfb_ctx *c, const dfb_bsr *m, uint64_t lev_fill, DfbIlu **out) {
  DFB_REQUIRE(m->pat->convention == 0, "ilu: needs the csrdia convention (separate diagonal)");
  DFB_REQUIRE(c->nranks <= 1, "ilu: single-GPU only (a distributed ILU would be a different preconditioner)");
  const uint64_t n = m->pat->num_blk, nnz_a = m->pat->nnz;
  const int N = m->blk_dim, NN = N * N;
  std::vector<uint32_t> a_r2i(n + 1), a_col(nnz_a);
  DFB_CUDA(cudaMemcpyAsync(a_r2i.data(), m->pat->d_row2idx, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (nnz_a) DFB_CUDA(cudaMemcpyAsync(a_col.data(), m->pat->d_idx2col, nnz_a * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  for (uint64_t i = 0; i < n; ++i)
    for (uint32_t k = a_r2i[i]; k < a_r2i[i + 1]; ++k)
      if (a_col[k] == i) return dfb_fail(DFB_ERR_BAD_ARG, "ilu: the pattern must not contain the diagonal (assert!(j_col < i_row) sparse_ilu.rs:199)");
  // ---- symbolic phase: r2i / scol (sorted columns per row) / dia (first strictly-upper entry)
  std::vector<uint32_t> r2i(n + 1, 0), scol, dia(n);
  if (lev_fill == 0) {
    r2i = a_r2i;
    scol = a_col;
    for (uint64_t i = 0; i < n; ++i) std::sort(scol.begin() + r2i[i], scol.begin() + r2i[i + 1]);
  } else {
    // level of fill: row i = A's columns at level 0, then for each strictly-lower column k (ascending) the strictly-upper part
    // of the finished row k, entries (k,j) with level <= lev_fill - 1 reached through (i,k) with level <= lev_fill - 1, at level
    // max + 1 (minimum over paths). Worklist and row are ordered containers, as the reference's BTreeSet / BTreeMap.
    std::vector<uint32_t> lev;
    scol.reserve(nnz_a * 2);
    lev.reserve(nnz_a * 2);
    std::map<uint32_t, uint64_t> row;
    std::set<uint32_t> todo;
    for (uint64_t i = 0; i < n; ++i) {
      row.clear();
      todo.clear();
      for (uint32_t k = a_r2i[i]; k < a_r2i[i + 1]; ++k) {
        row[a_col[k]] = 0;
        if (a_col[k] < i) todo.insert(a_col[k]);
      }
      while (!todo.empty()) {
        const uint32_t k = *todo.begin();
        todo.erase(todo.begin());
        const uint64_t lik = row[k];
        if (lik + 1 > lev_fill) continue;
        for (uint32_t kj = dia[k]; kj < r2i[k + 1]; ++kj) {
          const uint64_t lkj = lev[kj];
          if (lkj + 1 > lev_fill) continue;
          const uint32_t j = scol[kj];
          if (j == i) continue;
          const uint64_t l = std::max(lik, lkj) + 1;
          if (j < i) todo.insert(j);
          auto it = row.find(j);
          if (it == row.end()) row[j] = l;
          else if (l < it->second) it->second = l;
        }
      }
      if (scol.size() + row.size() > 0xFFFFFFF0ull) return dfb_fail(DFB_ERR_INDEX_OVERFLOW, "iluk: fill exceeds the u32 index range");
      for (const auto &cl : row) {
        scol.push_back(cl.first);
        lev.push_back((uint32_t)std::min<uint64_t>(cl.second, 0xFFFFFFFFull));
      }
      r2i[i + 1] = (uint32_t)scol.size();
      dia[i] = r2i[i + 1];
      for (uint32_t q = r2i[i]; q < r2i[i + 1]; ++q)
        if (scol[q] > i) {
          dia[i] = q;
          break;
        }
    }
  }
  const uint64_t nnz = scol.size();
  std::vector<uint32_t> a2ilu(nnz_a), lf(n, 0), lb(n, 0);
  for (uint64_t i = 0; i < n; ++i) {
    const auto rb = scol.begin() + r2i[i], re = scol.begin() + r2i[i + 1];
    if (lev_fill == 0) {
      dia[i] = r2i[i + 1];
      for (uint32_t q = r2i[i]; q < r2i[i + 1]; ++q)
        if (scol[q] > i) {
          dia[i] = q;
          break;
        }
    }
    for (uint32_t k = a_r2i[i]; k < a_r2i[i + 1]; ++k) {
      const auto it = std::lower_bound(rb, re, a_col[k]);  // copy_value's col2idx: A's entry must exist in the ILU row
      if (it == re || *it != a_col[k]) return dfb_fail(DFB_ERR_PATTERN_MISS, "ilu copy_value: entry outside the ILU pattern (panic!() sparse_ilu.rs:115)");
      a2ilu[k] = (uint32_t)(it - scol.begin());
    }
  }
  uint32_t max_f = 0, max_b = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t l = 0;
    for (uint32_t k = r2i[i]; k < dia[i]; ++k) l = std::max(l, lf[scol[k]] + 1);
    lf[i] = l;
    max_f = std::max(max_f, l);
  }
  for (uint64_t ii = n; ii-- > 0;) {
    uint32_t l = 0;
    for (uint32_t k = dia[ii]; k < r2i[ii + 1]; ++k) l = std::max(l, lb[scol[k]] + 1);
    lb[ii] = l;
    max_b = std::max(max_b, l);
  }
  auto order = [&](const std::vector<uint32_t> &lev, uint32_t nlev, std::vector<uint32_t> &ptr, std::vector<uint32_t> &rows) {
    ptr.assign(nlev + 2, 0);
    for (uint64_t i = 0; i < n; ++i) ptr[lev[i] + 1] += 1;
    for (uint32_t l = 0; l <= nlev; ++l) ptr[l + 1] += ptr[l];
    rows.resize(n);
    std::vector<uint32_t> cur(ptr.begin(), ptr.end() - 1);
    for (uint64_t i = 0; i < n; ++i) rows[cur[lev[i]]++] = (uint32_t)i;  // ascending row index inside a level
  };
  DfbIlu *s = new DfbIlu();
  std::vector<uint32_t> rows_f, rows_b;
  order(lf, max_f, s->lev_f, rows_f);
  order(lb, max_b, s->lev_b, rows_b);
  s->lev_f.resize(n ? max_f + 2 : 1);
  s->lev_b.resize(n ? max_b + 2 : 1);
  dfb_status st = DFB_OK;
  auto up = [&](uint32_t **d, const std::vector<uint32_t> &h) {
    if (st != DFB_OK) return;
    st = dfb_dev_alloc_t(d, h.size() + 2);
    if (st == DFB_OK && !h.empty()) {
      cudaError_t e = cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "ilu: %s", cudaGetErrorString(e));
    }
  };
  s->nnz = nnz;
  up(&s->d_r2i, r2i);
  up(&s->d_col, scol);
  up(&s->d_a2ilu, a2ilu);
  up(&s->d_dia, dia);
  up(&s->d_rows_f, rows_f);
  up(&s->d_rows_b, rows_b);
  // records of the sweep-ordered factor copy: RPW consecutive rows of a level
  auto records = [&](const std::vector<uint32_t> &lev, bool forward, DfbIlu::Sweep &w) {
    if (st != DFB_OK) return;
    const uint32_t rpw = N == 1 ? 32u : N == 2 ? 16u : 8u;
    std::vector<uint2> g;
    w.lev_rec.assign(1, 0u);
    for (size_t l = 0; l + 1 < lev.size(); ++l) {
      for (uint32_t q = lev[l]; q < lev[l + 1]; q += rpw) g.push_back(make_uint2(q, std::min(rpw, lev[l + 1] - q)));
      w.lev_rec.push_back((uint32_t)g.size());
    }
    w.n_rec = (uint32_t)g.size();
    w.rec_bytes = (uint64_t)g.size() * ilu_rec_bytes(N, forward);
    st = dfb_dev_alloc_t(&w.d_groups, g.size() + 1);
    if (st == DFB_OK && !g.empty()) {
      cudaError_t e = cudaMemcpy(w.d_groups, g.data(), g.size() * sizeof(uint2), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "ilu: %s", cudaGetErrorString(e));
    }
    if (st == DFB_OK) st = dfb_dev_alloc_t(&w.d_rec, w.rec_bytes + 128);
  };
  records(s->lev_f, true, s->fwd);
  records(s->lev_b, false, s->bwd);
  {
    std::vector<uint32_t> pos(n);
    for (uint64_t q = 0; q < n; ++q) pos[rows_f[q]] = (uint32_t)q;
    up(&s->d_pos, pos);
    if (st == DFB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "ilu0: upload failed");   // pos is a local
    if (st == DFB_OK) st = dfb_dev_alloc_t(&s->d_yp, n * N + 8);
  }
  if (st == DFB_OK) st = dfb_dev_alloc_t(&s->d_epoch, 8);
  if (st == DFB_OK) cudaMemsetAsync(s->d_epoch, 0, 8 * sizeof(uint32_t), c->stream);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&s->d_val, nnz * NN + 8);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&s->d_z, n * N + 8);
  if (st == DFB_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "ilu0: upload failed");
  if (st != DFB_OK) {
    dfb_ilu_destroy(s);
    return st;
  }
  *out = s;
  return DFB_OK;
}

static int level_grid(uint32_t count) {
  int g = (int)((count + 127) / 128);
  return g < 1 ? 1 : (g > 1184 ? 1184 : g);
}

static dfb_status dfb_ilu_reslice(dfb_ctx *c, DfbIlu *s, int N, const double *d_dia);

// copy_value + decompose
dfb_status dfb_ilu_update(dfb_ctx *c, DfbIlu *s, const dfb_bsr *m, double *d_dia_out) {
  const uint64_t n = m->pat->num_blk, nnz = m->pat->nnz;
  const int N = m->blk_dim, NN = N * N;
  if (s->nnz != nnz) DFB_CUDA(cudaMemsetAsync(s->d_val, 0, s->nnz * NN * sizeof(double), c->stream));  // fill entries start at zero
  DFB_TRY(dfb_bsr_canon_read(m));
  DFB_LAUNCH(c, k_ilu_copy, c->sm_count * 8, 256, 0, m->d_idx2val, m->d_row2val, s->d_a2ilu, nnz, n, NN, s->d_val, d_dia_out);
  const uint32_t nlev = s->lev_f.empty() ? 0 : (uint32_t)s->lev_f.size() - 1;
  for (uint32_t l = 0; l < nlev; ++l) {
    const uint32_t a = s->lev_f[l], b = s->lev_f[l + 1];
    if (a == b) continue;
#define DEC(NV)                                                                                                                     \
  case NV:                                                                                                                          \
    DFB_LAUNCH(c, k_ilu_decompose_level<NV>, level_grid(b - a), 128, 0, s->d_rows_f, a, b, s->d_r2i, s->d_col, s->d_dia, s->d_val, \
               d_dia_out, c->d_errflag + 2);                                                                                        \
    break;
    switch (N) {
      DEC(1) DEC(2) DEC(3) DEC(4)
    }
#undef DEC
  }
  DFB_CHECK_LAUNCH();
  DFB_TRY(dfb_check_error_flag(c, 2, DFB_ERR_SINGULAR_DIAG, "ilu0 decompose: singular pivot block (try_inverse().unwrap())"));
  return dfb_ilu_reslice(c, s, N, d_dia_out);
}

// work items of a sweep for a launch shape: the records of every non-empty level are cut into passes of `pass` records (one per
// warp of the grid); the first pass of every level but the first carries the barrier flag. Cached per direction.
static dfb_status ilu_items(dfb_ctx *c, const DfbIlu *cs, bool forward, uint32_t pass, const uint2 **d_items, uint32_t *n_items) {
  DfbIlu *s = const_cast<DfbIlu *>(cs);
  DfbIlu::Sweep &W = forward ? s->fwd : s->bwd;
  if (W.pass != pass || !W.d_items) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    DFB_CUDA(cudaStreamIsCapturing(c->stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) return dfb_fail(DFB_ERR_BAD_ARG, "ilu: the sweep shape changed inside a stream capture");
    std::vector<uint2> h;
    bool first = true;
    for (size_t l = 0; l + 1 < W.lev_rec.size(); ++l) {
      const uint32_t a = W.lev_rec[l], b = W.lev_rec[l + 1];
      for (uint32_t g = a; g < b; g += pass) {
        h.push_back(make_uint2(g, std::min(pass, b - g) | ((g == a && !first) ? ILU_ITEM_BARRIER : 0u)));
        first = false;
      }
    }
    DFB_CUDA(cudaStreamSynchronize(c->stream));   // a sweep with the old list may be in flight
    dfb_dev_free(W.d_items);
    W.d_items = nullptr;
    DFB_TRY(dfb_dev_alloc_t(&W.d_items, h.size() + 1));
    if (!h.empty()) DFB_CUDA(cudaMemcpy(W.d_items, h.data(), h.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    W.n_items = (uint32_t)h.size();
    W.pass = pass;
  }
  *d_items = W.d_items;
  *n_items = W.n_items;
  return DFB_OK;
}

// the sweep-ordered records of both directions from the canonical factor (after decompose)
template <int N>
static dfb_status ilu_slice(dfb_ctx *c, DfbIlu *s, const double *d_dia) {
  const int g = c->sm_count * 8;
  if (s->fwd.n_rec)
    DFB_LAUNCH(c, (k_ilu_slice<N, true>), g, 256, 0, s->fwd.d_groups, s->fwd.n_rec, s->d_rows_f, s->d_r2i, s->d_dia, s->d_col, s->d_val, d_dia, s->d_pos, s->fwd.d_rec);
  if (s->bwd.n_rec)
    DFB_LAUNCH(c, (k_ilu_slice<N, false>), g, 256, 0, s->bwd.d_groups, s->bwd.n_rec, s->d_rows_b, s->d_r2i, s->d_dia, s->d_col, s->d_val, d_dia, s->d_pos, s->bwd.d_rec);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

static dfb_status dfb_ilu_reslice(dfb_ctx *c, DfbIlu *s, int N, const double *d_dia) {
  switch (N) {
    case 1: return ilu_slice<1>(c, s, d_dia);
    case 2: return ilu_slice<2>(c, s, d_dia);
    case 3: return ilu_slice<3>(c, s, d_dia);
    case 4: return ilu_slice<4>(c, s, d_dia);
  }
  return dfb_fail(DFB_ERR_UNSUPPORTED, "ilu: blk_dim %d", N);
}

// solve_preconditioning_vec, in place: forward then backward sweep, one launch each
template <int N>
static dfb_status ilu_sweeps(dfb_ctx *c, const DfbIlu *s, double *d_vec) {
  constexpr int THREADS = N == 4 ? 256 : 512;   // two record buffers per warp: 180 KB of shared memory (N = 4: 9.6 KB records)
  const int n_cta = c->sm_count;                // one resident CTA per SM: the grid barrier needs all of them running
  const uint32_t pass = (uint32_t)n_cta * (THREADS / 32);
  const uint2 *it_f = nullptr, *it_b = nullptr;
  uint32_t nf = 0, nb = 0;
  DFB_TRY(ilu_items(c, s, true, pass, &it_f, &nf));
  DFB_TRY(ilu_items(c, s, false, pass, &it_b, &nb));
  const unsigned char *rec_f = s->fwd.d_rec, *rec_b = s->bwd.d_rec;
  const uint32_t *col = s->d_col;
  const double *val = s->d_val;
  unsigned int *bar = s->d_epoch + 4;
  const size_t smem_f = (size_t)(THREADS / 32) * IluSweepSmem<N, true>::WARP_BYTES, smem_b = (size_t)(THREADS / 32) * IluSweepSmem<N, false>::WARP_BYTES;
  // (the dynamic shared-memory opt-in is a per-device function attribute: set on every launch, cheap)
  DFB_CUDA(cudaFuncSetAttribute(k_ilu_sweep<N, true, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
  DFB_CUDA(cudaFuncSetAttribute(k_ilu_sweep<N, false, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  DFB_LAUNCH(c, k_ilu_barrier_reset, 1, 1, 0, bar);
  if (nf) DFB_LAUNCH(c, (k_ilu_sweep<N, true, THREADS>), n_cta, THREADS, smem_f, rec_f, it_f, nf, col, val, s->d_pos, d_vec, s->d_yp, bar);
  DFB_LAUNCH(c, k_ilu_barrier_reset, 1, 1, 0, bar);
  if (nb) DFB_LAUNCH(c, (k_ilu_sweep<N, false, THREADS>), n_cta, THREADS, smem_b, rec_b, it_b, nb, col, val, s->d_pos, d_vec, s->d_yp, bar);
  DFB_CHECK_LAUNCH();
#ifdef DFB_ILU_TRACE
  {
    static int calls = 0;
    if (++calls == 5) {
      cudaStreamSynchronize(c->stream);
      static long long h[8 * 1024];
      cudaMemcpyFromSymbol(h, g_ilu_trace, sizeof(h));
      const int lo = (int)nf / 4, hi = std::min((int)nf - 1, 1023) * 3 / 4;
      double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = lo; k < hi; ++k) {
        const long long *q = h + k * 8;
        d[0] += q[4] - q[0];   // wait for the record (TMA)
        d[1] += q[1] - q[4];   // header + gather loads issued
        d[2] += q[2] - q[1];   // wait for the values + arithmetic + store
        d[3] += q[3] - q[2];   // next copy issued + barrier
        d[4] += q[8] - q[3];   // loop
      }
      const double n = hi - lo;
      fprintf(stderr, "[ilu trace] forward items %d..%d, SM cycles: record wait %.0f | gather issue %.0f | values + arithmetic + store %.0f | barrier %.0f | loop %.0f\n",
              lo, hi, d[0] / n, d[1] / n, d[2] / n, d[3] / n, d[4] / n);
    }
  }
#endif
  return DFB_OK;
}

dfb_status dfb_ilu_apply(dfb_ctx *c, const DfbIlu *s, const dfb_pattern *pat, int N, const double *d_dia, double *d_vec) {
  if (c->ilu_level_launches == 0) {
    switch (N) {
      case 1: return ilu_sweeps<1>(c, s, d_vec);
      case 2: return ilu_sweeps<2>(c, s, d_vec);
      case 3: return ilu_sweeps<3>(c, s, d_vec);
      case 4: return ilu_sweeps<4>(c, s, d_vec);
    }
  }
  // one launch per level (option "ilu_level_launches" = 1): the reference point the sweeps are checked against
  const uint32_t nf = s->lev_f.empty() ? 0 : (uint32_t)s->lev_f.size() - 1, nb = s->lev_b.empty() ? 0 : (uint32_t)s->lev_b.size() - 1;
  for (uint32_t l = 0; l < nf; ++l) {
    const uint32_t a = s->lev_f[l], b = s->lev_f[l + 1];
    if (a == b) continue;
#define FWD(NV)                                                                                                                      \
  case NV:                                                                                                                           \
    DFB_LAUNCH(c, k_ilu_forward_level<NV>, level_grid(b - a), 128, 0, s->d_rows_f, a, b, s->d_r2i, s->d_col, s->d_dia, s->d_val, d_dia, \
               d_vec);                                                                                                               \
    break;
    switch (N) {
      FWD(1) FWD(2) FWD(3) FWD(4)
    }
#undef FWD
  }
  for (uint32_t l = 0; l < nb; ++l) {
    const uint32_t a = s->lev_b[l], b = s->lev_b[l + 1];
    if (a == b) continue;
#define BWD(NV)                                                                                                                       \
  case NV:                                                                                                                            \
    DFB_LAUNCH(c, k_ilu_backward_level<NV>, level_grid(b - a), 128, 0, s->d_rows_b, a, b, s->d_r2i, s->d_col, s->d_dia, s->d_val, d_vec); \
    break;
    switch (N) {
      BWD(1) BWD(2) BWD(3) BWD(4)
    }
#undef BWD
  }
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

double *dfb_ilu_z(DfbIlu *s) { return s->d_z; }

// factor values of src -> dst (both built for the same pattern and level of fill): Solver::clone copies the decomposed ILU
dfb_status dfb_ilu_copy_values(dfb_ctx *c, DfbIlu *dst, const DfbIlu *src, int NN) {
  DFB_REQUIRE(dst && src && dst->nnz == src->nnz, "ilu copy: the two factorisations have different patterns");
  DFB_CUDA(cudaMemcpyAsync(dst->d_val, src->d_val, src->nnz * (uint64_t)NN * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  DFB_REQUIRE(dst->fwd.rec_bytes == src->fwd.rec_bytes && dst->bwd.rec_bytes == src->bwd.rec_bytes, "ilu copy: different sweep records");
  if (src->fwd.rec_bytes) DFB_CUDA(cudaMemcpyAsync(dst->fwd.d_rec, src->fwd.d_rec, src->fwd.rec_bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (src->bwd.rec_bytes) DFB_CUDA(cudaMemcpyAsync(dst->bwd.d_rec, src->bwd.d_rec, src->bwd.rec_bytes, cudaMemcpyDeviceToDevice, c->stream));
  return DFB_OK;
}

dfb_status dfb_ilu_dot(dfb_ctx *c, const double *a, const double *b, uint64_t n, double *d_out) {
  uint64_t g = (n + 1023) / 1024;
  if (g < 1) g = 1;
  if (g > 1184) g = 1184;
  DFB_LAUNCH(c, k_ilu_dot, (unsigned)g, 256, 0, a, b, n, c->d_partials, c->d_ticket + 4, d_out);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}
