// dfb_internal.h — shared internals of libdelfem_b200.so (not part of the public ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/delfem_b200.h"
#include "peer.cuh"

#define DFB_API extern "C" __attribute__((visibility("default")))

void dfb_set_error(const char *fmt, ...);
dfb_status dfb_fail(dfb_status st, const char *fmt, ...);

#define DFB_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return dfb_fail(DFB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DFB_TRY(expr)                 \
  do {                                \
    dfb_status _s = (expr);           \
    if (_s != DFB_OK) return _s;      \
  } while (0)

#define DFB_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return dfb_fail(DFB_ERR_BAD_ARG, __VA_ARGS__);  \
  } while (0)

// kernel launch bookkeeping: every product kernel launch goes through this so gpu_launches is a real count
#define DFB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                  \
  do {                                                                                   \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                     \
    (ctx)->launches += 1;                                                                \
  } while (0)
#define DFB_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                         \
  do {                                                                                   \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                            \
    (ctx)->launches += 1;                                                                \
  } while (0)
#define DFB_CHECK_LAUNCH() DFB_CUDA(cudaGetLastError())

static const int DFB_RED_MAX_BLOCKS = 2048;  // upper bound of any reducing grid
static const int DFB_RED_K = 4;              // max scalars reduced by one kernel

// device-side solver state (lives in ctx->d_state).  The iteration kernels take NO per-iteration host arguments: the iteration
// index, its parity and the exchange ids of the peer all-reduces / halo pushes are derived from this struct on the device, so
// that a chunk of iterations is a CUDA graph of identical nodes that can be replayed.  Two-slot fields are indexed by iteration
// parity so that the slot a kernel reads is never the slot a thread of the same kernel writes; for the same reason `iter` is
// written by the direction kernel only and `iter_dir` (what the direction kernel reads) by the update kernel only.
struct DfbSolverState {
  double rho[2];      // r.z (PCG) / r.r (CG)
  double red[4];      // [0] p.Ap   [1] r.r   [2] r.z   (targets of the block reductions / all-reduce)
  double inv_rr0;     // 1 / (r0.r0)
  double rr0;
  double tol;
  double eps_sq;
  unsigned long long xbase, hbase;          // exchange-id bases of the running solve (scalar mailboxes / halo flags)
  unsigned long long xseq_next, hseq_next;  // first unused ids; persist across solves (identical on every rank: SPMD)
  unsigned long long gflag[2];              // local "total is in red[]" flags of the folded all-reduce: [0] p.Ap, [1] r.r/r.z
  int done[2];        // stop flag (parity slots): converged, max_it reached, or error
  int iter;           // completed iterations
  int iter_dir;
  int err;            // 1: p.Ap < 0 (ls flavour assert)   2: peer exchange timed out
  int check_pap;
  int converged;      // the stop was the convergence test (not max_it)
  int max_it;
};
// exchange ids: init posts xbase+1; iteration i posts p.Ap as xbase+2+2i and {r.r, r.z} as xbase+3+2i; the halo of the direction
// that iteration i multiplies is push hbase+1+i (landing-zone parity buffer = id & 1)
__host__ __device__ inline unsigned long long dfb_xid_pq(unsigned long long xbase, int it) { return xbase + 2ull + 2ull * (unsigned long long)it; }
__host__ __device__ inline unsigned long long dfb_xid_rr(unsigned long long xbase, int it) { return xbase + 3ull + 2ull * (unsigned long long)it; }
__host__ __device__ inline unsigned long long dfb_hid(unsigned long long hbase, int it) { return hbase + 1ull + (unsigned long long)it; }

// phases of the assembly / setup path that the library can time itself (CUDA events around its own launches)
enum { DFB_PH_ELEM = 0,      // element kernels (two-phase path: dfb_emat_batch / node records)
       DFB_PH_GATHER = 1,    // merge: gather of element matrices / node records into the matrix values
       DFB_PH_GRAD = 2,      // gradient gather + energy sum of the energy models
       DFB_PH_PACK = 3,      // (re)build of the tile-packed SpMV mirror
       DFB_PH_TILE = 4,      // fused row-tile assembly (element kernel + merge in one launch)
       DFB_PH_BC = 5,        // Dirichlet rows/columns
       DFB_PH_PRECOND = 6,   // preconditioner update
       DFB_N_PHASES = 8 };
struct dfb_ctx;
void dfb_phase_begin(dfb_ctx *c, int id, cudaStream_t strm = nullptr);
void dfb_phase_end(dfb_ctx *c, cudaStream_t strm = nullptr);

struct dfb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;   // main stream
  cudaStream_t stream_comm = nullptr;  // halo / boundary overlap stream
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
  cudaEvent_t ev_flag[2] = {nullptr, nullptr};
  uint64_t launches = 0;
  int sm_count = 148;
  int64_t l2_bytes = 0;
  int64_t total_mem = 0;
  // reductions
  double *d_partials = nullptr;   // [DFB_RED_MAX_BLOCKS * DFB_RED_K]
  unsigned int *d_ticket = nullptr;  // [8]
  DfbSolverState *d_state = nullptr;
  double *d_hist = nullptr;       // ring of the last hist_cap values of ||r_k||^2 (entry k at k % hist_cap), drained per chunk
  uint64_t hist_cap = 0;
  void *h_poll = nullptr;         // pinned: 2 slots of [DfbSolverState | ring snapshot]
  size_t h_poll_slot = 0;         // bytes per slot
  std::vector<double> last_hist;  // ||r_k||^2, k = 0..iter of the last solve (host copy)
  double *d_scratch = nullptr;    // small scalar scratch [64]
  void *h_pinned = nullptr;       // pinned host scratch (4 KB)
  int *d_errflag = nullptr;       // generic device error flag [4]
  // L2 flush buffer
  void *flush_buf = nullptr;
  size_t flush_bytes = 0;
  // options
  int64_t spmv_variant = 20;
  int64_t iters_per_sync = 16;
  int64_t use_graph = 1;          // (P)CG: replay chunks of iterations as a CUDA graph (0: plain launches)
  struct DfbGraphCache *graph = nullptr;
  int64_t assemble_variant = 0;
  int64_t assemble_tile_rows = 0;     // 0: chosen per mesh (about 224 off-diagonal slots per tile)
  int64_t assemble_tile_threads = 0;  // 0: one thread per off-diagonal slot + 4 per diagonal slot of the largest tile
  int64_t assemble_tile_neo = 0;  // 1: the neo-Hookean tet also takes the fused row-tile path (measured slower than element kernel + gather: 255 registers)
  int64_t spmv_ctas_per_sm = 0;  // 0 = auto
  int64_t spmv_tile_rows = 0;    // block-rows per packed tile: 0 = by row density (8, or 16 for <= 8 blocks per row), else 8 / 16
  int64_t ilu_level_launches = 0;  // 1: ILU sweeps as one launch per level (reference point) instead of the dependency-driven sweeps
  int64_t profile_spmv = 0;
  // SpMV profiling (events on the launching stream)
  std::vector<cudaEvent_t> prof_ev;   // pairs
  size_t prof_used = 0;
  double prof_ms = 0.0;
  uint64_t prof_count = 0;
  // phase profiling (option "profile_phases"): event pairs per phase id on the launching stream, folded lazily
  int64_t profile_phases = 0;
  std::vector<cudaEvent_t> phase_ev;   // pairs
  std::vector<int> phase_id;           // one per pair
  size_t phase_used = 0;               // events used
  double phase_ms[DFB_N_PHASES] = {};
  uint64_t phase_count[DFB_N_PHASES] = {};
  // multi-GPU
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // peer mailboxes (NVLink P2P scalar all-reduce fused into the PCG kernels)
  DfbMailbox *d_mbox = nullptr;                 // local, IPC-exported
  DfbMailbox *peer_mbox[DFB_MAX_RANKS] = {};    // mapped peers (own entry = d_mbox)
  int peer_enabled = 0;
  int64_t use_peer_allreduce = 1;               // option: 0 forces NCCL all-reduce even when mailboxes are mapped
  // peer halo push (ghost values stored straight into the neighbours' landing zones over NVLink; PCG only)
  int64_t use_peer_halo = 1;                    // option: ghost exchange by NVLink peer push inside (P)CG when the peers are mapped and dfb_halo_set_remote was called (0: NCCL send/recv)
  int64_t peer_halo_bytes = 8ll << 20;          // option (before dfb_ctx_peer_export): landing-zone bytes per parity buffer
  uint64_t landing_bytes = 0;                   // what the exported allocation actually holds per parity buffer
};

static inline double *dfb_peer_landing(const dfb_ctx *c, int r, int parity) {
  return (double *)((unsigned char *)c->peer_mbox[r] + sizeof(DfbMailbox) + (size_t)parity * c->landing_bytes);
}

// view handed to kernels; post = 0: the launch does not publish its reduction to the peers
static inline DfbPeerView dfb_peer_view(const dfb_ctx *c, int post) {
  DfbPeerView pv;
  pv.rank = c->rank;
  pv.nranks = (c->peer_enabled && c->use_peer_allreduce && c->nranks > 1) ? c->nranks : 1;
  pv.post = post;
  for (int r = 0; r < DFB_MAX_RANKS; ++r) pv.mbox[r] = c->peer_mbox[r];
  return pv;
}

struct dfb_vec {
  dfb_ctx *ctx;
  uint64_t n_blk, n_ghost;
  int blk_dim;
  double *d;
  uint64_t len() const { return n_blk * (uint64_t)blk_dim; }          // owned scalars
  uint64_t len_all() const { return (n_blk + n_ghost) * (uint64_t)blk_dim; }
};

struct dfb_bcflag {
  dfb_ctx *ctx;
  uint64_t n_blk_all;
  int blk_dim;
  int32_t *d;
};

struct dfb_mesh {
  dfb_ctx *ctx;
  uint64_t num_elem, num_vtx;
  int num_node, ndim;
  uint32_t *d_elem2vtx;
  double *d_xyz;
};

struct dfb_pattern {
  dfb_ctx *ctx;
  uint64_t num_blk;   // rows stored
  uint64_t nnz;
  int convention;
  uint32_t *d_row2idx;  // [num_blk+1]
  uint32_t *d_idx2col;  // [nnz] (+pad)
  int refcount;
};

struct dfb_halo {
  dfb_ctx *ctx;
  int n_peers;
  std::vector<int> peers;
  std::vector<uint64_t> send_ptr, recv_ptr;
  uint32_t *d_send_idx;
  double *d_sendbuf;     // packed, sized for blk_dim up to 4
  uint64_t n_send;
  int max_dim;
  // peer push (dfb_halo_set_remote): where this rank's packets land in each peer's ghost tail (in blocks)
  std::vector<uint64_t> remote_off;
  int peer_ready = 0;
};
struct DfbHaloPush {   // kernel argument of the pack-and-push launch
  int n_peers;
  int rank;
  uint64_t parity_stride;                     // doubles between the two parity buffers of a landing zone
  uint64_t send_ptr[DFB_MAX_RANKS + 1];
  double *dst[DFB_MAX_RANKS];                 // peer landing zone (parity 0) + remote offset, in doubles
  unsigned long long *flag[DFB_MAX_RANKS];    // &peer_mbox[peer]->halo_flag[rank]
};

// Matrix values live in ONE of two layouts (or both, when both are current):
//   canonical  idx2val[nnz*NN] + row2val[num_blk*NN]: exactly the reference's Vec<[T;NN]> arrays (upload/download, per-element merge,
//              ILU, sparse product, SpMV variants 0/4 work on these)
//   packed     the tile-packed SpMV mirror (bsr.cu): per 8 (or 16) block-rows one record [rowptr_rel | cols | diagonal blocks | value blocks]
// With spmv_variant 20 the packed layout is the NATIVE one: assembly, Dirichlet BC, add_to_diagonal, add_scaled, the Jacobi
// preconditioners and the SpMV all work on it directly, and the canonical arrays are only materialised (unpack kernel) when an entry
// point that needs them is called. Nothing is allocated until the values are first needed (a new matrix is "all zero").
struct dfb_bsr {
  dfb_ctx *ctx;
  dfb_pattern *pat;
  int blk_dim;
  double *d_idx2val = nullptr;  // [nnz * NN] (+pad)   canonical, lazily allocated
  double *d_row2val = nullptr;  // [num_blk * NN] (convention 0)
  int canon_valid = 0;          // canonical arrays hold the current values
  int32_t *d_flag = nullptr;    // uploaded BC flags (scratch, lazily allocated)
  uint64_t flag_cap = 0;
  dfb_halo *halo = nullptr;
  uint64_t int_begin = 0, int_end = 0;
  unsigned char *d_packed = nullptr;
  uint32_t *d_tile_off = nullptr;   // [n_tiles + 1], in 16-byte units
  uint64_t packed_cap = 0;          // bytes allocated
  uint64_t n_tiles = 0;
  int pk_rpt = 8;                   // block-rows per packed tile (8 or 16), fixed when the layout is built
  int packed_valid = 0;             // the packed records hold the current values
  int packed_ok = -1;               // -1 layout not built yet, 1 every tile fits the SpMV's shared-memory buffer, 0 some tile is oversized
                                    //  (its rows are multiplied straight from the canonical arrays, which then stay the native layout)
};

// device-side view of the packed layout. rpt = block-rows per tile: 8 (4 lanes per row in the SpMV) or, for sparse rows (<= 8 blocks per
// row on average, e.g. spring networks), 16 (2 lanes per row) so that a tile's single bulk copy stays ~8 KB
struct PackedView {
  unsigned char *base;
  const uint32_t *tile_off;
  int NN, has_dia, rpt, hdr;   // hdr = bytes of the record header: rpt + 1 relative row pointers, padded to 16
};
__host__ __device__ inline int dfb_pk_hdr_bytes(int rpt) { return ((rpt + 1) * 4 + 15) / 16 * 16; }
__device__ __forceinline__ unsigned char *pk_record(const PackedView &v, uint64_t tile) { return v.base + ((uint64_t)v.tile_off[tile] << 4); }
// diagonal block of row r
__device__ __forceinline__ double *pk_diag(const PackedView &v, uint64_t r) {
  unsigned char *rec = pk_record(v, r / v.rpt);
  const uint32_t nb = reinterpret_cast<const uint32_t *>(rec)[v.rpt];
  return reinterpret_cast<double *>(rec + v.hdr + ((nb * 4u + 15u) & ~15u)) + (r % v.rpt) * v.NN;
}
// off-diagonal blocks of row r: values, column indices, count
__device__ __forceinline__ double *pk_row(const PackedView &v, uint64_t r, const uint32_t **cols, uint32_t *n) {
  unsigned char *rec = pk_record(v, r / v.rpt);
  const uint32_t *hdr = reinterpret_cast<const uint32_t *>(rec);
  const uint32_t nb = hdr[v.rpt], kb = hdr[r % v.rpt], ke = hdr[r % v.rpt + 1];
  *cols = reinterpret_cast<const uint32_t *>(rec + v.hdr) + kb;
  *n = ke - kb;
  return reinterpret_cast<double *>(rec + v.hdr + ((nb * 4u + 15u) & ~15u)) + (v.has_dia ? v.rpt * v.NN : 0) + (uint64_t)kb * v.NN;
}
// first value (the tile's diagonal blocks, then its off-diagonal blocks, contiguous) of tile t
__device__ __forceinline__ double *pk_tile_values(const PackedView &v, uint64_t tile) {
  unsigned char *rec = pk_record(v, tile);
  const uint32_t nb = reinterpret_cast<const uint32_t *>(rec)[v.rpt];
  return reinterpret_cast<double *>(rec + v.hdr + ((nb * 4u + 15u) & ~15u));
}

// ---- layout management (bsr.cu)
bool dfb_bsr_packed_native(const dfb_bsr *m);        // spmv_variant 20 and every tile fits: the packed layout is the one entry points write
dfb_status dfb_bsr_canon_read(const dfb_bsr *m);     // canonical arrays allocated and current (zero-filled or unpacked on demand)
dfb_status dfb_bsr_canon_write(dfb_bsr *m);          // ... and about to be modified in place: the packed records become stale
dfb_status dfb_bsr_canon_overwrite(dfb_bsr *m);      // ... about to be overwritten completely (no unpack needed)
dfb_status dfb_bsr_packed_overwrite(dfb_bsr *m);     // packed layout built, every value about to be overwritten: canonical arrays become stale
dfb_status dfb_bsr_packed_modify(dfb_bsr *m);        // packed records current and about to be modified in place: canonical arrays become stale
PackedView dfb_bsr_packed_view(const dfb_bsr *m);

struct dfb_plan {
  dfb_ctx *ctx;
  dfb_pattern *pat;
  uint64_t num_elem;
  int num_node;
  int flavour;
  uint64_t num_slots;    // nnz (+ num_blk for convention 0)
  uint64_t num_contrib;
  uint32_t *d_nz2ptr;    // [num_slots+1]
  uint32_t *d_contrib;   // [num_contrib] code = e*nn2 + i*num_node + j
  const dfb_mesh *mesh;  // borrowed (elem2vtx used to rebuild elem2nz on demand)
  void *d_geo = nullptr;   // node-record scratch of the slot-centric gather (variant 4 / fallback; lazily allocated)
  uint64_t geo_cap = 0;
  // fused row-tile assembly (assemble_tile.cu, the default): per tile of 16 rows the sorted unique incident elements, and the
  // contribution lists recoded to 16 bits (tile-local element << 4 | i << 2 | j), parallel to d_contrib
  uint32_t *d_tile_elem = nullptr;   // tile t's unique elements (sorted), at offset (first pair of the tile) - pair_base
  uint32_t *d_tile_meta = nullptr;   // [n_atiles][8]: k0, k1, ob, oe, pb, pe, n_elem, pad
  uint16_t *d_code16 = nullptr;      // [num_contrib]
  uint64_t n_atiles = 0;
  uint32_t pair_base = 0, tile_max_elem = 0, tile_max_slots = 0, tile_rows = 16;
  int tile_state = 0;      // 0 not built yet, 1 usable, -1 not applicable
};
// assemble_tile.cu: DFB_ERR_UNSUPPORTED (no error string, matrix untouched) = the tile path does not apply, use the gather of assemble.cu
dfb_status dfb_plan_build_tiles(dfb_plan *plan);
void dfb_plan_free_tiles(dfb_plan *plan);
dfb_status dfb_assemble_tile_linear(dfb_plan *plan, int kind, const dfb_mesh *mesh, double p0, double p1, dfb_bsr *m);
dfb_status dfb_assemble_tile_energy(dfb_plan *plan, int kind, const dfb_mesh *mesh, const dfb_vec *disp, double p0, double p1, dfb_bsr *m,
                                    double *dw, double *w_elem);

struct DfbIlu;
struct dfb_precond {
  dfb_ctx *ctx;
  int kind;
  int blk_dim;
  uint64_t num_blk;
  double *d_inv;  // jacobi: [num_blk*N] ; block-jacobi / ilu0: [num_blk*NN] inverse (pivot) blocks (column-major)
  DfbIlu *ilu = nullptr;        // kind == DFB_PRECOND_ILU0
  dfb_pattern *pat = nullptr;   // pattern of the matrix the ILU was built for
};
// ilu.cu
dfb_status dfb_ilu_create(dfb_ctx *c, const dfb_bsr *m, uint64_t lev_fill, DfbIlu **out);
void dfb_ilu_destroy(DfbIlu *s);
dfb_status dfb_ilu_update(dfb_ctx *c, DfbIlu *s, const dfb_bsr *m, double *d_dia_out);
dfb_status dfb_ilu_apply(dfb_ctx *c, const DfbIlu *s, const dfb_pattern *pat, int N, const double *d_dia, double *d_vec);
double *dfb_ilu_z(DfbIlu *s);
dfb_status dfb_ilu_dot(dfb_ctx *c, const double *a, const double *b, uint64_t n, double *d_out);

// ---- helpers implemented in ctx.cu
dfb_status dfb_dev_alloc(void **p, size_t bytes);
void dfb_dev_free(void *p);
uint64_t dfb_free_epoch();  // number of dfb_dev_free calls so far: cached graphs bake raw device pointers and must not outlive them
template <typename T>
static inline dfb_status dfb_dev_alloc_t(T **p, size_t count) {
  return dfb_dev_alloc((void **)p, count * sizeof(T) + 256);
}
dfb_status dfb_check_error_flag(dfb_ctx *ctx, int slot, dfb_status st, const char *what);

// ---- cross-file entry points
dfb_status dfb_exclusive_scan_u32(dfb_ctx *ctx, uint32_t *d_data, uint64_t n, uint64_t *total_out);
dfb_status dfb_nccl_allreduce_sum(dfb_ctx *ctx, double *d_buf, int count, cudaStream_t strm);
dfb_status dfb_halo_exchange_on(dfb_halo *h, double *d_vec, uint64_t n_blk, int blk_dim, cudaStream_t strm);
bool dfb_halo_peer_usable(const dfb_halo *h, int blk_dim);
dfb_status dfb_halo_push(dfb_halo *h, const double *d_vec, int blk_dim, cudaStream_t strm);
void dfb_halo_wait_view(const dfb_halo *h, int blk_dim, uint64_t n_own, DfbHaloWait *wait_out);
// in_solver: the launch belongs to a (P)CG iteration (reads iteration / stop flag / exchange ids from ctx->d_state)
dfb_status dfb_spmv_launch(const dfb_bsr *m, double *y, double beta, double alpha, const double *x, uint64_t row_begin,
                           uint64_t row_end, int fuse_dot, int in_solver, cudaStream_t strm, int post = 0,
                           const DfbHaloWait *hw = nullptr);
dfb_status dfb_bsr_ensure_packed(const dfb_bsr *m);

static inline int dfb_div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }
