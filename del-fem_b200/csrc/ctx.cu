// ctx.cu — context, errors, timers, device memory helpers, exclusive scan
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "dfb_internal.h"

void dfb_comm_teardown(dfb_ctx *ctx);      // dist.cu
void dfb_graph_cache_destroy(dfb_ctx *ctx);  // solver.cu

static thread_local char g_err[1024] = "";

void dfb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

dfb_status dfb_fail(dfb_status st, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return st;
}

DFB_API const char *dfb_last_error_message(void) { return g_err; }
DFB_API const char *dfb_version(void) { return "delfem_b200 0.1.0 (sm_100a, fp64)"; }

// Device memory comes from the device's stream-ordered pool with an unlimited release threshold: buffers that a handle
// frees are recycled by the next handle without a driver round trip (a one-shot "host buffers in, solution out" step at S10
// allocates and frees ~14 GB; plain cudaFree alone cost 13-113 ms of it). Semantics are kept identical to cudaMalloc/cudaFree:
// the allocation is complete when dfb_dev_alloc returns, and dfb_dev_free waits for the device like cudaFree does.
static cudaStream_t g_alloc_stream[64] = {};

static dfb_status alloc_stream(cudaStream_t *out) {
  int dev = 0;
  DFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return dfb_fail(DFB_ERR_CUDA, "device ordinal %d out of range", dev);
  if (!g_alloc_stream[dev]) {
    DFB_CUDA(cudaStreamCreateWithFlags(&g_alloc_stream[dev], cudaStreamNonBlocking));
    cudaMemPool_t pool;
    DFB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long thr = ~0ull;
    DFB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
  *out = g_alloc_stream[dev];
  return DFB_OK;
}

dfb_status dfb_dev_alloc(void **p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 256;
  cudaStream_t s;
  DFB_TRY(alloc_stream(&s));
  DFB_CUDA(cudaMallocAsync(p, bytes, s));
  DFB_CUDA(cudaStreamSynchronize(s));
  return DFB_OK;
}

static uint64_t g_free_epoch = 0;
uint64_t dfb_free_epoch() { return g_free_epoch; }

void dfb_dev_free(void *p) {
  if (!p) return;
  g_free_epoch += 1;
  cudaStream_t s;
  if (alloc_stream(&s) != DFB_OK) return;
  cudaDeviceSynchronize();  // like cudaFree: nobody is still using the buffer
  cudaFreeAsync(p, s);
}

DFB_API dfb_status dfb_ctx_create(int32_t device, dfb_ctx **out) {
  DFB_REQUIRE(out != nullptr, "dfb_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return dfb_fail(DFB_ERR_CUDA, "dfb_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
  DFB_REQUIRE(device >= 0 && device < ndev, "dfb_ctx_create: device %d out of range [0,%d)", device, ndev);
  DFB_CUDA(cudaSetDevice(device));
  dfb_ctx *c = new dfb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  DFB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->l2_bytes = prop.l2CacheSize;
  c->total_mem = (int64_t)prop.totalGlobalMem;
  DFB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  DFB_CUDA(cudaStreamCreateWithFlags(&c->stream_comm, cudaStreamNonBlocking));
  DFB_CUDA(cudaEventCreate(&c->ev_t0));
  DFB_CUDA(cudaEventCreate(&c->ev_t1));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_flag[0], cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&c->ev_flag[1], cudaEventDisableTiming));
  DFB_TRY(dfb_dev_alloc_t(&c->d_partials, (size_t)DFB_RED_MAX_BLOCKS * DFB_RED_K));
  DFB_TRY(dfb_dev_alloc_t(&c->d_ticket, 8));
  DFB_CUDA(cudaMemsetAsync(c->d_ticket, 0, 8 * sizeof(unsigned int), c->stream));
  DFB_TRY(dfb_dev_alloc_t(&c->d_state, 1));
  DFB_CUDA(cudaMemsetAsync(c->d_state, 0, sizeof(DfbSolverState), c->stream));
  DFB_TRY(dfb_dev_alloc_t(&c->d_scratch, 64));
  DFB_TRY(dfb_dev_alloc_t(&c->d_errflag, 4));
  DFB_CUDA(cudaMemsetAsync(c->d_errflag, 0, 4 * sizeof(int), c->stream));
  DFB_CUDA(cudaMallocHost(&c->h_pinned, 4096));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  *out = c;
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_destroy(dfb_ctx *c) {
  if (!c) return DFB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream_comm);
  dfb_graph_cache_destroy(c);
  dfb_comm_teardown(c);
  if (c->h_poll) cudaFreeHost(c->h_poll);
  dfb_dev_free(c->d_partials);
  dfb_dev_free(c->d_ticket);
  dfb_dev_free(c->d_state);
  dfb_dev_free(c->d_hist);
  dfb_dev_free(c->d_scratch);
  dfb_dev_free(c->d_errflag);
  dfb_dev_free(c->flush_buf);
  cudaFreeHost(c->h_pinned);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->phase_ev) cudaEventDestroy(e);
  cudaEventDestroy(c->ev_t0);
  cudaEventDestroy(c->ev_t1);
  cudaEventDestroy(c->ev_a);
  cudaEventDestroy(c->ev_b);
  cudaEventDestroy(c->ev_c);
  cudaEventDestroy(c->ev_flag[0]);
  cudaEventDestroy(c->ev_flag[1]);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->stream_comm);
  delete c;
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_sync(dfb_ctx *c) {
  DFB_REQUIRE(c, "dfb_ctx_sync: ctx is NULL");
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream_comm));
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_launch_count(dfb_ctx *c, uint64_t *out) {
  DFB_REQUIRE(c && out, "dfb_ctx_launch_count: NULL argument");
  *out = c->launches;
  return DFB_OK;
}

static int64_t *option_slot(dfb_ctx *c, const char *key) {
  if (!strcmp(key, "spmv_variant")) return &c->spmv_variant;
  if (!strcmp(key, "iters_per_sync")) return &c->iters_per_sync;
  if (!strcmp(key, "use_graph")) return &c->use_graph;
  if (!strcmp(key, "assemble_variant")) return &c->assemble_variant;
  if (!strcmp(key, "assemble_tile_neo")) return &c->assemble_tile_neo;
  if (!strcmp(key, "assemble_tile_rows")) return &c->assemble_tile_rows;
  if (!strcmp(key, "assemble_tile_threads")) return &c->assemble_tile_threads;
  if (!strcmp(key, "spmv_ctas_per_sm")) return &c->spmv_ctas_per_sm;
  if (!strcmp(key, "spmv_tile_rows")) return &c->spmv_tile_rows;
  if (!strcmp(key, "profile_spmv")) return &c->profile_spmv;
  if (!strcmp(key, "ilu_level_launches")) return &c->ilu_level_launches;
  if (!strcmp(key, "profile_phases")) return &c->profile_phases;
  if (!strcmp(key, "use_peer_allreduce")) return &c->use_peer_allreduce;
  if (!strcmp(key, "use_peer_halo")) return &c->use_peer_halo;
  if (!strcmp(key, "peer_halo_bytes")) return &c->peer_halo_bytes;
  return nullptr;
}

DFB_API dfb_status dfb_ctx_set_option(dfb_ctx *c, const char *key, int64_t value) {
  DFB_REQUIRE(c && key, "dfb_ctx_set_option: NULL argument");
  int64_t *s = option_slot(c, key);
  DFB_REQUIRE(s, "dfb_ctx_set_option: unknown option '%s'", key);
  *s = value;
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_get_option(dfb_ctx *c, const char *key, int64_t *value) {
  DFB_REQUIRE(c && key && value, "dfb_ctx_get_option: NULL argument");
  int64_t *s = option_slot(c, key);
  DFB_REQUIRE(s, "dfb_ctx_get_option: unknown option '%s'", key);
  *value = *s;
  return DFB_OK;
}

DFB_API dfb_status dfb_timer_start(dfb_ctx *c) {
  DFB_REQUIRE(c, "dfb_timer_start: ctx is NULL");
  DFB_CUDA(cudaEventRecord(c->ev_t0, c->stream));
  return DFB_OK;
}

DFB_API dfb_status dfb_timer_stop(dfb_ctx *c, float *ms) {
  DFB_REQUIRE(c && ms, "dfb_timer_stop: NULL argument");
  DFB_CUDA(cudaEventRecord(c->ev_t1, c->stream));
  DFB_CUDA(cudaEventSynchronize(c->ev_t1));
  DFB_CUDA(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
  return DFB_OK;
}

__global__ void k_flush_fill(uint4 *p, size_t n, unsigned v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = make_uint4(v, v, v, v);
}

DFB_API dfb_status dfb_flush_l2(dfb_ctx *c) {
  DFB_REQUIRE(c, "dfb_flush_l2: ctx is NULL");
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)(c->l2_bytes > 0 ? c->l2_bytes : (128ll << 20)) * 2;
    DFB_TRY(dfb_dev_alloc(&c->flush_buf, c->flush_bytes));
  }
  static unsigned counter = 0;
  DFB_LAUNCH(c, k_flush_fill, c->sm_count * 8, 256, 0, (uint4 *)c->flush_buf, c->flush_bytes / 16, ++counter);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

DFB_API dfb_status dfb_device_info(dfb_ctx *c, int32_t *sm_count, int64_t *l2_bytes, int64_t *total_mem_bytes) {
  DFB_REQUIRE(c, "dfb_device_info: ctx is NULL");
  if (sm_count) *sm_count = c->sm_count;
  if (l2_bytes) *l2_bytes = c->l2_bytes;
  if (total_mem_bytes) *total_mem_bytes = c->total_mem;
  return DFB_OK;
}

DFB_API dfb_status dfb_host_register(dfb_ctx *c, void *host, uint64_t bytes) {
  DFB_REQUIRE(c && host, "dfb_host_register: NULL argument");
  DFB_CUDA(cudaSetDevice(c->device));
  DFB_CUDA(cudaHostRegister(host, bytes, cudaHostRegisterDefault));
  return DFB_OK;
}
DFB_API dfb_status dfb_host_unregister(dfb_ctx *c, void *host) {
  DFB_REQUIRE(c && host, "dfb_host_unregister: NULL argument");
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  DFB_CUDA(cudaHostUnregister(host));
  return DFB_OK;
}

// fold finished event pairs into the running totals (synchronises the stream)
static dfb_status prof_collect(dfb_ctx *c, size_t max_pairs = (size_t)-1) {
  if (c->prof_used == 0) return DFB_OK;
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i + 1 < c->prof_used && i / 2 < max_pairs; i += 2) {
    float ms = 0.f;
    DFB_CUDA(cudaEventElapsedTime(&ms, c->prof_ev[i], c->prof_ev[i + 1]));
    c->prof_ms += ms;
    c->prof_count += 1;
  }
  c->prof_used = 0;
  return DFB_OK;
}

DFB_API dfb_status dfb_profile_spmv(dfb_ctx *c, double *total_ms, uint64_t *count, int32_t reset) {
  DFB_REQUIRE(c, "dfb_profile_spmv: ctx is NULL");
  DFB_TRY(prof_collect(c));
  if (total_ms) *total_ms = c->prof_ms;
  if (count) *count = c->prof_count;
  if (reset) {
    c->prof_ms = 0.0;
    c->prof_count = 0;
  }
  return DFB_OK;
}

// called by the solvers around an SpMV launch; returns false when the per-solve budget of event pairs is used up
bool dfb_prof_begin(dfb_ctx *c) {
  if (!c->profile_spmv) return false;
  if (c->prof_used + 2 > 32768) return false;
  while (c->prof_ev.size() < c->prof_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return false;
    c->prof_ev.push_back(e);
  }
  cudaEventRecord(c->prof_ev[c->prof_used], c->stream);
  return true;
}
void dfb_prof_end(dfb_ctx *c) {
  cudaEventRecord(c->prof_ev[c->prof_used + 1], c->stream);
  c->prof_used += 2;
}
// the solvers enqueue iterations ahead of the convergence test: only the first `max_pairs` recorded launches did real work
dfb_status dfb_prof_flush(dfb_ctx *c, size_t max_pairs) { return prof_collect(c, max_pairs); }

// ---- phase profiler: dfb_phase_begin/_end bracket a group of launches with events when option "profile_phases" is set
void dfb_phase_begin(dfb_ctx *c, int id, cudaStream_t strm) {
  if (!c->profile_phases) return;
  if (c->phase_used + 2 > 65536) return;
  while (c->phase_ev.size() < c->phase_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c->phase_ev.push_back(e);
  }
  if (c->phase_id.size() < c->phase_ev.size() / 2) c->phase_id.resize(c->phase_ev.size() / 2, 0);
  c->phase_id[c->phase_used / 2] = id;
  cudaEventRecord(c->phase_ev[c->phase_used], strm ? strm : c->stream);
  c->phase_used += 1;
}
void dfb_phase_end(dfb_ctx *c, cudaStream_t strm) {
  if (!c->profile_phases || (c->phase_used & 1) == 0) return;
  cudaEventRecord(c->phase_ev[c->phase_used], strm ? strm : c->stream);
  c->phase_used += 1;
}

DFB_API dfb_status dfb_ctx_phase_times(dfb_ctx *c, double *ms_out, uint64_t *count_out, int32_t reset) {
  DFB_REQUIRE(c, "dfb_ctx_phase_times: ctx is NULL");
  if (c->phase_used) {
    DFB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < c->phase_used; i += 2) {
      float ms = 0.f;
      DFB_CUDA(cudaEventElapsedTime(&ms, c->phase_ev[i], c->phase_ev[i + 1]));
      const int id = c->phase_id[i / 2];
      c->phase_ms[id] += ms;
      c->phase_count[id] += 1;
    }
    c->phase_used = 0;
  }
  for (int k = 0; k < DFB_N_PHASES; ++k) {
    if (ms_out) ms_out[k] = c->phase_ms[k];
    if (count_out) count_out[k] = c->phase_count[k];
    if (reset) {
      c->phase_ms[k] = 0.0;
      c->phase_count[k] = 0;
    }
  }
  return DFB_OK;
}

DFB_API dfb_status dfb_malloc(dfb_ctx *c, uint64_t bytes, void **dev_ptr) {
  DFB_REQUIRE(c && dev_ptr, "dfb_malloc: NULL argument");
  DFB_CUDA(cudaSetDevice(c->device));
  return dfb_dev_alloc(dev_ptr, bytes + 256);
}
DFB_API dfb_status dfb_free(dfb_ctx *c, void *dev_ptr) {
  DFB_REQUIRE(c, "dfb_free: ctx is NULL");
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  dfb_dev_free(dev_ptr);
  return DFB_OK;
}
DFB_API dfb_status dfb_memcpy_d2h(dfb_ctx *c, void *host, const void *dev, uint64_t bytes) {
  DFB_REQUIRE(c && host && dev, "dfb_memcpy_d2h: NULL argument");
  DFB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  return DFB_OK;
}
DFB_API dfb_status dfb_memcpy_h2d(dfb_ctx *c, void *dev, const void *host, uint64_t bytes) {
  DFB_REQUIRE(c && host && dev, "dfb_memcpy_h2d: NULL argument");
  DFB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  return DFB_OK;
}

// reads (and clears) a device error flag; returns `st` if it was raised
dfb_status dfb_check_error_flag(dfb_ctx *c, int slot, dfb_status st, const char *what) {
  int *h = (int *)c->h_pinned + 512 + slot;
  DFB_CUDA(cudaMemcpyAsync(h, c->d_errflag + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  if (*h != 0) {
    DFB_CUDA(cudaMemsetAsync(c->d_errflag + slot, 0, sizeof(int), c->stream));
    return dfb_fail(st, "%s", what);
  }
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of a u32 array in place (three-phase: tile sums, scan of tile sums, rescan with offsets).
// Used by the pattern / plan builders. data[n] (one past the end) receives the total as well.
// ------------------------------------------------------------------------------------------------
static const int SCAN_THREADS = 512;
static const int SCAN_ITEMS = 8;
static const int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (unsigned)o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned *total) {
  __shared__ unsigned wsum[33];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const unsigned w = (lane < nw) ? wsum[lane] : 0u;
    const unsigned winc = warp_incl_scan(w);
    wsum[lane] = winc - w;           // exclusive offset of warp `lane`
    if (lane == 31) wsum[32] = winc; // block total
  }
  __syncthreads();
  *total = wsum[32];
  const unsigned r = wsum[wid] + inc - v;
  __syncthreads();
  return r;
}

__global__ void k_scan_tile_sums(const uint32_t *data, uint64_t n, unsigned long long *tile_sums) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += data[i];
  }
  unsigned tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of tile_sums (64-bit), total to tile_sums[n_tiles]
__global__ void k_scan_tile_offsets(unsigned long long *tile_sums, uint64_t n_tiles) {
  __shared__ unsigned long long carry;
  __shared__ unsigned long long buf[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_tiles; base += blockDim.x) {
    uint64_t i = base + threadIdx.x;
    unsigned long long v = (i < n_tiles) ? tile_sums[i] : 0ull;
    buf[threadIdx.x] = v;
    __syncthreads();
    // Hillis-Steele inclusive scan in shared memory
    for (unsigned o = 1; o < blockDim.x; o <<= 1) {
      unsigned long long t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0ull;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    unsigned long long inc = buf[threadIdx.x];
    if (i < n_tiles) tile_sums[i] = carry + inc - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void k_scan_apply(uint32_t *data, uint64_t n, const unsigned long long *tile_offsets) {
  // each thread owns SCAN_ITEMS consecutive items so that the in-thread prefix is sequential
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  unsigned v[SCAN_ITEMS];
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + k;
    v[k] = (i < n) ? data[i] : 0u;
    s += v[k];
  }
  unsigned tot;
  unsigned excl = block_excl_scan(s, &tot);
  unsigned long long off = tile_offsets[blockIdx.x] + excl;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint64_t i = base + k;
    if (i < n) data[i] = (uint32_t)off;
    off += v[k];
  }
}

__global__ void k_scan_write_total(uint32_t *data, uint64_t n, const unsigned long long *tile_offsets, uint64_t n_tiles) {
  data[n] = (uint32_t)tile_offsets[n_tiles];
}

// k_scan_tile_sums reads items strided, k_scan_apply reads them blocked: the per-tile SUM is the same either way.
dfb_status dfb_exclusive_scan_u32(dfb_ctx *c, uint32_t *d_data, uint64_t n, uint64_t *total_out) {
  const uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  unsigned long long *d_tiles = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_tiles, n_tiles + 2));
  if (n_tiles > 0) {
    DFB_LAUNCH(c, k_scan_tile_sums, (unsigned)n_tiles, SCAN_THREADS, 0, d_data, n, d_tiles);
    DFB_LAUNCH(c, k_scan_tile_offsets, 1, 1024, 0, d_tiles, n_tiles);
    DFB_LAUNCH(c, k_scan_apply, (unsigned)n_tiles, SCAN_THREADS, 0, d_data, n, d_tiles);
  } else {
    DFB_CUDA(cudaMemsetAsync(d_tiles, 0, 16, c->stream));
  }
  DFB_LAUNCH(c, k_scan_write_total, 1, 1, 0, d_data, n, d_tiles, n_tiles);
  DFB_CHECK_LAUNCH();
  unsigned long long *h = (unsigned long long *)c->h_pinned;
  DFB_CUDA(cudaMemcpyAsync(h, d_tiles + n_tiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  const uint64_t total = *h;
  dfb_dev_free(d_tiles);
  if (total_out) *total_out = total;
  if (total > 0xFFFFFFF0ull) return dfb_fail(DFB_ERR_INDEX_OVERFLOW, "scan total %llu does not fit u32", (unsigned long long)total);
  return DFB_OK;
}
