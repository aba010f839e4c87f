// dist.cu — multi-GPU plumbing: one rank (process) per GPU, NCCL over NVLink 5 / NVSwitch.
//   * SpMV halo: pack kernel -> grouped ncclSend/ncclRecv; the receive lands directly in the ghost tail of the vector
//   * PCG scalars: ncclAllReduce(sum, fp64, 1..2) on the solver-state buffer, stream-ordered (no host sync)
// NCCL is resolved lazily with dlopen so that single-GPU users need no NCCL at all (and so that a process that has
// already loaded PyTorch's bundled libnccl.so.2 shares it).  The reference has no multi-device code.
#include <dlfcn.h>
#include <string.h>

#include "dfb_internal.h"

typedef void *nccl_comm_t;
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_GetUniqueId)(nccl_uid_t *);
typedef int (*fn_CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int);
typedef int (*fn_CommDestroy)(nccl_comm_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char *(*fn_ErrStr)(int);

static struct {
  void *h;
  fn_GetUniqueId GetUniqueId;
  fn_CommInitRank CommInitRank;
  fn_CommDestroy CommDestroy;
  fn_AllReduce AllReduce;
  fn_Send Send;
  fn_Recv Recv;
  fn_Group GroupStart, GroupEnd;
  fn_ErrStr GetErrorString;
} g_nccl = {};

static const int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

static dfb_status nccl_load() {
  if (g_nccl.h) return DFB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.h; ++i) g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.h) return dfb_fail(DFB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
  g_nccl.field = (decltype(g_nccl.field))dlsym(g_nccl.h, name);                  \
  if (!g_nccl.field) return dfb_fail(DFB_ERR_NCCL, "libnccl: missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return DFB_OK;
}

#define DFB_NCCL(expr)                                                                                   \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != 0) return dfb_fail(DFB_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r));       \
  } while (0)

DFB_API dfb_status dfb_comm_unique_id(uint8_t id_out[128]) {
  DFB_REQUIRE(id_out, "dfb_comm_unique_id: NULL argument");
  DFB_TRY(nccl_load());
  nccl_uid_t uid;
  DFB_NCCL(g_nccl.GetUniqueId(&uid));
  memcpy(id_out, uid.internal, 128);
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_comm_init(dfb_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t nranks) {
  DFB_REQUIRE(ctx && id, "dfb_ctx_comm_init: NULL argument");
  DFB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "dfb_ctx_comm_init: bad rank %d / %d", rank, nranks);
  DFB_TRY(nccl_load());
  DFB_CUDA(cudaSetDevice(ctx->device));
  nccl_uid_t uid;
  memcpy(uid.internal, id, 128);
  nccl_comm_t comm = nullptr;
  DFB_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return DFB_OK;
}

dfb_status dfb_nccl_allreduce_sum(dfb_ctx *ctx, double *d_buf, int count, cudaStream_t strm) {
  if (ctx->nranks <= 1) return DFB_OK;
  DFB_REQUIRE(ctx->nccl_comm, "all-reduce requested but the context has no communicator");
  DFB_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, strm));
  return DFB_OK;
}

// called by dfb_ctx_destroy: unmap the peers' mailboxes, free the own one, destroy the communicator
void dfb_comm_teardown(dfb_ctx *ctx) {
  for (int r = 0; r < DFB_MAX_RANKS; ++r) {
    if (ctx->peer_mbox[r] && ctx->peer_mbox[r] != ctx->d_mbox) cudaIpcCloseMemHandle(ctx->peer_mbox[r]);
    ctx->peer_mbox[r] = nullptr;
  }
  ctx->peer_enabled = 0;
  if (ctx->d_mbox) cudaFree(ctx->d_mbox);
  ctx->d_mbox = nullptr;
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

// ------------------------------------------------------------------------------------------------ peer mailboxes
DFB_API dfb_status dfb_ctx_peer_export(dfb_ctx *ctx, uint8_t handle_out[64]) {
  DFB_REQUIRE(ctx && handle_out, "dfb_ctx_peer_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DFB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->d_mbox) {
    // one IPC allocation: [mailbox | landing zone parity 0 | landing zone parity 1]
    ctx->landing_bytes = ctx->peer_halo_bytes > 0 ? (((uint64_t)ctx->peer_halo_bytes + 127) & ~127ull) : 0;
    DFB_CUDA(cudaMalloc((void **)&ctx->d_mbox, sizeof(DfbMailbox) + 2 * ctx->landing_bytes));
    DFB_CUDA(cudaMemset(ctx->d_mbox, 0, sizeof(DfbMailbox)));
  }
  cudaIpcMemHandle_t h;
  DFB_CUDA(cudaIpcGetMemHandle(&h, ctx->d_mbox));
  memcpy(handle_out, &h, 64);
  return DFB_OK;
}

DFB_API dfb_status dfb_ctx_peer_import(dfb_ctx *ctx, const uint8_t *handles, int32_t nranks) {
  DFB_REQUIRE(ctx && handles, "dfb_ctx_peer_import: NULL argument");
  DFB_REQUIRE(nranks == ctx->nranks && nranks <= DFB_MAX_RANKS, "dfb_ctx_peer_import: nranks %d does not match the communicator (%d, max %d)",
              nranks, ctx->nranks, DFB_MAX_RANKS);
  DFB_REQUIRE(ctx->d_mbox, "dfb_ctx_peer_import: call dfb_ctx_peer_export first");
  DFB_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < nranks; ++r) {
    if (r == ctx->rank) {
      ctx->peer_mbox[r] = ctx->d_mbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    void *p = nullptr;
    DFB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_mbox[r] = (DfbMailbox *)p;
  }
  ctx->peer_enabled = 1;
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ halo
// Peer push: the boundary entries of the vector are stored straight into the neighbours' landing zones over NVLink; the last
// block to finish fences and then raises this rank's flag in every neighbour's mailbox.  The matching wait lives in the
// boundary tiles of the neighbour's SpMV (bsr.cu), so a ghost exchange costs one small launch and no collective.
__global__ void __launch_bounds__(256) k_halo_push(const double *__restrict__ v, const uint32_t *__restrict__ idx, int dim,
                                                    const DfbHaloPush hp, const DfbSolverState *st, unsigned int *ticket) {
  // runs after the kernel that produced v (init: iteration 0's direction; the direction kernel of iteration i-1: st->iter == i)
  const int it = st->iter;
  if (st->done[it & 1]) return;
  const unsigned long long seq = dfb_hid(st->hbase, it);
  const uint64_t par_off = (seq & 1ull) ? hp.parity_stride : 0;
  __shared__ bool s_last;
  const uint64_t total = hp.send_ptr[hp.n_peers] * (uint64_t)dim;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = q / dim;
    const int d = (int)(q - s * dim);
    int p = 0;
    while (p + 1 < hp.n_peers && s >= hp.send_ptr[p + 1]) ++p;
    hp.dst[p][par_off + (s - hp.send_ptr[p]) * dim + d] = v[(uint64_t)idx[s] * dim + d];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (s_last && (int)threadIdx.x < hp.n_peers) {
    __threadfence_system();
    volatile unsigned long long *f = hp.flag[threadIdx.x];
    *f = seq;
  }
}

__global__ void k_halo_pack(const double *__restrict__ v, const uint32_t *__restrict__ idx, uint64_t n_send, int dim,
                            double *__restrict__ buf) {
  const uint64_t total = n_send * dim;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t s = q / dim;
    const int d = (int)(q - s * dim);
    buf[q] = v[(uint64_t)idx[s] * dim + d];
  }
}

DFB_API dfb_status dfb_halo_create(dfb_ctx *ctx, int32_t n_peers, const int32_t *peers, const uint64_t *send_ptr,
                                   const uint32_t *send_idx, const uint64_t *recv_ptr, dfb_halo **out) {
  DFB_REQUIRE(ctx && out, "dfb_halo_create: NULL argument");
  DFB_REQUIRE(n_peers >= 0, "dfb_halo_create: n_peers < 0");
  DFB_REQUIRE(n_peers == 0 || (peers && send_ptr && recv_ptr), "dfb_halo_create: NULL array");
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_halo *h = new dfb_halo();
  h->ctx = ctx;
  h->n_peers = n_peers;
  h->peers.assign(peers, peers + n_peers);
  h->send_ptr.assign(1, 0);
  h->recv_ptr.assign(1, 0);
  if (n_peers > 0) {
    h->send_ptr.assign(send_ptr, send_ptr + n_peers + 1);
    h->recv_ptr.assign(recv_ptr, recv_ptr + n_peers + 1);
  }
  h->n_send = h->send_ptr.back();
  h->max_dim = 4;
  h->d_send_idx = nullptr;
  h->d_sendbuf = nullptr;
  dfb_status st = dfb_dev_alloc_t(&h->d_send_idx, h->n_send + 2);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&h->d_sendbuf, h->n_send * h->max_dim + 2);
  if (st == DFB_OK && h->n_send > 0 && !send_idx) st = dfb_fail(DFB_ERR_BAD_ARG, "dfb_halo_create: send_idx is NULL");
  if (st == DFB_OK && h->n_send > 0) {
    cudaError_t e = cudaMemcpyAsync(h->d_send_idx, send_idx, h->n_send * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_halo_create: %s", cudaGetErrorString(e));
  }
  if (st != DFB_OK) {
    dfb_halo_destroy(h);
    return st;
  }
  *out = h;
  return DFB_OK;
}

DFB_API dfb_status dfb_halo_set_remote(dfb_halo *h, const uint64_t *remote_off) {
  DFB_REQUIRE(h, "dfb_halo_set_remote: NULL argument");
  DFB_REQUIRE(h->n_peers == 0 || remote_off, "dfb_halo_set_remote: NULL array");
  DFB_REQUIRE(h->n_peers <= DFB_MAX_RANKS, "dfb_halo_set_remote: more than %d peers", DFB_MAX_RANKS);
  h->remote_off.assign(remote_off, remote_off + h->n_peers);
  h->peer_ready = 1;
  return DFB_OK;
}

// the push path needs mapped peers, remote offsets, and landing zones large enough for what is sent AND received
bool dfb_halo_peer_usable(const dfb_halo *h, int blk_dim) {
  const dfb_ctx *c = h->ctx;
  if (!h->peer_ready || !c->peer_enabled || !c->use_peer_halo || c->nranks <= 1 || c->landing_bytes == 0) return false;
  const uint64_t cap_blk = c->landing_bytes / ((uint64_t)blk_dim * sizeof(double));
  if (h->recv_ptr.back() > cap_blk) return false;
  for (int p = 0; p < h->n_peers; ++p)
    if (h->remote_off[p] + (h->send_ptr[p + 1] - h->send_ptr[p]) > cap_blk) return false;
  return true;
}

// push the boundary entries of d_vec into the neighbours' landing zones; exchange id and parity buffer come from the solver state
dfb_status dfb_halo_push(dfb_halo *h, const double *d_vec, int blk_dim, cudaStream_t strm) {
  dfb_ctx *c = h->ctx;
  if (h->n_peers == 0) return DFB_OK;
  DfbHaloPush hp;
  hp.n_peers = h->n_peers;
  hp.rank = c->rank;
  hp.parity_stride = c->landing_bytes / sizeof(double);
  for (int p = 0; p <= h->n_peers; ++p) hp.send_ptr[p] = h->send_ptr[p];
  for (int p = 0; p < h->n_peers; ++p) {
    const int pr = h->peers[p];
    hp.dst[p] = dfb_peer_landing(c, pr, 0) + h->remote_off[p] * (uint64_t)blk_dim;
    hp.flag[p] = &c->peer_mbox[pr]->halo_flag[c->rank];
  }
  int g = dfb_div_up(h->n_send * blk_dim, 256);
  if (g < 1) g = 1;
  if (g > c->sm_count * 2) g = c->sm_count * 2;
  DFB_LAUNCH_ON(c, strm, k_halo_push, g, 256, 0, d_vec, h->d_send_idx, blk_dim, hp, c->d_state, c->d_ticket + 5);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// receiver side: where the neighbours' pushes land on THIS rank and whose flags to wait for
void dfb_halo_wait_view(const dfb_halo *h, int blk_dim, uint64_t n_own, DfbHaloWait *w) {
  const dfb_ctx *c = h->ctx;
  uint32_t mask = 0;
  for (int p = 0; p < h->n_peers; ++p)
    if (h->recv_ptr[p + 1] > h->recv_ptr[p]) mask |= 1u << h->peers[p];
  w->wait_mask = mask;
  w->halo_flag = c->d_mbox->halo_flag;
  w->n_own = (uint32_t)n_own;
  for (int par = 0; par < 2; ++par) w->xg[par] = dfb_peer_landing(c, c->rank, par) - n_own * (uint64_t)blk_dim;
  w->tile_lo_end = 0;
  w->tile_hi_begin = 0xFFFFFFFFu;
}

DFB_API dfb_status dfb_halo_destroy(dfb_halo *h) {
  if (!h) return DFB_OK;
  cudaStreamSynchronize(h->ctx->stream);
  cudaStreamSynchronize(h->ctx->stream_comm);
  dfb_dev_free(h->d_send_idx);
  dfb_dev_free(h->d_sendbuf);
  delete h;
  return DFB_OK;
}

dfb_status dfb_halo_exchange_on(dfb_halo *h, double *d_vec, uint64_t n_blk, int blk_dim, cudaStream_t strm) {
  dfb_ctx *c = h->ctx;
  if (c->nranks <= 1 || h->n_peers == 0) return DFB_OK;
  DFB_REQUIRE(c->nccl_comm, "halo exchange requested but the context has no communicator");
  DFB_REQUIRE(blk_dim <= h->max_dim, "halo exchange: blk_dim too large");
  if (h->n_send > 0) {
    int g = dfb_div_up(h->n_send * blk_dim, 256);
    if (g > c->sm_count * 4) g = c->sm_count * 4;
    DFB_LAUNCH_ON(c, strm, k_halo_pack, g, 256, 0, d_vec, h->d_send_idx, h->n_send, blk_dim, h->d_sendbuf);
    DFB_CHECK_LAUNCH();
  }
  DFB_NCCL(g_nccl.GroupStart());
  for (int p = 0; p < h->n_peers; ++p) {
    const uint64_t ns = h->send_ptr[p + 1] - h->send_ptr[p], nr = h->recv_ptr[p + 1] - h->recv_ptr[p];
    if (ns) DFB_NCCL(g_nccl.Send(h->d_sendbuf + h->send_ptr[p] * blk_dim, ns * blk_dim, NCCL_FLOAT64, h->peers[p], c->nccl_comm, strm));
    if (nr) DFB_NCCL(g_nccl.Recv(d_vec + (n_blk + h->recv_ptr[p]) * blk_dim, nr * blk_dim, NCCL_FLOAT64, h->peers[p], c->nccl_comm, strm));
  }
  DFB_NCCL(g_nccl.GroupEnd());
  return DFB_OK;
}

DFB_API dfb_status dfb_halo_exchange(dfb_halo *h, dfb_vec *v) {
  DFB_REQUIRE(h && v, "dfb_halo_exchange: NULL argument");
  DFB_REQUIRE(v->n_ghost >= h->recv_ptr.back(), "dfb_halo_exchange: vector has fewer ghost blocks than the halo receives");
  return dfb_halo_exchange_on(h, v->d, v->n_blk, v->blk_dim, h->ctx->stream);
}
