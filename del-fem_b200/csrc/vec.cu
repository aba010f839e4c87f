// vec.cu — device vectors `[[T;N]]` and the level-1 operations of del-fem-cpu/src/slice_of_array.rs:1-47
// (del-fem-ls/src/slice.rs:4-55).  All kernels are HBM-bound streaming kernels: 128-bit loads, grid = multiple of
// the SM count, fp64 throughout.
#include "dfb_internal.h"
#include "reduce.cuh"

static inline int stream_grid(const dfb_ctx *c, uint64_t n_items, int per_thread, int threads) {
  uint64_t want = (n_items + (uint64_t)per_thread * threads - 1) / ((uint64_t)per_thread * threads);
  uint64_t cap = (uint64_t)c->sm_count * 8;
  if (want < 1) want = 1;
  if (want > cap) want = cap;
  if (want > DFB_RED_MAX_BLOCKS) want = DFB_RED_MAX_BLOCKS;
  return (int)want;
}

DFB_API dfb_status dfb_vec_create(dfb_ctx *ctx, uint64_t n_blk, uint64_t n_ghost, int32_t blk_dim, dfb_vec **out) {
  DFB_REQUIRE(ctx && out, "dfb_vec_create: NULL argument");
  DFB_REQUIRE(blk_dim >= 1 && blk_dim <= 4, "dfb_vec_create: blk_dim %d not in 1..4", blk_dim);
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_vec *v = new dfb_vec();
  v->ctx = ctx;
  v->n_blk = n_blk;
  v->n_ghost = n_ghost;
  v->blk_dim = blk_dim;
  dfb_status st = dfb_dev_alloc_t(&v->d, v->len_all() + 8);
  if (st != DFB_OK) {
    delete v;
    return st;
  }
  DFB_CUDA(cudaMemsetAsync(v->d, 0, (v->len_all() + 8) * sizeof(double), ctx->stream));
  *out = v;
  return DFB_OK;
}

DFB_API dfb_status dfb_vec_destroy(dfb_vec *v) {
  if (!v) return DFB_OK;
  cudaStreamSynchronize(v->ctx->stream);
  dfb_dev_free(v->d);
  delete v;
  return DFB_OK;
}

DFB_API double *dfb_vec_device_ptr(dfb_vec *v) { return v ? v->d : nullptr; }

DFB_API dfb_status dfb_vec_upload(dfb_vec *v, const double *host) {
  DFB_REQUIRE(v && host, "dfb_vec_upload: NULL argument");
  DFB_CUDA(cudaMemcpyAsync(v->d, host, v->len() * sizeof(double), cudaMemcpyHostToDevice, v->ctx->stream));
  DFB_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return DFB_OK;
}

DFB_API dfb_status dfb_vec_download(dfb_vec *v, double *host) {
  DFB_REQUIRE(v && host, "dfb_vec_download: NULL argument");
  DFB_CUDA(cudaMemcpyAsync(host, v->d, v->len() * sizeof(double), cudaMemcpyDeviceToHost, v->ctx->stream));
  DFB_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return DFB_OK;
}

DFB_API dfb_status dfb_vec_set_zero(dfb_vec *v) {
  DFB_REQUIRE(v, "dfb_vec_set_zero: NULL argument");
  DFB_CUDA(cudaMemsetAsync(v->d, 0, v->len_all() * sizeof(double), v->ctx->stream));
  return DFB_OK;
}

static dfb_status same_shape(const dfb_vec *a, const dfb_vec *b, const char *who) {
  DFB_REQUIRE(a && b, "%s: NULL argument", who);
  DFB_REQUIRE(a->n_blk == b->n_blk && a->blk_dim == b->blk_dim, "%s: shape mismatch (%llu x %d vs %llu x %d)", who,
              (unsigned long long)a->n_blk, a->blk_dim, (unsigned long long)b->n_blk, b->blk_dim);
  DFB_REQUIRE(a->ctx == b->ctx, "%s: vectors belong to different contexts", who);
  return DFB_OK;
}

DFB_API dfb_status dfb_vec_copy(dfb_vec *dst, const dfb_vec *src) {
  DFB_TRY(same_shape(dst, src, "dfb_vec_copy"));
  DFB_CUDA(cudaMemcpyAsync(dst->d, src->d, src->len() * sizeof(double), cudaMemcpyDeviceToDevice, dst->ctx->stream));
  return DFB_OK;
}

// ---- dot
__global__ void __launch_bounds__(256) k_dot(const double *__restrict__ a, const double *__restrict__ b, uint64_t n,
                                             double *partials, unsigned int *ticket, double *out) {
  __shared__ double s_buf[32];
  double acc[1] = {0.0};
  const uint64_t n2 = n >> 1;
  const double2 *a2 = reinterpret_cast<const double2 *>(a);
  const double2 *b2 = reinterpret_cast<const double2 *>(b);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    const double2 x = a2[i], y = b2[i];
    acc[0] += x.x * y.x + x.y * y.y;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) acc[0] += a[n - 1] * b[n - 1];
  if (grid_sum_last_block<1>(acc, partials, ticket, s_buf) && threadIdx.x == 0) out[0] = acc[0];
}

DFB_API dfb_status dfb_vec_dot(const dfb_vec *a, const dfb_vec *b, double *out) {
  DFB_TRY(same_shape(a, b, "dfb_vec_dot"));
  DFB_REQUIRE(out, "dfb_vec_dot: out is NULL");
  dfb_ctx *c = a->ctx;
  const uint64_t n = a->len();
  const int grid = stream_grid(c, n, 8, 256);
  DFB_LAUNCH(c, k_dot, grid, 256, 0, a->d, b->d, n, c->d_partials, c->d_ticket + 0, c->d_scratch);
  DFB_CHECK_LAUNCH();
  if (c->nranks > 1) DFB_TRY(dfb_nccl_allreduce_sum(c, c->d_scratch, 1, c->stream));
  double *h = (double *)c->h_pinned;
  DFB_CUDA(cudaMemcpyAsync(h, c->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  *out = *h;
  return DFB_OK;
}

// ---- u += alpha p
__global__ void __launch_bounds__(256) k_axpy(double *__restrict__ u, double alpha, const double *__restrict__ p, uint64_t n) {
  const uint64_t n2 = n >> 1;
  double2 *u2 = reinterpret_cast<double2 *>(u);
  const double2 *p2 = reinterpret_cast<const double2 *>(p);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 x = u2[i];
    const double2 y = p2[i];
    x.x += alpha * y.x;
    x.y += alpha * y.y;
    u2[i] = x;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) u[n - 1] += alpha * p[n - 1];
}

DFB_API dfb_status dfb_vec_add_scaled(dfb_vec *u, double alpha, const dfb_vec *p) {
  DFB_TRY(same_shape(u, p, "dfb_vec_add_scaled"));
  dfb_ctx *c = u->ctx;
  DFB_LAUNCH(c, k_axpy, stream_grid(c, u->len(), 4, 256), 256, 0, u->d, alpha, p->d, u->len());
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ---- p = r + beta p
__global__ void __launch_bounds__(256) k_xpby(double *__restrict__ p, double beta, const double *__restrict__ r, uint64_t n) {
  const uint64_t n2 = n >> 1;
  double2 *p2 = reinterpret_cast<double2 *>(p);
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
    double2 x = p2[i];
    const double2 y = r2[i];
    x.x = y.x + beta * x.x;
    x.y = y.y + beta * x.y;
    p2[i] = x;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) p[n - 1] = r[n - 1] + beta * p[n - 1];
}

DFB_API dfb_status dfb_vec_scale_and_add(dfb_vec *p, double beta, const dfb_vec *r) {
  DFB_TRY(same_shape(p, r, "dfb_vec_scale_and_add"));
  dfb_ctx *c = p->ctx;
  DFB_LAUNCH(c, k_xpby, stream_grid(c, p->len(), 4, 256), 256, 0, p->d, beta, r->d, p->len());
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ---- v *= s
__global__ void __launch_bounds__(256) k_scale(double *__restrict__ v, double s, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] *= s;
}

DFB_API dfb_status dfb_vec_scale(dfb_vec *v, double s) {
  DFB_REQUIRE(v, "dfb_vec_scale: NULL argument");
  dfb_ctx *c = v->ctx;
  DFB_LAUNCH(c, k_scale, stream_grid(c, v->len(), 4, 256), 256, 0, v->d, s, v->len());
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ---- rhs[d] = 0 where flag
__global__ void k_zero_flagged(double *__restrict__ v, const int32_t *__restrict__ flag, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (flag[i] != 0) v[i] = 0.0;
}

DFB_API dfb_status dfb_vec_set_fixed_dof_zero(dfb_vec *rhs, const int32_t *flag_host) {
  DFB_REQUIRE(rhs && flag_host, "dfb_vec_set_fixed_dof_zero: NULL argument");
  dfb_ctx *c = rhs->ctx;
  int32_t *d_flag = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_flag, rhs->len()));
  DFB_CUDA(cudaMemcpyAsync(d_flag, flag_host, rhs->len() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  DFB_LAUNCH(c, k_zero_flagged, stream_grid(c, rhs->len(), 4, 256), 256, 0, rhs->d, d_flag, rhs->len());
  DFB_CHECK_LAUNCH();
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  dfb_dev_free(d_flag);
  return DFB_OK;
}

DFB_API dfb_status dfb_bcflag_create(dfb_ctx *ctx, const int32_t *flag_host, uint64_t n_blk_all, int32_t blk_dim, dfb_bcflag **out) {
  DFB_REQUIRE(ctx && flag_host && out, "dfb_bcflag_create: NULL argument");
  DFB_REQUIRE(blk_dim >= 1 && blk_dim <= 4, "dfb_bcflag_create: blk_dim %d not in 1..4", blk_dim);
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_bcflag *f = new dfb_bcflag();
  f->ctx = ctx;
  f->n_blk_all = n_blk_all;
  f->blk_dim = blk_dim;
  f->d = nullptr;
  dfb_status st = dfb_dev_alloc_t(&f->d, n_blk_all * blk_dim + 4);
  if (st == DFB_OK) {
    cudaError_t e = cudaMemcpyAsync(f->d, flag_host, n_blk_all * blk_dim * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_bcflag_create: %s", cudaGetErrorString(e));
  }
  if (st != DFB_OK) {
    dfb_dev_free(f->d);
    delete f;
    return st;
  }
  *out = f;
  return DFB_OK;
}

DFB_API dfb_status dfb_bcflag_destroy(dfb_bcflag *f) {
  if (!f) return DFB_OK;
  cudaStreamSynchronize(f->ctx->stream);
  dfb_dev_free(f->d);
  delete f;
  return DFB_OK;
}

DFB_API dfb_status dfb_vec_set_fixed_dof_zero_dev(dfb_vec *rhs, const dfb_bcflag *f) {
  DFB_REQUIRE(rhs && f, "dfb_vec_set_fixed_dof_zero_dev: NULL argument");
  DFB_REQUIRE(f->blk_dim == rhs->blk_dim && f->n_blk_all >= rhs->n_blk, "dfb_vec_set_fixed_dof_zero_dev: flag shape mismatch");
  dfb_ctx *c = rhs->ctx;
  DFB_LAUNCH(c, k_zero_flagged, stream_grid(c, rhs->len(), 4, 256), 256, 0, rhs->d, f->d, rhs->len());
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ---- synthetic fill (bit-identical to oracle orc_splitmix_fill)
__global__ void k_fill_splitmix(double *v, uint64_t n, uint64_t seed, uint64_t first, double lo, double hi) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned long long z = (seed + first + i + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    const double u = __dmul_rn((double)(z >> 11), 1.0 / 9007199254740992.0);
    v[i] = __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u));
  }
}

DFB_API dfb_status dfb_vec_fill_splitmix(dfb_vec *v, uint64_t seed, uint64_t first_idx, double lo, double hi) {
  DFB_REQUIRE(v, "dfb_vec_fill_splitmix: NULL argument");
  dfb_ctx *c = v->ctx;
  DFB_LAUNCH(c, k_fill_splitmix, stream_grid(c, v->len(), 4, 256), 256, 0, v->d, v->len(), seed, first_idx, lo, hi);
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}
