// mesh_pattern.cu — device meshes and the sparsity pattern.
//   * dfb_mesh_*: device-resident elem2vtx (u32) + vtx2xyz, synthetic Kuhn tet grid generator
//   * dfb_pattern_*: row2idx/idx2col of sparse_square::Matrix (del-fem-cpu/src/sparse_square.rs:75-81)
//   * dfb_pattern_from_uniform_mesh: del_msh_cpu::vtx2vtx::from_uniform_mesh restated on the device:
//     vtx2elem (elements ascending per vertex) -> per vertex, elements ascending x nodes in order, first-seen order.
//     One thread per vertex; bit-identical to the serial construction because each row is built by one thread.
#include "dfb_internal.h"

// ------------------------------------------------------------------------------------------------ mesh
__global__ void k_narrow_u64(const unsigned long long *__restrict__ in, uint32_t *__restrict__ out, uint64_t n,
                             unsigned long long limit, int *err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = in[i];
    if (v >= limit) atomicExch(err, 1);
    out[i] = (uint32_t)v;
  }
}
__global__ void k_widen_u32(const uint32_t *__restrict__ in, unsigned long long *__restrict__ out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}
__global__ void k_check_u32(const uint32_t *__restrict__ in, uint64_t n, uint32_t limit, int *err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (in[i] >= limit) atomicExch(err, 1);
}

// upload a host usize array as device u32 (narrowing on the device; values must be < limit)
static dfb_status upload_narrow(dfb_ctx *c, const uint64_t *host, uint64_t n, uint64_t limit, uint32_t *d_out,
                                const char *what) {
  if (n == 0) return DFB_OK;
  unsigned long long *d_tmp = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_tmp, n));
  DFB_CUDA(cudaMemcpyAsync(d_tmp, host, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
  DFB_LAUNCH(c, k_narrow_u64, c->sm_count * 8, 256, 0, d_tmp, d_out, n, (unsigned long long)limit, c->d_errflag + 0);
  DFB_CHECK_LAUNCH();
  dfb_status st = dfb_check_error_flag(c, 0, DFB_ERR_BAD_ARG, what);
  dfb_dev_free(d_tmp);
  return st;
}

static dfb_status mesh_alloc(dfb_ctx *ctx, uint64_t num_elem, int num_node, uint64_t num_vtx, int ndim, dfb_mesh **out) {
  DFB_REQUIRE(ctx && out, "dfb_mesh_create: NULL argument");
  DFB_REQUIRE(num_node >= 2 && num_node <= 8, "dfb_mesh_create: num_node %d not in 2..8", num_node);
  DFB_REQUIRE(ndim == 2 || ndim == 3, "dfb_mesh_create: ndim %d not in {2,3}", ndim);
  if (num_vtx >= 0xFFFFFFF0ull || num_elem * (uint64_t)(num_node * num_node) >= 0xFFFFFFF0ull)
    return dfb_fail(DFB_ERR_INDEX_OVERFLOW, "dfb_mesh_create: mesh too large for u32 device indices");
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_mesh *m = new dfb_mesh();
  m->ctx = ctx;
  m->num_elem = num_elem;
  m->num_vtx = num_vtx;
  m->num_node = num_node;
  m->ndim = ndim;
  m->d_elem2vtx = nullptr;
  m->d_xyz = nullptr;
  dfb_status st = dfb_dev_alloc_t(&m->d_elem2vtx, num_elem * num_node + 4);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&m->d_xyz, num_vtx * ndim + 4);
  if (st != DFB_OK) {
    dfb_dev_free(m->d_elem2vtx);
    delete m;
    return st;
  }
  *out = m;
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_create(dfb_ctx *ctx, const uint64_t *elem2vtx, uint64_t num_elem, int32_t num_node,
                                   const double *vtx2xyz, uint64_t num_vtx, int32_t ndim, dfb_mesh **out) {
  DFB_REQUIRE(elem2vtx && vtx2xyz, "dfb_mesh_create: NULL array");
  dfb_mesh *m = nullptr;
  DFB_TRY(mesh_alloc(ctx, num_elem, num_node, num_vtx, ndim, &m));
  dfb_status st = upload_narrow(ctx, elem2vtx, num_elem * num_node, num_vtx, m->d_elem2vtx,
                                "dfb_mesh_create: elem2vtx entry >= num_vtx");
  if (st == DFB_OK) {
    cudaError_t e = cudaMemcpyAsync(m->d_xyz, vtx2xyz, num_vtx * ndim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_mesh_create: %s", cudaGetErrorString(e));
  }
  if (st != DFB_OK) {
    dfb_mesh_destroy(m);
    return st;
  }
  *out = m;
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_create_u32(dfb_ctx *ctx, const uint32_t *elem2vtx, uint64_t num_elem, int32_t num_node,
                                       const double *vtx2xyz, uint64_t num_vtx, int32_t ndim, dfb_mesh **out) {
  DFB_REQUIRE(elem2vtx && vtx2xyz, "dfb_mesh_create_u32: NULL array");
  dfb_mesh *m = nullptr;
  DFB_TRY(mesh_alloc(ctx, num_elem, num_node, num_vtx, ndim, &m));
  cudaError_t e = cudaMemcpyAsync(m->d_elem2vtx, elem2vtx, num_elem * num_node * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_xyz, vtx2xyz, num_vtx * ndim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  dfb_status st = DFB_OK;
  if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_mesh_create_u32: %s", cudaGetErrorString(e));
  if (st == DFB_OK) {
    DFB_LAUNCH(ctx, k_check_u32, ctx->sm_count * 8, 256, 0, m->d_elem2vtx, num_elem * num_node, (uint32_t)num_vtx, ctx->d_errflag + 0);
    st = dfb_check_error_flag(ctx, 0, DFB_ERR_BAD_ARG, "dfb_mesh_create_u32: elem2vtx entry >= num_vtx");
  }
  if (st != DFB_OK) {
    dfb_mesh_destroy(m);
    return st;
  }
  *out = m;
  return DFB_OK;
}

// Kuhn-6 grid: per cell with base corner o, for each axis permutation (lexicographic) tet = [o, o+e_p1, o+e_p1+e_p2,
// o+e_x+e_y+e_z], middle nodes swapped for odd permutations (positive volume).
__global__ void k_kuhn_elems(uint32_t *__restrict__ e2v, uint64_t nx, uint64_t ny, uint64_t nzl) {
  const uint64_t ncell = (nx - 1) * (ny - 1) * (nzl - 1);
  const int perm[6][2] = {{0, 1}, {0, 2}, {1, 0}, {1, 2}, {2, 0}, {2, 1}};
  const int odd[6] = {0, 1, 1, 0, 0, 1};
  const uint64_t stride[3] = {1, nx, nx * ny};
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ncell * 6; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t cell = q / 6;
    const int t = (int)(q - cell * 6);
    const uint64_t i = cell % (nx - 1);
    const uint64_t j = (cell / (nx - 1)) % (ny - 1);
    const uint64_t k = cell / ((nx - 1) * (ny - 1));
    const uint64_t o = i + nx * (j + ny * k);
    uint64_t a = o + stride[perm[t][0]];
    uint64_t b = a + stride[perm[t][1]];
    const uint64_t c = o + stride[0] + stride[1] + stride[2];
    if (odd[t]) {
      const uint64_t s = a;
      a = b;
      b = s;
    }
    reinterpret_cast<uint4 *>(e2v)[q] = make_uint4((uint32_t)o, (uint32_t)a, (uint32_t)b, (uint32_t)c);
  }
}

__global__ void k_kuhn_xyz(double *__restrict__ xyz, uint64_t nx, uint64_t ny, uint64_t nzl, uint64_t k_begin, double h) {
  const uint64_t nv = nx * ny * nzl;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = v % nx, j = (v / nx) % ny, k = v / (nx * ny) + k_begin;
    xyz[v * 3 + 0] = __dmul_rn((double)i, h);
    xyz[v * 3 + 1] = __dmul_rn((double)j, h);
    xyz[v * 3 + 2] = __dmul_rn((double)k, h);
  }
}

DFB_API dfb_status dfb_mesh_create_kuhn(dfb_ctx *ctx, uint64_t nx, uint64_t ny, uint64_t nz, double h, uint64_t k_begin,
                                        uint64_t k_end, dfb_mesh **out) {
  DFB_REQUIRE(nx >= 2 && ny >= 2 && nz >= 2, "dfb_mesh_create_kuhn: grid must be at least 2x2x2");
  DFB_REQUIRE(k_begin < k_end && k_end <= nz && k_end - k_begin >= 2, "dfb_mesh_create_kuhn: bad plane range");
  const uint64_t nzl = k_end - k_begin;
  dfb_mesh *m = nullptr;
  DFB_TRY(mesh_alloc(ctx, 6 * (nx - 1) * (ny - 1) * (nzl - 1), 4, nx * ny * nzl, 3, &m));
  DFB_LAUNCH(ctx, k_kuhn_elems, ctx->sm_count * 8, 256, 0, m->d_elem2vtx, nx, ny, nzl);
  DFB_LAUNCH(ctx, k_kuhn_xyz, ctx->sm_count * 8, 256, 0, m->d_xyz, nx, ny, nzl, k_begin, h);
  DFB_CHECK_LAUNCH();
  *out = m;
  return DFB_OK;
}

__global__ void k_rotate_ids(uint32_t *__restrict__ e2v, uint64_t n_entries, uint32_t shift, uint32_t nv) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_entries; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = e2v[q];
    e2v[q] = (v >= shift) ? (v - shift) : (v + nv - shift);
  }
}
__global__ void k_rotate_xyz(const double *__restrict__ in, double *__restrict__ out, uint64_t nv, int ndim, uint64_t shift) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv * ndim; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t v = q / ndim;
    const int d = (int)(q - v * ndim);
    const uint64_t nvn = (v >= shift) ? (v - shift) : (v + nv - shift);
    out[nvn * ndim + d] = in[q];
  }
}

DFB_API dfb_status dfb_mesh_rotate_vertex_ids(dfb_mesh *m, uint64_t shift) {
  DFB_REQUIRE(m, "dfb_mesh_rotate_vertex_ids: NULL argument");
  DFB_REQUIRE(shift <= m->num_vtx, "dfb_mesh_rotate_vertex_ids: shift > num_vtx");
  if (shift == 0 || shift == m->num_vtx) return DFB_OK;
  dfb_ctx *c = m->ctx;
  double *d_new = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_new, m->num_vtx * m->ndim + 4));
  DFB_LAUNCH(c, k_rotate_ids, c->sm_count * 8, 256, 0, m->d_elem2vtx, m->num_elem * m->num_node, (uint32_t)shift, (uint32_t)m->num_vtx);
  DFB_LAUNCH(c, k_rotate_xyz, c->sm_count * 8, 256, 0, m->d_xyz, d_new, m->num_vtx, m->ndim, shift);
  DFB_CHECK_LAUNCH();
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  dfb_dev_free(m->d_xyz);
  m->d_xyz = d_new;
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_destroy(dfb_mesh *m) {
  if (!m) return DFB_OK;
  cudaStreamSynchronize(m->ctx->stream);
  dfb_dev_free(m->d_elem2vtx);
  dfb_dev_free(m->d_xyz);
  delete m;
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_update_xyz(dfb_mesh *m, const double *vtx2xyz) {
  DFB_REQUIRE(m && vtx2xyz, "dfb_mesh_update_xyz: NULL argument");
  DFB_CUDA(cudaMemcpyAsync(m->d_xyz, vtx2xyz, m->num_vtx * m->ndim * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
  DFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_sizes(const dfb_mesh *m, uint64_t *num_elem, int32_t *num_node, uint64_t *num_vtx, int32_t *ndim) {
  DFB_REQUIRE(m, "dfb_mesh_sizes: NULL argument");
  if (num_elem) *num_elem = m->num_elem;
  if (num_node) *num_node = m->num_node;
  if (num_vtx) *num_vtx = m->num_vtx;
  if (ndim) *ndim = m->ndim;
  return DFB_OK;
}

static dfb_status download_widen(dfb_ctx *c, const uint32_t *d_in, uint64_t n, uint64_t *host) {
  if (n == 0) return DFB_OK;
  unsigned long long *d_tmp = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_tmp, n));
  DFB_LAUNCH(c, k_widen_u32, c->sm_count * 8, 256, 0, d_in, d_tmp, n);
  DFB_CHECK_LAUNCH();
  DFB_CUDA(cudaMemcpyAsync(host, d_tmp, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  dfb_dev_free(d_tmp);
  return DFB_OK;
}

DFB_API dfb_status dfb_mesh_download(const dfb_mesh *m, uint64_t *elem2vtx, double *vtx2xyz) {
  DFB_REQUIRE(m, "dfb_mesh_download: NULL argument");
  if (elem2vtx) DFB_TRY(download_widen(m->ctx, m->d_elem2vtx, m->num_elem * m->num_node, elem2vtx));
  if (vtx2xyz) {
    DFB_CUDA(cudaMemcpyAsync(vtx2xyz, m->d_xyz, m->num_vtx * m->ndim * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    DFB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  }
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ pattern
__global__ void k_check_rowptr(const uint32_t *__restrict__ row2idx, uint64_t num_blk, uint64_t nnz, int *err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_blk; i += (uint64_t)gridDim.x * blockDim.x)
    if (row2idx[i] > row2idx[i + 1] || row2idx[i + 1] > nnz) atomicExch(err, 1);
}

static dfb_status pattern_alloc(dfb_ctx *ctx, uint64_t num_blk, uint64_t nnz, int convention, dfb_pattern **out) {
  DFB_REQUIRE(ctx && out, "dfb_pattern_create: NULL argument");
  DFB_REQUIRE(convention == 0 || convention == 1, "dfb_pattern_create: convention %d not in {0,1}", convention);
  if (num_blk >= 0xFFFFFFF0ull || nnz + num_blk >= 0xFFFFFFF0ull)
    return dfb_fail(DFB_ERR_INDEX_OVERFLOW, "dfb_pattern_create: pattern too large for u32 device indices");
  DFB_CUDA(cudaSetDevice(ctx->device));
  dfb_pattern *p = new dfb_pattern();
  p->ctx = ctx;
  p->num_blk = num_blk;
  p->nnz = nnz;
  p->convention = convention;
  p->refcount = 1;
  p->d_row2idx = nullptr;
  p->d_idx2col = nullptr;
  dfb_status st = dfb_dev_alloc_t(&p->d_row2idx, num_blk + 2);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&p->d_idx2col, nnz + 8);
  if (st != DFB_OK) {
    dfb_dev_free(p->d_row2idx);
    delete p;
    return st;
  }
  *out = p;
  return DFB_OK;
}

DFB_API dfb_status dfb_pattern_create(dfb_ctx *ctx, const uint64_t *row2idx, const uint64_t *idx2col, uint64_t num_blk,
                                      int32_t convention, dfb_pattern **out) {
  DFB_REQUIRE(row2idx, "dfb_pattern_create: row2idx is NULL");
  const uint64_t nnz = row2idx[num_blk];
  DFB_REQUIRE(row2idx[0] == 0, "dfb_pattern_create: row2idx[0] != 0");
  DFB_REQUIRE(nnz == 0 || idx2col, "dfb_pattern_create: idx2col is NULL");
  dfb_pattern *p = nullptr;
  DFB_TRY(pattern_alloc(ctx, num_blk, nnz, convention, &p));
  dfb_status st = upload_narrow(ctx, row2idx, num_blk + 1, nnz + 1, p->d_row2idx, "dfb_pattern_create: row2idx entry > nnz");
  // NOTE: the column range check is against the number of rows; distributed matrices (ghost columns) are built with
  // dfb_pattern_create_u32 / dfb_pattern_from_uniform_mesh which take the column bound separately.
  if (st == DFB_OK) st = upload_narrow(ctx, idx2col, nnz, num_blk, p->d_idx2col, "dfb_pattern_create: idx2col entry >= num_blk");
  if (st == DFB_OK) {
    DFB_LAUNCH(ctx, k_check_rowptr, ctx->sm_count * 4, 256, 0, p->d_row2idx, num_blk, nnz, ctx->d_errflag + 0);
    st = dfb_check_error_flag(ctx, 0, DFB_ERR_BAD_ARG, "dfb_pattern_create: row2idx is not non-decreasing");
  }
  if (st != DFB_OK) {
    dfb_pattern_destroy(p);
    return st;
  }
  *out = p;
  return DFB_OK;
}

DFB_API dfb_status dfb_pattern_create_u32(dfb_ctx *ctx, const uint32_t *row2idx, const uint32_t *idx2col, uint64_t num_blk,
                                          int32_t convention, dfb_pattern **out) {
  DFB_REQUIRE(row2idx, "dfb_pattern_create_u32: row2idx is NULL");
  const uint64_t nnz = row2idx[num_blk];
  DFB_REQUIRE(row2idx[0] == 0, "dfb_pattern_create_u32: row2idx[0] != 0");
  DFB_REQUIRE(nnz == 0 || idx2col, "dfb_pattern_create_u32: idx2col is NULL");
  dfb_pattern *p = nullptr;
  DFB_TRY(pattern_alloc(ctx, num_blk, nnz, convention, &p));
  cudaError_t e = cudaMemcpyAsync(p->d_row2idx, row2idx, (num_blk + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(p->d_idx2col, idx2col, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
  dfb_status st = DFB_OK;
  if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_pattern_create_u32: %s", cudaGetErrorString(e));
  if (st == DFB_OK) {
    DFB_LAUNCH(ctx, k_check_rowptr, ctx->sm_count * 4, 256, 0, p->d_row2idx, num_blk, nnz, ctx->d_errflag + 0);
    st = dfb_check_error_flag(ctx, 0, DFB_ERR_BAD_ARG, "dfb_pattern_create_u32: row2idx is not non-decreasing");
  }
  if (st != DFB_OK) {
    dfb_pattern_destroy(p);
    return st;
  }
  *out = p;
  return DFB_OK;
}

DFB_API dfb_status dfb_pattern_destroy(dfb_pattern *p) {
  if (!p) return DFB_OK;
  if (--p->refcount > 0) return DFB_OK;
  cudaStreamSynchronize(p->ctx->stream);
  dfb_dev_free(p->d_row2idx);
  dfb_dev_free(p->d_idx2col);
  delete p;
  return DFB_OK;
}

DFB_API dfb_status dfb_pattern_sizes(const dfb_pattern *p, uint64_t *num_blk, uint64_t *nnz, int32_t *convention) {
  DFB_REQUIRE(p, "dfb_pattern_sizes: NULL argument");
  if (num_blk) *num_blk = p->num_blk;
  if (nnz) *nnz = p->nnz;
  if (convention) *convention = p->convention;
  return DFB_OK;
}

DFB_API dfb_status dfb_pattern_download(const dfb_pattern *p, uint64_t *row2idx, uint64_t *idx2col) {
  DFB_REQUIRE(p, "dfb_pattern_download: NULL argument");
  if (row2idx) DFB_TRY(download_widen(p->ctx, p->d_row2idx, p->num_blk + 1, row2idx));
  if (idx2col) DFB_TRY(download_widen(p->ctx, p->d_idx2col, p->nnz, idx2col));
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ vtx2vtx on device
__global__ void k_count_vtx2elem(const uint32_t *__restrict__ e2v, uint64_t n_entries, uint32_t *__restrict__ cnt) {
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_entries; q += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&cnt[e2v[q]], 1u);
}

__global__ void k_fill_vtx2elem(const uint32_t *__restrict__ e2v, uint64_t num_elem, int num_node,
                                const uint32_t *__restrict__ ptr, uint32_t *__restrict__ cursor, uint32_t *__restrict__ v2e) {
  const uint64_t n_entries = num_elem * num_node;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_entries; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = e2v[q];
    const uint32_t pos = atomicAdd(&cursor[v], 1u);
    v2e[ptr[v] + pos] = (uint32_t)(q / num_node);
  }
}

// ascending insertion sort of each vertex's element list (restores the counting-sort order of the serial algorithm)
__global__ void k_sort_lists(const uint32_t *__restrict__ ptr, uint32_t *__restrict__ items, uint64_t n_lists) {
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_lists; v += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = ptr[v], e = ptr[v + 1];
    for (uint32_t i = b + 1; i < e; ++i) {
      const uint32_t key = items[i];
      uint32_t j = i;
      while (j > b && items[j - 1] > key) {
        items[j] = items[j - 1];
        --j;
      }
      items[j] = key;
    }
  }
}

// pass 0: count distinct neighbours per vertex; pass 1: emit them in first-seen order.
// Distinctness is checked against the entries this thread has already emitted / counted (local list), which is
// equivalent to the serial algorithm's stamp array.
template <bool EMIT>
__global__ void k_vtx2vtx(const uint32_t *__restrict__ e2v, int num_node, const uint32_t *__restrict__ v2e_ptr,
                          const uint32_t *__restrict__ v2e, int is_self, uint64_t num_rows, uint32_t *__restrict__ row2idx,
                          uint32_t *__restrict__ idx2col, uint32_t *__restrict__ scratch, uint32_t scratch_stride) {
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < num_rows; v += (uint64_t)gridDim.x * blockDim.x) {
    // EMIT: write into idx2col[row2idx[v]...]; COUNT: need a temporary list: use the per-thread scratch row
    uint32_t *list = EMIT ? (idx2col + row2idx[v]) : (scratch + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * scratch_stride);
    uint32_t n = 0;
    bool overflow = false;
    for (uint32_t q = v2e_ptr[v]; q < v2e_ptr[v + 1]; ++q) {
      const uint32_t e = v2e[q];
      for (int k = 0; k < num_node; ++k) {
        const uint32_t j = e2v[(uint64_t)e * num_node + k];
        if (!is_self && j == (uint32_t)v) continue;
        bool seen = false;
        for (uint32_t t = 0; t < n; ++t)
          if (list[t] == j) {
            seen = true;
            break;
          }
        if (seen) continue;
        if (!EMIT && n >= scratch_stride) {
          overflow = true;
          continue;
        }
        list[n++] = j;
      }
    }
    if (!EMIT) row2idx[v] = overflow ? 0xFFFFFFFFu : n;
  }
}

DFB_API dfb_status dfb_pattern_from_uniform_mesh(dfb_ctx *ctx, const dfb_mesh *mesh, int32_t is_self, uint64_t num_rows,
                                                 dfb_pattern **out) {
  DFB_REQUIRE(ctx && mesh && out, "dfb_pattern_from_uniform_mesh: NULL argument");
  DFB_REQUIRE(num_rows <= mesh->num_vtx, "dfb_pattern_from_uniform_mesh: num_rows > num_vtx");
  DFB_CUDA(cudaSetDevice(ctx->device));
  const uint64_t nv = mesh->num_vtx, n_entries = mesh->num_elem * mesh->num_node;
  uint32_t *d_ptr = nullptr, *d_cur = nullptr, *d_v2e = nullptr, *d_scratch = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_ptr, nv + 2));
  DFB_TRY(dfb_dev_alloc_t(&d_cur, nv + 2));
  DFB_TRY(dfb_dev_alloc_t(&d_v2e, n_entries + 2));
  DFB_CUDA(cudaMemsetAsync(d_ptr, 0, (nv + 2) * sizeof(uint32_t), ctx->stream));
  DFB_CUDA(cudaMemsetAsync(d_cur, 0, (nv + 2) * sizeof(uint32_t), ctx->stream));
  const int G = ctx->sm_count * 8;
  DFB_LAUNCH(ctx, k_count_vtx2elem, G, 256, 0, mesh->d_elem2vtx, n_entries, d_ptr);
  DFB_CHECK_LAUNCH();
  DFB_TRY(dfb_exclusive_scan_u32(ctx, d_ptr, nv, nullptr));
  DFB_LAUNCH(ctx, k_fill_vtx2elem, G, 256, 0, mesh->d_elem2vtx, mesh->num_elem, mesh->num_node, d_ptr, d_cur, d_v2e);
  DFB_LAUNCH(ctx, k_sort_lists, G, 256, 0, d_ptr, d_v2e, nv);
  DFB_CHECK_LAUNCH();
  // count pass with a bounded per-thread scratch list (128 neighbours); rows beyond that are rejected
  const uint32_t stride = 128;
  const int Gc = ctx->sm_count * 2, Tc = 128;
  DFB_TRY(dfb_dev_alloc_t(&d_scratch, (uint64_t)Gc * Tc * stride));
  uint32_t *d_row2idx = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_row2idx, num_rows + 2));
  DFB_LAUNCH(ctx, k_vtx2vtx<false>, Gc, Tc, 0, mesh->d_elem2vtx, mesh->num_node, d_ptr, d_v2e, (int)is_self, num_rows,
             d_row2idx, (uint32_t *)nullptr, d_scratch, stride);
  DFB_CHECK_LAUNCH();
  uint64_t nnz = 0;
  dfb_status st = dfb_exclusive_scan_u32(ctx, d_row2idx, num_rows, &nnz);
  dfb_pattern *p = nullptr;
  if (st == DFB_OK) st = pattern_alloc(ctx, num_rows, nnz, is_self ? 1 : 0, &p);
  else st = dfb_fail(DFB_ERR_UNSUPPORTED, "dfb_pattern_from_uniform_mesh: a vertex has more than 128 neighbours");
  if (st == DFB_OK) {
    cudaMemcpyAsync(p->d_row2idx, d_row2idx, (num_rows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream);
    DFB_LAUNCH(ctx, k_vtx2vtx<true>, G, 128, 0, mesh->d_elem2vtx, mesh->num_node, d_ptr, d_v2e, (int)is_self, num_rows,
               p->d_row2idx, p->d_idx2col, (uint32_t *)nullptr, 0u);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_pattern_from_uniform_mesh: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(ctx->stream);
  dfb_dev_free(d_ptr);
  dfb_dev_free(d_cur);
  dfb_dev_free(d_v2e);
  dfb_dev_free(d_scratch);
  dfb_dev_free(d_row2idx);
  if (st != DFB_OK) {
    if (p) dfb_pattern_destroy(p);
    return st;
  }
  *out = p;
  return DFB_OK;
}
