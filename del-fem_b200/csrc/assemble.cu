// assemble.cu — element kernels + the merge as a deterministic, atomic-free gather.            (compiled --fmad=false)
//
// Reference being replaced: the serial element loop + scatter-add
//   emat = kernel(p0..);  Matrix::merge_for_array_blk(&emat, node2vtx, &mut col2idx)      del-fem-cpu/src/sparse_square.rs:180-214
//   merge::csrdia / merge::blkcsr                                                          del-fem-cpu/src/merge.rs:3-88
//   Matrix::merge                                                                          del-fem-ls/src/sparse_square.rs:66-104
// whose floating-point accumulation order into a given nonzero is "ascending element index".
//
// Here: dfb_plan_create inverts the element->nonzero map once per (pattern, mesh): for every nonzero slot (off-diagonal
// idx, or nnz+row for the separate diagonal) the list of contributions (e,i,j), sorted ascending.  The default assembly is the
// fused row-tile kernel of assemble_tile.cu; this file holds the plan, the batched element kernels of the two-phase path (energy
// models, user-supplied element matrices), and the slot-centric gathers (one thread per slot walks its list in that order and
// writes the sum once: no atomics, no set_zero pass, bit-reproducible) that serve as fallback and reference point.
// This translation unit is compiled with --fmad=false and the element formulas are written operation-for-operation like
// the reference's (no contraction in rustc either), so linear kernels reproduce the serial result bit-for-bit.
#include "dfb_internal.h"
#include "elements.cuh"
#include "reduce.cuh"

static const uint32_t NO_SLOT = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------ plan (index maps)
// slot of (row r, column vertex cv): last occurrence wins like the reference's col2idx scatter; NO_SLOT if absent
__device__ __forceinline__ uint32_t find_slot(const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col,
                                              uint32_t r, uint32_t cv) {
  uint32_t s = NO_SLOT;
  for (uint32_t k = row2idx[r]; k < row2idx[r + 1]; ++k)
    if (idx2col[k] == cv) s = k;
  return s;
}

// MODE 0: count contributions per slot; MODE 1: fill contribution lists; MODE 2: write the forward map elem2nz
template <int MODE>
__global__ void k_plan_pass(const uint32_t *__restrict__ e2v, uint64_t num_elem, int n_node, int flavour, int convention,
                            const uint32_t *__restrict__ row2idx, const uint32_t *__restrict__ idx2col, uint64_t num_rows,
                            uint64_t nnz, uint32_t *__restrict__ cnt_or_cursor, const uint32_t *__restrict__ nz2ptr,
                            uint32_t *__restrict__ out, int *err) {
  const uint64_t total = num_elem * n_node;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t e = q / n_node;
    const int i = (int)(q - e * n_node);
    const uint32_t *conn = e2v + e * n_node;
    const uint32_t r = conn[i];
    if (r >= num_rows) {  // row owned by another rank: not assembled here
      if (MODE == 2)
        for (int j = 0; j < n_node; ++j) out[q * n_node + j] = NO_SLOT;
      continue;
    }
    for (int j = 0; j < n_node; ++j) {
      const uint32_t cv = conn[j];
      const bool is_dia = convention == 0 && (flavour == 0 ? (i == j) : (r == cv));
      uint32_t slot;
      if (is_dia) slot = (uint32_t)nnz + r;
      else {
        slot = find_slot(row2idx, idx2col, r, cv);
        if (slot == NO_SLOT) {
          atomicExch(err, 1);
          if (MODE == 2) out[q * n_node + j] = NO_SLOT;
          continue;
        }
      }
      if (MODE == 0) atomicAdd(&cnt_or_cursor[slot], 1u);
      if (MODE == 1) {
        const uint32_t pos = atomicAdd(&cnt_or_cursor[slot], 1u);
        out[nz2ptr[slot] + pos] = (uint32_t)(q * n_node + j);  // code = (e*n_node + i)*n_node + j
      }
      if (MODE == 2) out[q * n_node + j] = slot;
    }
  }
}

static __global__ void k_sort_contrib(const uint32_t *__restrict__ ptr, uint32_t *__restrict__ items, uint64_t n_lists) {
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

DFB_API dfb_status dfb_plan_create(dfb_ctx *ctx, dfb_pattern *pattern, const dfb_mesh *mesh, int32_t flavour, dfb_plan **out) {
  DFB_REQUIRE(ctx && pattern && mesh && out, "dfb_plan_create: NULL argument");
  DFB_REQUIRE(flavour == 0 || flavour == 1, "dfb_plan_create: flavour %d not in {0,1}", flavour);
  DFB_REQUIRE(pattern->num_blk <= mesh->num_vtx, "dfb_plan_create: pattern has more rows than the mesh has vertices");
  DFB_CUDA(cudaSetDevice(ctx->device));
  const int n_node = mesh->num_node;
  const uint64_t num_slots = pattern->nnz + (pattern->convention == 0 ? pattern->num_blk : 0);
  dfb_plan *p = new dfb_plan();
  p->ctx = ctx;
  p->pat = pattern;
  pattern->refcount += 1;
  p->num_elem = mesh->num_elem;
  p->num_node = n_node;
  p->flavour = flavour;
  p->num_slots = num_slots;
  p->num_contrib = 0;
  p->d_nz2ptr = nullptr;
  p->d_contrib = nullptr;
  p->mesh = mesh;
  uint32_t *d_cursor = nullptr;
  dfb_status st = dfb_dev_alloc_t(&p->d_nz2ptr, num_slots + 2);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_cursor, num_slots + 2);
  const int G = ctx->sm_count * 8;
  if (st == DFB_OK) {
    cudaMemsetAsync(p->d_nz2ptr, 0, (num_slots + 2) * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(d_cursor, 0, (num_slots + 2) * sizeof(uint32_t), ctx->stream);
    DFB_LAUNCH(ctx, k_plan_pass<0>, G, 256, 0, mesh->d_elem2vtx, mesh->num_elem, n_node, flavour, pattern->convention,
               pattern->d_row2idx, pattern->d_idx2col, pattern->num_blk, pattern->nnz, p->d_nz2ptr, (const uint32_t *)nullptr,
               (uint32_t *)nullptr, ctx->d_errflag + 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_plan_create: %s", cudaGetErrorString(e));
  }
  if (st == DFB_OK)
    st = dfb_check_error_flag(ctx, 1, DFB_ERR_PATTERN_MISS,
                              "dfb_plan_create: an element couples two vertices whose (row, col) is not in the pattern "
                              "(assert_ne!(idx0, usize::MAX))");
  uint64_t total = 0;
  if (st == DFB_OK) st = dfb_exclusive_scan_u32(ctx, p->d_nz2ptr, num_slots, &total);
  if (st == DFB_OK) {
    p->num_contrib = total;
    st = dfb_dev_alloc_t(&p->d_contrib, total + 2);
  }
  if (st == DFB_OK) {
    DFB_LAUNCH(ctx, k_plan_pass<1>, G, 256, 0, mesh->d_elem2vtx, mesh->num_elem, n_node, flavour, pattern->convention,
               pattern->d_row2idx, pattern->d_idx2col, pattern->num_blk, pattern->nnz, d_cursor, p->d_nz2ptr, p->d_contrib,
               ctx->d_errflag + 1);
    DFB_LAUNCH(ctx, k_sort_contrib, G, 128, 0, p->d_nz2ptr, p->d_contrib, num_slots);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_plan_create: %s", cudaGetErrorString(e));
  }
  cudaStreamSynchronize(ctx->stream);
  dfb_dev_free(d_cursor);
  if (st != DFB_OK) {
    dfb_plan_destroy(p);
    return st;
  }
  *out = p;
  return DFB_OK;
}

DFB_API dfb_status dfb_plan_destroy(dfb_plan *p) {
  if (!p) return DFB_OK;
  cudaStreamSynchronize(p->ctx->stream);
  dfb_dev_free(p->d_nz2ptr);
  dfb_dev_free(p->d_contrib);
  dfb_dev_free(p->d_geo);
  dfb_plan_free_tiles(p);
  dfb_pattern_destroy(p->pat);
  delete p;
  return DFB_OK;
}

DFB_API dfb_status dfb_plan_sizes(const dfb_plan *p, uint64_t *num_slots, uint64_t *num_contrib) {
  DFB_REQUIRE(p, "dfb_plan_sizes: NULL argument");
  if (num_slots) *num_slots = p->num_slots;
  if (num_contrib) *num_contrib = p->num_contrib;
  return DFB_OK;
}

DFB_API dfb_status dfb_plan_download_elem2nz(const dfb_plan *p, uint32_t *elem2nz) {
  DFB_REQUIRE(p && elem2nz, "dfb_plan_download_elem2nz: NULL argument");
  dfb_ctx *c = p->ctx;
  const uint64_t n = p->num_elem * p->num_node * p->num_node;
  uint32_t *d_out = nullptr;
  DFB_TRY(dfb_dev_alloc_t(&d_out, n + 2));
  DFB_LAUNCH(c, k_plan_pass<2>, c->sm_count * 8, 256, 0, p->mesh->d_elem2vtx, p->num_elem, p->num_node, p->flavour,
             p->pat->convention, p->pat->d_row2idx, p->pat->d_idx2col, p->pat->num_blk, p->pat->nnz, (uint32_t *)nullptr,
             (const uint32_t *)nullptr, d_out, c->d_errflag + 1);
  DFB_CHECK_LAUNCH();
  DFB_CUDA(cudaMemcpyAsync(elem2nz, d_out, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  DFB_CUDA(cudaStreamSynchronize(c->stream));
  dfb_dev_free(d_out);
  DFB_CUDA(cudaMemsetAsync(c->d_errflag + 1, 0, sizeof(int), c->stream));
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ fused gather-assembly
template <int KIND>
__global__ void __launch_bounds__(128) k_assemble_fused(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib,
                                                        uint64_t num_slots, uint64_t nnz, const uint32_t *__restrict__ e2v,
                                                        const double *__restrict__ xyz, double p0, double p1,
                                                        double *__restrict__ idx2val, double *__restrict__ row2val) {
  using G = ElemGeo<KIND>;
  constexpr int NN = G::N * G::N, NNODE = G::NNODE, NDIM = G::NDIM;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_slots; s += (uint64_t)gridDim.x * blockDim.x) {
    double acc[NN];
#pragma unroll
    for (int q = 0; q < NN; ++q) acc[q] = 0.0;
    const uint32_t cb = nz2ptr[s], ce = nz2ptr[s + 1];
    for (uint32_t c = cb; c < ce; ++c) {
      const uint32_t code = contrib[c];
      const uint32_t e = code / (NNODE * NNODE);
      const int ij = (int)(code - e * (NNODE * NNODE));
      const int i = ij / NNODE, j = ij - i * NNODE;
      const double *p[NNODE];
      double coords[NNODE][NDIM];
#pragma unroll
      for (int n = 0; n < NNODE; ++n) {
        const uint32_t v = e2v[(uint64_t)e * NNODE + n];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) coords[n][d] = __ldg(&xyz[(uint64_t)v * NDIM + d]);
        p[n] = coords[n];
      }
      G geo;
      geo.init(p, p0, p1);
      double blk[NN];
      geo.block(i, j, p0, p1, blk);
#pragma unroll
      for (int q = 0; q < NN; ++q) acc[q] += blk[q];
    }
    double *dst = (s < nnz) ? (idx2val + s * NN) : (row2val + (s - nnz) * NN);
#pragma unroll
    for (int q = 0; q < NN; ++q) dst[q] = acc[q];
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) k_node_recs(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz, uint64_t num_elem,
                                                   double p0, double p1, NodeRec *__restrict__ recs) {
  using G = ElemGeo<KIND>;
  constexpr int NNODE = G::NNODE, NDIM = G::NDIM;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    const double *p[NNODE];
    double coords[NNODE][NDIM];
#pragma unroll
    for (int n = 0; n < NNODE; ++n) {
      const uint32_t v = e2v[e * NNODE + n];
#pragma unroll
      for (int d = 0; d < NDIM; ++d) coords[n][d] = __ldg(&xyz[(uint64_t)v * NDIM + d]);
      p[n] = coords[n];
    }
    G g;
    g.init(p, p0, p1);
#pragma unroll
    for (int n = 0; n < NNODE; ++n) {
      const NodeRec r = NodeOps<KIND>::make(g, n);
      double4 *dst = reinterpret_cast<double4 *>(&recs[e * NNODE + n]);
      dst[0] = make_double4(r.v[0], r.v[1], r.v[2], r.v[3]);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) k_assemble_recs(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib,
                                                       uint64_t num_slots, uint64_t nnz, const NodeRec *__restrict__ recs, double p0,
                                                       double p1, double *__restrict__ idx2val, double *__restrict__ row2val) {
  using G = ElemGeo<KIND>;
  constexpr int NN = G::N * G::N, NNODE = G::NNODE;
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_slots; s += (uint64_t)gridDim.x * blockDim.x) {
    double acc[NN];
#pragma unroll
    for (int q = 0; q < NN; ++q) acc[q] = 0.0;
    const uint32_t cb = nz2ptr[s], ce = nz2ptr[s + 1];
    for (uint32_t c = cb; c < ce; ++c) {
      const uint32_t code = contrib[c];          // (e*NNODE + i)*NNODE + j
      const uint32_t ei = code / NNODE;          // e*NNODE + i
      const uint32_t j = code - ei * NNODE;
      const uint32_t ej = (ei / NNODE) * NNODE + j;
      const double2 *pa = reinterpret_cast<const double2 *>(&recs[ei]);
      const double2 *pb = reinterpret_cast<const double2 *>(&recs[ej]);
      const double2 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pb), b1 = __ldg(pb + 1);
      const NodeRec ra = {{a0.x, a0.y, a1.x, a1.y}}, rb = {{b0.x, b0.y, b1.x, b1.y}};
      double blk[NN];
      NodeOps<KIND>::block(ra, rb, p0, p1, blk);
#pragma unroll
      for (int q = 0; q < NN; ++q) acc[q] += blk[q];
    }
    // (a 4-way unrolled variant with all loads hoisted measured no faster: ncu shows the kernel bound by L1 sector throughput
    //  (86 %), not by latency: 2 record sectors + ~1 list sector per contribution and 4x-amplified 72-byte strided stores)
    double *dst = (s < nnz) ? (idx2val + s * NN) : (row2val + (s - nnz) * NN);
#pragma unroll
    for (int q = 0; q < NN; ++q) dst[q] = acc[q];
  }
}

DFB_API dfb_status dfb_assemble(dfb_plan *plan, int32_t kind, const dfb_mesh *mesh, double p0, double p1, dfb_bsr *m) {
  DFB_REQUIRE(plan && mesh && m, "dfb_assemble: NULL argument");
  const ElemInfo info = elem_info(kind);
  DFB_REQUIRE(info.n_node != 0, "dfb_assemble: unknown element kind %d", kind);
  DFB_REQUIRE(m->pat == plan->pat, "dfb_assemble: matrix and plan use different patterns");
  DFB_REQUIRE(mesh->num_elem == plan->num_elem && mesh->num_node == plan->num_node, "dfb_assemble: mesh does not match the plan");
  DFB_REQUIRE(mesh->num_node == info.n_node && mesh->ndim == info.ndim, "dfb_assemble: mesh (%d nodes, %d-D) does not fit element kind %d",
              mesh->num_node, mesh->ndim, kind);
  DFB_REQUIRE(m->blk_dim == info.N, "dfb_assemble: matrix block dim %d != element block dim %d", m->blk_dim, info.N);
  dfb_ctx *c = plan->ctx;
  if (c->assemble_variant == 0) {
    // default: fused row-tile assembly (element records in shared memory, values written straight into the SpMV's native layout)
    const dfb_status ts = dfb_assemble_tile_linear(plan, kind, mesh, p0, p1, m);
    if (ts != DFB_ERR_UNSUPPORTED) return ts;
  }
  // fallbacks / reference points, all bit-identical to the tile path: "assemble_variant" 4 (also where the tile path does not apply:
  // convention 1, flavour 1, very high valence) = node records + slot-centric gather; 1 = single kernel, geometry recomputed per
  // contribution (no scratch memory)
  DFB_TRY(dfb_bsr_canon_overwrite(m));  // every value is about to be rewritten in the canonical arrays
  const uint64_t ns = plan->num_slots;
  uint64_t g = (ns + 127) / 128;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
  if (c->assemble_variant != 1 && kind != DFB_ELEM_COT_LAPLACE_TRI3) {
    // node-record path: 32-byte record per (element, node), two aligned sectors per contribution
    uint64_t ge = (mesh->num_elem + 127) / 128;
    if (ge < 1) ge = 1;
    if (ge > (uint64_t)c->sm_count * 16) ge = (uint64_t)c->sm_count * 16;
    const uint64_t need = mesh->num_elem * info.n_node * sizeof(NodeRec) + 256;
    if (plan->geo_cap < need) {
      dfb_dev_free(plan->d_geo);
      plan->d_geo = nullptr;
      plan->geo_cap = 0;
      DFB_TRY(dfb_dev_alloc(&plan->d_geo, need));
      plan->geo_cap = need;
    }
#define REC_CASE(K)                                                                                                                   \
  case K:                                                                                                                             \
    dfb_phase_begin(c, DFB_PH_ELEM);                                                                                                  \
    DFB_LAUNCH(c, k_node_recs<K>, (unsigned)ge, 128, 0, mesh->d_elem2vtx, mesh->d_xyz, mesh->num_elem, p0, p1, (NodeRec *)plan->d_geo); \
    dfb_phase_end(c);                                                                                                                 \
    dfb_phase_begin(c, DFB_PH_GATHER);                                                                                                \
    DFB_LAUNCH(c, k_assemble_recs<K>, (unsigned)g, 128, 0, plan->d_nz2ptr, plan->d_contrib, ns, plan->pat->nnz,                       \
               (const NodeRec *)plan->d_geo, p0, p1, m->d_idx2val, m->d_row2val);                                                     \
    dfb_phase_end(c);                                                                                                                 \
    break;
    switch (kind) {
      REC_CASE(DFB_ELEM_LAPLACE_TRI2)
      REC_CASE(DFB_ELEM_SOLID_LINEAR_TRI2)
      REC_CASE(DFB_ELEM_LAPLACE_TET)
      REC_CASE(DFB_ELEM_SOLID_LINEAR_TET)
      REC_CASE(DFB_ELEM_MASS_TET)
      default:
        return dfb_fail(DFB_ERR_UNSUPPORTED, "dfb_assemble: kind %d has no fused path (use dfb_emat_batch + dfb_assemble_from_emat)", kind);
    }
#undef REC_CASE
    DFB_CHECK_LAUNCH();
    return DFB_OK;
  }
#define ASM_CASE(K)                                                                                                          \
  case K:                                                                                                                    \
    dfb_phase_begin(c, DFB_PH_GATHER);                                                                                       \
    DFB_LAUNCH(c, k_assemble_fused<K>, (unsigned)g, 128, 0, plan->d_nz2ptr, plan->d_contrib, ns, plan->pat->nnz, mesh->d_elem2vtx, \
               mesh->d_xyz, p0, p1, m->d_idx2val, m->d_row2val);                                                             \
    dfb_phase_end(c);                                                                                                        \
    break;
  switch (kind) {
    ASM_CASE(DFB_ELEM_LAPLACE_TRI2)
    ASM_CASE(DFB_ELEM_COT_LAPLACE_TRI3)
    ASM_CASE(DFB_ELEM_SOLID_LINEAR_TRI2)
    ASM_CASE(DFB_ELEM_LAPLACE_TET)
    ASM_CASE(DFB_ELEM_SOLID_LINEAR_TET)
    ASM_CASE(DFB_ELEM_MASS_TET)
    default:
      return dfb_fail(DFB_ERR_UNSUPPORTED, "dfb_assemble: kind %d has no fused path (use dfb_emat_batch + dfb_assemble_from_emat)", kind);
  }
#undef ASM_CASE
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------------ two-phase path
// batched element kernels: one thread per element writes the whole [n_node][n_node][NN] element matrix
template <int KIND>
__global__ void __launch_bounds__(128) k_emat_batch(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                    uint64_t num_elem, double p0, double p1, double *__restrict__ emat) {
  using G = ElemGeo<KIND>;
  constexpr int NN = G::N * G::N, NNODE = G::NNODE, NDIM = G::NDIM;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    const double *p[NNODE];
    double coords[NNODE][NDIM];
#pragma unroll
    for (int n = 0; n < NNODE; ++n) {
      const uint32_t v = e2v[e * NNODE + n];
#pragma unroll
      for (int d = 0; d < NDIM; ++d) coords[n][d] = __ldg(&xyz[(uint64_t)v * NDIM + d]);
      p[n] = coords[n];
    }
    G geo;
    geo.init(p, p0, p1);
    double *out = emat + e * (NNODE * NNODE * NN);
#pragma unroll
    for (int i = 0; i < NNODE; ++i)
#pragma unroll
      for (int j = 0; j < NNODE; ++j) {
        double blk[NN];
        geo.block(i, j, p0, p1, blk);
#pragma unroll
        for (int q = 0; q < NN; ++q) out[(i * NNODE + j) * NN + q] = blk[q];
      }
  }
}

// ---- DERIVED tet neo-Hookean: energy density + the reference's chain rule
// (solid_incompressive_hex.rs:86-148) with 4 nodes, dndx = g, detwei = V. Writes the Hessian already transposed to the
// column-major block convention of the BSR (a + 3b), the gradient and the energy.
__device__ __forceinline__ void mat3_det_inv(double &det, double (&inv)[3][3], const double (&c)[3][3]) {
  const double d = c[0][0] * (c[1][1] * c[2][2] - c[1][2] * c[2][1]) - c[0][1] * (c[1][0] * c[2][2] - c[1][2] * c[2][0]) +
                   c[0][2] * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
  const double t = 1.0 / d;
  inv[0][0] = t * (c[1][1] * c[2][2] - c[1][2] * c[2][1]);
  inv[0][1] = t * (c[0][2] * c[2][1] - c[0][1] * c[2][2]);
  inv[0][2] = t * (c[0][1] * c[1][2] - c[0][2] * c[1][1]);
  inv[1][0] = t * (c[1][2] * c[2][0] - c[1][0] * c[2][2]);
  inv[1][1] = t * (c[0][0] * c[2][2] - c[0][2] * c[2][0]);
  inv[1][2] = t * (c[0][2] * c[1][0] - c[0][0] * c[1][2]);
  inv[2][0] = t * (c[1][0] * c[2][1] - c[1][1] * c[2][0]);
  inv[2][1] = t * (c[0][1] * c[2][0] - c[0][0] * c[2][1]);
  inv[2][2] = t * (c[0][0] * c[1][1] - c[0][1] * c[1][0]);
  det = d;
}

__global__ void __launch_bounds__(64) k_neohookean_tet(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                       const double *__restrict__ disp, uint64_t num_elem, double myu,
                                                       double lambda, double *__restrict__ emat, double *__restrict__ evec,
                                                       double *__restrict__ w_out) {
  const int IJ[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {1, 2}, {2, 0}};  // dudx.rs:36
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    double X[4][3], U[4][3];
    for (int n = 0; n < 4; ++n) {
      const uint32_t v = e2v[e * 4 + n];
      for (int d = 0; d < 3; ++d) {
        X[n][d] = xyz[(uint64_t)v * 3 + d];
        U[n][d] = disp[(uint64_t)v * 3 + d];
      }
    }
    double g[4][3];
    const double vol = tet_vol_grad(g, X[0], X[1], X[2], X[3]);
    double dudx[3][3];  // dndx.rs:2-21
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double t = 0.0;
        for (int n = 0; n < 4; ++n) t += U[n][i] * g[n][j];
        dudx[i][j] = t;
      }
    double C[3][3], Ci[3][3], detc;  // dudx.rs:17-34
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        double t = dudx[i][j] + dudx[j][i];
        for (int k = 0; k < 3; ++k) t += dudx[k][i] * dudx[k][j];
        C[i][j] = t;
      }
      C[i][i] += 1.0;
    }
    mat3_det_inv(detc, Ci, C);
    const double lnj = 0.5 * log(detc);
    const double wr = 0.5 * myu * (C[0][0] + C[1][1] + C[2][2] - 3.0) - myu * lnj + 0.5 * lambda * lnj * lnj;
    const double sco = 0.5 * lambda * lnj - 0.5 * myu;
    const double tco = 0.5 * myu - 0.5 * lambda * lnj;
    double dwrdc[6], ddwrddc[6][6];
    for (int a = 0; a < 6; ++a) {
      const int i = IJ[a][0], j = IJ[a][1];
      dwrdc[a] = ((i == j) ? 0.5 * myu : 0.0) + sco * Ci[i][j];
      for (int b = 0; b < 6; ++b) {
        const int k = IJ[b][0], l = IJ[b][1];
        const double v1 = 0.25 * lambda * Ci[i][j] * Ci[k][l];
        const double v2 = 0.5 * tco * Ci[i][l] * Ci[j][k];
        const double v3 = 0.5 * tco * Ci[i][k] * Ci[j][l];
        ddwrddc[a][b] = v1 + v2 + v3;
      }
    }
    double z[3][3];  // def_grad_tensor dudx.rs:1-14
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) z[i][j] = dudx[i][j];
      z[i][i] += 1.0;
    }
    double dcdu0[4][3][6];
    for (int n = 0; n < 4; ++n)
      for (int d = 0; d < 3; ++d) {
        dcdu0[n][d][0] = g[n][0] * z[d][0];
        dcdu0[n][d][1] = g[n][1] * z[d][1];
        dcdu0[n][d][2] = g[n][2] * z[d][2];
        dcdu0[n][d][3] = g[n][0] * z[d][1] + g[n][1] * z[d][0];
        dcdu0[n][d][4] = g[n][1] * z[d][2] + g[n][2] * z[d][1];
        dcdu0[n][d][5] = g[n][2] * z[d][0] + g[n][0] * z[d][2];
      }
    const double detwei = vol;
    if (emat) {
      double *out = emat + e * 144;
      for (int ino = 0; ino < 4; ++ino)
        for (int jno = 0; jno < 4; ++jno) {
          double blk[9];  // reference convention a*3+b
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
              double dtmp1 = 0.0;
              for (int gg = 0; gg < 6; ++gg)
                for (int hh = 0; hh < 6; ++hh) dtmp1 += 2.0 * ddwrddc[gg][hh] * dcdu0[ino][a][gg] * dcdu0[jno][b][hh];
              blk[a * 3 + b] = 0.0 + 2.0 * detwei * dtmp1;
            }
          double dtmp2 = 0.0;
          dtmp2 += dwrdc[0] * g[ino][0] * g[jno][0];
          dtmp2 += dwrdc[1] * g[ino][1] * g[jno][1];
          dtmp2 += dwrdc[2] * g[ino][2] * g[jno][2];
          dtmp2 += dwrdc[3] * (g[ino][0] * g[jno][1] + g[ino][1] * g[jno][0]);
          dtmp2 += dwrdc[4] * (g[ino][1] * g[jno][2] + g[ino][2] * g[jno][1]);
          dtmp2 += dwrdc[5] * (g[ino][2] * g[jno][0] + g[ino][0] * g[jno][2]);
          for (int a = 0; a < 3; ++a) blk[a * 3 + a] += 2.0 * detwei * dtmp2;
          // transpose into the column-major BSR convention
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) out[(ino * 4 + jno) * 9 + a + 3 * b] = blk[a * 3 + b];
        }
    }
    if (evec) {
      for (int n = 0; n < 4; ++n)
        for (int d = 0; d < 3; ++d) {
          double dtmp1 = 0.0;
          for (int gg = 0; gg < 6; ++gg) dtmp1 += dcdu0[n][d][gg] * dwrdc[gg];
          evec[e * 12 + n * 3 + d] = 0.0 + 2.0 * detwei * dtmp1;
        }
    }
    if (w_out) w_out[e] = 0.0 + wr * detwei;
  }
}

// ---- spring3::wdwddw_squared_length_difference (del-fem-cpu/src/spring3.rs:1-28), batched over the edges of a 2-node mesh.
// Rest positions = mesh coordinates, deformed positions = rest + disp; rest length = |rest_0 - rest_1|.
__global__ void __launch_bounds__(128) k_spring3(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                 const double *__restrict__ disp, uint64_t num_elem, double stiffness,
                                                 double *__restrict__ emat, double *__restrict__ evec, double *__restrict__ w_out) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v0 = e2v[e * 2 + 0], v1 = e2v[e * 2 + 1];
    double r[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double a0 = xyz[(uint64_t)v0 * 3 + k], a1 = xyz[(uint64_t)v1 * 3 + k];
      r[k] = a0 - a1;
      v[k] = (a0 + disp[(uint64_t)v0 * 3 + k]) - (a1 + disp[(uint64_t)v1 * 3 + k]);
    }
    const double l0 = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    const double l = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const double c = l0 - l;
    if (evec) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        evec[e * 6 + k] = v[k] * (-c * stiffness / l);
        evec[e * 6 + 3 + k] = v[k] * (c * stiffness / l);
      }
    }
    if (emat) {
      const double t0 = stiffness * l0 / (l * l * l);
      const double t1 = stiffness * (l - l0) / l;
      double *out = emat + e * 36;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const double m = (v[a] * v[b]) * t0 + ((a == b) ? t1 : 0.0);
          out[0 * 9 + a + 3 * b] = m;
          out[1 * 9 + a + 3 * b] = -m;
          out[2 * 9 + a + 3 * b] = -m;
          out[3 * 9 + a + 3 * b] = m;
        }
    }
    if (w_out) w_out[e] = 0.5 * stiffness * c * c;
  }
}

__device__ __forceinline__ void cross3(double *o, const double *a, const double *b) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---- stvk_tri3::wdwddw_ (del-fem-cpu/src/stvk_tri3.rs:152-330), batched over the triangles of a 3-D surface mesh.
// St.Venant-Kirchhoff membrane: rest = mesh coordinates, deformed = rest + disp, p0 = lambda, p1 = myu.
// The 16 material terms of :282-313 are summed in the reference's order; the dN factors there are -1/0/+1, so each term is
// +-((gd[p][a] * E[x][y]) * gd[q][b]) exactly or vanishes — the signs are resolved at compile time instead of multiplied.
// Blocks are written transposed (reference index a*3+b -> BSR a+3b).
struct StvkTerm {
  int p, r, x, y, q, s;
};
__global__ void __launch_bounds__(64) k_stvk_tri3(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                  const double *__restrict__ disp, uint64_t num_elem, double lambda, double myu,
                                                  double *__restrict__ emat, double *__restrict__ evec, double *__restrict__ w_out) {
  constexpr int dN[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
  constexpr StvkTerm T[16] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 1, 1, 1}, {0, 0, 0, 2, 0, 1}, {0, 0, 0, 2, 1, 0},
                              {1, 1, 1, 0, 0, 0}, {1, 1, 1, 1, 1, 1}, {1, 1, 1, 2, 0, 1}, {1, 1, 1, 2, 1, 0},
                              {0, 1, 2, 0, 0, 0}, {0, 1, 2, 1, 1, 1}, {0, 1, 2, 2, 0, 1}, {0, 1, 2, 2, 1, 0},
                              {1, 0, 2, 0, 0, 0}, {1, 0, 2, 1, 1, 1}, {1, 0, 2, 2, 0, 1}, {1, 0, 2, 2, 1, 0}};
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    double P[3][3], Q[3][3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const uint64_t v = e2v[e * 3 + n];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        P[n][k] = xyz[v * 3 + k];
        Q[n][k] = P[n][k] + disp[v * 3 + k];
      }
    }
    double gd0[3][3], nrm[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gd0[0][k] = P[1][k] - P[0][k];
      gd0[1][k] = P[2][k] - P[0][k];
    }
    cross3(nrm, gd0[0], gd0[1]);
    const double area0 = sqrt(dot3(nrm, nrm)) * 0.5;
    {
      const double inv = 0.5 / area0;
#pragma unroll
      for (int k = 0; k < 3; ++k) gd0[2][k] = nrm[k] * inv;
    }
    double gu0[2][3];
    cross3(gu0[0], gd0[1], gd0[2]);
    {
      const double inv = 1.0 / dot3(gu0[0], gd0[0]);
#pragma unroll
      for (int k = 0; k < 3; ++k) gu0[0][k] *= inv;
    }
    cross3(gu0[1], gd0[2], gd0[0]);
    {
      const double inv = 1.0 / dot3(gu0[1], gd0[1]);
#pragma unroll
      for (int k = 0; k < 3; ++k) gu0[1][k] *= inv;
    }
    double gd[2][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gd[0][k] = Q[1][k] - Q[0][k];
      gd[1][k] = Q[2][k] - Q[0][k];
    }
    const double gl[3] = {0.5 * (dot3(gd[0], gd[0]) - dot3(gd0[0], gd0[0])), 0.5 * (dot3(gd[1], gd[1]) - dot3(gd0[1], gd0[1])),
                          1.0 * (dot3(gd[0], gd[1]) - dot3(gd0[0], gd0[1]))};
    const double g[3] = {dot3(gu0[0], gu0[0]), dot3(gu0[1], gu0[1]), dot3(gu0[1], gu0[0])};
    const double E[3][3] = {
        {lambda * g[0] * g[0] + 2.0 * myu * (g[0] * g[0]), lambda * g[0] * g[1] + 2.0 * myu * (g[2] * g[2]),
         lambda * g[0] * g[2] + 2.0 * myu * (g[0] * g[2])},
        {lambda * g[1] * g[0] + 2.0 * myu * (g[2] * g[2]), lambda * g[1] * g[1] + 2.0 * myu * (g[1] * g[1]),
         lambda * g[1] * g[2] + 2.0 * myu * (g[2] * g[1])},
        {lambda * g[2] * g[0] + 2.0 * myu * (g[0] * g[2]), lambda * g[2] * g[1] + 2.0 * myu * (g[2] * g[1]),
         lambda * g[2] * g[2] + 1.0 * myu * (g[0] * g[1] + g[2] * g[2])}};
    double S[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) S[s] = E[s][0] * gl[0] + E[s][1] * gl[1] + E[s][2] * gl[2];
    if (w_out) w_out[e] = 0.5 * area0 * (gl[0] * S[0] + gl[1] * S[1] + gl[2] * S[2]);
    if (evec) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int a = 0; a < 3; ++a)
          evec[e * 9 + i * 3 + a] = area0 * (S[0] * gd[0][a] * (double)dN[i][0] + S[2] * gd[0][a] * (double)dN[i][1] +
                                             S[2] * gd[1][a] * (double)dN[i][0] + S[1] * gd[1][a] * (double)dN[i][1]);
    }
    if (emat) {
      double *out = emat + e * 81;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          double M[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) M[k] = (gd[T[k].p][a] * E[T[k].x][T[k].y]) * gd[T[k].q][b];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              double t = 0.0;
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const int sg = dN[i][T[k].r] * dN[j][T[k].s];
                if (sg > 0) t += M[k];
                if (sg < 0) t -= M[k];
              }
              t *= area0;
              if (a == b)
                t += area0 * (S[0] * (double)dN[i][0] * (double)dN[j][0] + S[2] * (double)dN[i][0] * (double)dN[j][1] +
                              S[2] * (double)dN[i][1] * (double)dN[j][0] + S[1] * (double)dN[i][1] * (double)dN[j][1]);
              out[(i * 3 + j) * 9 + a + 3 * b] = t;
            }
        }
    }
  }
}

// ---- quadratic bending hinge: energy p0/2 |w|^2 of the hinge vector w = sum_n k_n x_n of quadratic_bending::wdw
// (del-fem-cpu/src/quadratic_bending.rs:1-43; the energy/Hessian is DERIVED: w is linear in x, so H_ij = p0 k_i k_j I_3).
// Element = 4-node hinge (0,1 = wing vertices, 2-3 = shared edge; triangles (0,2,3) and (1,3,2)).
__device__ __forceinline__ double tri3_area(const double *a, const double *b, const double *c) {
  double u[3], v[3], n[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    u[k] = b[k] - a[k];
    v[k] = c[k] - a[k];
  }
  cross3(n, u, v);
  return sqrt(dot3(n, n)) * 0.5;
}
__global__ void __launch_bounds__(128) k_bending_quad(const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                      const double *__restrict__ disp, uint64_t num_elem, double stiffness,
                                                      double *__restrict__ emat, double *__restrict__ evec, double *__restrict__ w_out) {
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < num_elem; e += (uint64_t)gridDim.x * blockDim.x) {
    double P[4][3], Q[4][3];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const uint64_t v = e2v[e * 4 + n];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        P[n][k] = xyz[v * 3 + k];
        Q[n][k] = P[n][k] + disp[v * 3 + k];
      }
    }
    const double area0 = tri3_area(P[0], P[2], P[3]);
    const double area1 = tri3_area(P[1], P[3], P[2]);
    double e23[3], e02[3], e03[3], e12[3], e13[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      e23[k] = P[3][k] - P[2][k];
      e02[k] = P[2][k] - P[0][k];
      e03[k] = P[3][k] - P[0][k];
      e12[k] = P[2][k] - P[1][k];
      e13[k] = P[3][k] - P[1][k];
    }
    const double len = sqrt(dot3(e23, e23));
    const double h0 = 2.0 * area0 / len;
    const double h1 = 2.0 * area1 / len;
    const double cot023 = -dot3(e02, e23) / h0;
    const double cot032 = dot3(e03, e23) / h0;
    const double cot123 = -dot3(e12, e23) / h1;
    const double cot132 = dot3(e13, e23) / h1;
    const double tmp0 = sqrt(3.0) / (sqrt(area0 + area1) * len);
    const double k[4] = {(-cot023 - cot032) * tmp0, (-cot123 - cot132) * tmp0, (cot032 + cot132) * tmp0, (cot023 + cot123) * tmp0};
    double c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) c[d] = k[0] * Q[0][d] + k[1] * Q[1][d] + k[2] * Q[2][d] + k[3] * Q[3][d];
    if (w_out) w_out[e] = 0.5 * stiffness * (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    if (evec) {
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int d = 0; d < 3; ++d) evec[e * 12 + n * 3 + d] = (stiffness * k[n]) * c[d];
    }
    if (emat) {
      double *out = emat + e * 144;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double v = (stiffness * k[i]) * k[j];
#pragma unroll
          for (int q = 0; q < 9; ++q) out[(i * 4 + j) * 9 + q] = (q == 0 || q == 4 || q == 8) ? v : 0.0;
        }
    }
  }
}

DFB_API dfb_status dfb_emat_batch(dfb_ctx *ctx, int32_t kind, const dfb_mesh *mesh, double p0, double p1, const dfb_vec *disp,
                                  double *emat_dev, double *evec_dev, double *w_dev) {
  DFB_REQUIRE(ctx && mesh, "dfb_emat_batch: NULL argument");
  const ElemInfo info = elem_info(kind);
  DFB_REQUIRE(info.n_node != 0, "dfb_emat_batch: unknown element kind %d", kind);
  DFB_REQUIRE(mesh->num_node == info.n_node && mesh->ndim == info.ndim, "dfb_emat_batch: mesh does not fit element kind %d", kind);
  uint64_t g = (mesh->num_elem + 127) / 128;
  if (g < 1) g = 1;
  if (g > (uint64_t)ctx->sm_count * 16) g = (uint64_t)ctx->sm_count * 16;
#define EM_CASE(K)                                                                                                        \
  case K:                                                                                                                 \
    DFB_REQUIRE(emat_dev, "dfb_emat_batch: emat_dev is NULL");                                                            \
    DFB_LAUNCH(ctx, k_emat_batch<K>, (unsigned)g, 128, 0, mesh->d_elem2vtx, mesh->d_xyz, mesh->num_elem, p0, p1, emat_dev); \
    break;
  switch (kind) {
    EM_CASE(DFB_ELEM_LAPLACE_TRI2)
    EM_CASE(DFB_ELEM_COT_LAPLACE_TRI3)
    EM_CASE(DFB_ELEM_SOLID_LINEAR_TRI2)
    EM_CASE(DFB_ELEM_LAPLACE_TET)
    EM_CASE(DFB_ELEM_SOLID_LINEAR_TET)
    EM_CASE(DFB_ELEM_MASS_TET)
    case DFB_ELEM_NEOHOOKEAN_TET: {
      DFB_REQUIRE(disp && disp->blk_dim == 3 && disp->n_blk + disp->n_ghost >= mesh->num_vtx,
                  "dfb_emat_batch: neo-Hookean needs a displacement vector covering every mesh vertex");
      uint64_t g2 = (mesh->num_elem + 63) / 64;
      if (g2 < 1) g2 = 1;
      if (g2 > (uint64_t)ctx->sm_count * 32) g2 = (uint64_t)ctx->sm_count * 32;
      DFB_LAUNCH(ctx, k_neohookean_tet, (unsigned)g2, 64, 0, mesh->d_elem2vtx, mesh->d_xyz, disp->d, mesh->num_elem, p0, p1,
                 emat_dev, evec_dev, w_dev);
      break;
    }
    case DFB_ELEM_SPRING3: {
      DFB_REQUIRE(disp && disp->blk_dim == 3 && disp->n_blk + disp->n_ghost >= mesh->num_vtx,
                  "dfb_emat_batch: spring3 needs a displacement vector covering every mesh vertex");
      DFB_LAUNCH(ctx, k_spring3, (unsigned)g, 128, 0, mesh->d_elem2vtx, mesh->d_xyz, disp->d, mesh->num_elem, p0, emat_dev, evec_dev, w_dev);
      break;
    }
    case DFB_ELEM_STVK_TRI3: {
      DFB_REQUIRE(disp && disp->blk_dim == 3 && disp->n_blk + disp->n_ghost >= mesh->num_vtx,
                  "dfb_emat_batch: stvk_tri3 needs a displacement vector covering every mesh vertex");
      uint64_t g2 = (mesh->num_elem + 63) / 64;
      if (g2 < 1) g2 = 1;
      if (g2 > (uint64_t)ctx->sm_count * 32) g2 = (uint64_t)ctx->sm_count * 32;
      DFB_LAUNCH(ctx, k_stvk_tri3, (unsigned)g2, 64, 0, mesh->d_elem2vtx, mesh->d_xyz, disp->d, mesh->num_elem, p0, p1, emat_dev, evec_dev,
                 w_dev);
      break;
    }
    case DFB_ELEM_BENDING_QUAD: {
      DFB_REQUIRE(disp && disp->blk_dim == 3 && disp->n_blk + disp->n_ghost >= mesh->num_vtx,
                  "dfb_emat_batch: bending_quad needs a displacement vector covering every mesh vertex");
      DFB_LAUNCH(ctx, k_bending_quad, (unsigned)g, 128, 0, mesh->d_elem2vtx, mesh->d_xyz, disp->d, mesh->num_elem, p0, emat_dev, evec_dev,
                 w_dev);
      break;
    }
    default:
      return dfb_fail(DFB_ERR_UNSUPPORTED, "dfb_emat_batch: kind %d not implemented", kind);
  }
#undef EM_CASE
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// gather of precomputed element matrices through the plan (any element kind / user-supplied emat)
template <int NN>
__global__ void __launch_bounds__(128) k_assemble_from_emat(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib,
                                                            uint64_t num_slots, uint64_t nnz, const double *__restrict__ emat,
                                                            double *__restrict__ idx2val, double *__restrict__ row2val) {
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < num_slots; s += (uint64_t)gridDim.x * blockDim.x) {
    double acc[NN];
#pragma unroll
    for (int q = 0; q < NN; ++q) acc[q] = 0.0;
    for (uint32_t c = nz2ptr[s]; c < nz2ptr[s + 1]; ++c) {
      const double *b = emat + (uint64_t)contrib[c] * NN;
#pragma unroll
      for (int q = 0; q < NN; ++q) acc[q] += b[q];
    }
    double *dst = (s < nnz) ? (idx2val + s * NN) : (row2val + (s - nnz) * NN);
#pragma unroll
    for (int q = 0; q < NN; ++q) dst[q] = acc[q];
  }
}

DFB_API dfb_status dfb_assemble_from_emat(dfb_plan *plan, const double *emat_dev, dfb_bsr *m) {
  DFB_REQUIRE(plan && emat_dev && m, "dfb_assemble_from_emat: NULL argument");
  DFB_REQUIRE(m->pat == plan->pat, "dfb_assemble_from_emat: matrix and plan use different patterns");
  DFB_TRY(dfb_bsr_canon_overwrite(m));  // every value is about to be rewritten in the canonical arrays
  dfb_ctx *c = plan->ctx;
  const uint64_t ns = plan->num_slots;
  uint64_t g = (ns + 127) / 128;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
#define FE_CASE(NNV)                                                                                                       \
  case NNV:                                                                                                                \
    DFB_LAUNCH(c, k_assemble_from_emat<NNV>, (unsigned)g, 128, 0, plan->d_nz2ptr, plan->d_contrib, ns, plan->pat->nnz, emat_dev, \
               m->d_idx2val, m->d_row2val);                                                                                \
    break;
  switch (m->blk_dim * m->blk_dim) {
    FE_CASE(1)
    FE_CASE(4)
    FE_CASE(9)
    FE_CASE(16)
  }
#undef FE_CASE
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// gradient gather: dw[row] = sum over the diagonal slot's contributions (e,i,i) of evec[e][i], ascending e
__global__ void k_gather_evec(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib, uint64_t nnz,
                              uint64_t num_rows, int n_node, int N, const double *__restrict__ evec, double *__restrict__ out) {
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < num_rows; r += (uint64_t)gridDim.x * blockDim.x) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint32_t c = nz2ptr[nnz + r]; c < nz2ptr[nnz + r + 1]; ++c) {
      const uint32_t code = contrib[c];
      const uint32_t ei = code / n_node;  // e*n_node + i
      for (int d = 0; d < N; ++d) acc[d] += evec[(uint64_t)ei * N + d];
    }
    for (int d = 0; d < N; ++d) out[r * N + d] = acc[d];
  }
}

__global__ void __launch_bounds__(256) k_sum(const double *__restrict__ a, uint64_t n, double *partials, unsigned int *ticket, double *out) {
  __shared__ double s_buf[32];
  double acc[1] = {0.0};
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc[0] += a[i];
  if (grid_sum_last_block<1>(acc, partials, ticket, s_buf) && threadIdx.x == 0) out[0] = acc[0];
}

// energy-model assembly (two-phase): Hessian -> m (deterministic gather), gradient -> dw, energy -> *w.
// kinds: DFB_ELEM_NEOHOOKEAN_TET (p0 = myu, p1 = lambda), DFB_ELEM_SPRING3 (p0 = stiffness), DFB_ELEM_STVK_TRI3 (p0 = lambda,
// p1 = myu), DFB_ELEM_BENDING_QUAD (p0 = stiffness)
DFB_API dfb_status dfb_assemble_energy(dfb_plan *plan, int32_t kind, const dfb_mesh *mesh, const dfb_vec *disp, double p0, double p1,
                                       dfb_bsr *m, dfb_vec *dw, double *w) {
  DFB_REQUIRE(plan && mesh && disp && m, "dfb_assemble_energy: NULL argument");
  DFB_REQUIRE(kind == DFB_ELEM_NEOHOOKEAN_TET || kind == DFB_ELEM_SPRING3 || kind == DFB_ELEM_STVK_TRI3 || kind == DFB_ELEM_BENDING_QUAD,
              "dfb_assemble_energy: kind %d is not an energy model", kind);
  const ElemInfo info = elem_info(kind);
  DFB_REQUIRE(plan->pat->convention == 0 && plan->flavour == 0, "dfb_assemble_energy: needs a convention-0 / flavour-0 plan");
  DFB_REQUIRE(m->pat == plan->pat, "dfb_assemble_energy: matrix and plan use different patterns");
  DFB_REQUIRE(m->blk_dim == 3 && mesh->num_node == info.n_node && mesh->ndim == 3, "dfb_assemble_energy: needs 3x3 blocks and a matching mesh");
  DFB_REQUIRE(mesh->num_elem == plan->num_elem && mesh->num_node == plan->num_node, "dfb_assemble_energy: mesh does not match the plan");
  dfb_ctx *c = plan->ctx;
  const uint64_t nn = (uint64_t)info.n_node;
  double *d_emat = nullptr, *d_evec = nullptr, *d_w = nullptr;
  if (c->assemble_variant == 0 && (kind == DFB_ELEM_SPRING3 || kind == DFB_ELEM_NEOHOOKEAN_TET)) {
    // fused row-tile path: Hessian, gradient and per-element energies in one launch, no element-matrix scratch
    DFB_TRY(dfb_dev_alloc_t(&d_w, mesh->num_elem));
    dfb_status ts = dfb_assemble_tile_energy(plan, kind, mesh, disp, p0, p1, m, dw ? dw->d : nullptr, w ? d_w : nullptr);
    if (ts != DFB_ERR_UNSUPPORTED) {
      if (ts == DFB_OK && w) {
        dfb_phase_begin(c, DFB_PH_GRAD);
        uint64_t g = (mesh->num_elem + 1023) / 1024;
        if (g < 1) g = 1;
        if (g > 1024) g = 1024;
        DFB_LAUNCH(c, k_sum, (unsigned)g, 256, 0, d_w, mesh->num_elem, c->d_partials, c->d_ticket + 3, c->d_scratch + 8);
        cudaMemcpyAsync(c->h_pinned, c->d_scratch + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        dfb_phase_end(c);
      }
      cudaError_t e = cudaStreamSynchronize(c->stream);
      if (ts == DFB_OK && e != cudaSuccess) ts = dfb_fail(DFB_ERR_CUDA, "dfb_assemble_energy: %s", cudaGetErrorString(e));
      if (ts == DFB_OK && w) *w = *(double *)c->h_pinned;
      dfb_dev_free(d_w);
      return ts;
    }
    dfb_dev_free(d_w);
    d_w = nullptr;
  }
  dfb_status st = dfb_dev_alloc_t(&d_emat, mesh->num_elem * nn * nn * 9);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_evec, mesh->num_elem * nn * 3);
  if (st == DFB_OK) st = dfb_dev_alloc_t(&d_w, mesh->num_elem);
  if (st == DFB_OK) {
    dfb_phase_begin(c, DFB_PH_ELEM);
    st = dfb_emat_batch(c, kind, mesh, p0, p1, disp, d_emat, d_evec, d_w);
    dfb_phase_end(c);
  }
  if (st == DFB_OK) {
    dfb_phase_begin(c, DFB_PH_GATHER);
    st = dfb_assemble_from_emat(plan, d_emat, m);
    dfb_phase_end(c);
  }
  dfb_phase_begin(c, DFB_PH_GRAD);
  if (st == DFB_OK && dw) {
    DFB_LAUNCH(c, k_gather_evec, c->sm_count * 8, 128, 0, plan->d_nz2ptr, plan->d_contrib, plan->pat->nnz, plan->pat->num_blk, info.n_node, 3,
               d_evec, dw->d);
  }
  if (st == DFB_OK && w) {
    uint64_t g = (mesh->num_elem + 1023) / 1024;
    if (g < 1) g = 1;
    if (g > 1024) g = 1024;
    DFB_LAUNCH(c, k_sum, (unsigned)g, 256, 0, d_w, mesh->num_elem, c->d_partials, c->d_ticket + 3, c->d_scratch + 8);
    cudaMemcpyAsync(c->h_pinned, c->d_scratch + 8, sizeof(double), cudaMemcpyDeviceToHost, c->stream);
  }
  dfb_phase_end(c);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (st == DFB_OK && e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_assemble_energy: %s", cudaGetErrorString(e));
  if (st == DFB_OK && w) *w = *(double *)c->h_pinned;
  dfb_dev_free(d_emat);
  dfb_dev_free(d_evec);
  dfb_dev_free(d_w);
  return st;
}

DFB_API dfb_status dfb_assemble_neohookean_tet(dfb_plan *plan, const dfb_mesh *mesh, const dfb_vec *disp, double myu,
                                               double lambda, dfb_bsr *m, dfb_vec *dw, double *w) {
  return dfb_assemble_energy(plan, DFB_ELEM_NEOHOOKEAN_TET, mesh, disp, myu, lambda, m, dw, w);
}

// ---- lumped mass (DERIVED, SURVEY A9'; loop shape of mitc_tri3::mass_lumped_plate_bending, mitc_tri3.rs:321-331): every node of an
// element receives measure / n_node * rho.  Deterministic gather: row r walks the contribution list of its DIAGONAL slot, whose
// entries (e, i, i) are in ascending element order — the same order as the reference-style serial loop, hence the same bits.
template <int NNODE, int NDIM>
__global__ void __launch_bounds__(128) k_lumped_mass(const uint32_t *__restrict__ nz2ptr, const uint32_t *__restrict__ contrib, uint64_t nnz,
                                                     uint64_t num_rows, const uint32_t *__restrict__ e2v, const double *__restrict__ xyz,
                                                     double rho, int N, double *__restrict__ out) {
  for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < num_rows; r += (uint64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (uint32_t c = nz2ptr[nnz + r]; c < nz2ptr[nnz + r + 1]; ++c) {
      const uint64_t e = contrib[c] / (NNODE * NNODE);
      double P[NNODE][NDIM];
#pragma unroll
      for (int n = 0; n < NNODE; ++n) {
        const uint64_t v = e2v[e * NNODE + n];
#pragma unroll
        for (int d = 0; d < NDIM; ++d) P[n][d] = __ldg(&xyz[v * NDIM + d]);
      }
      double meas;
      if constexpr (NNODE == 3 && NDIM == 2) {
        meas = tri2_area(P[0], P[1], P[2]);
      } else if constexpr (NNODE == 3 && NDIM == 3) {
        const double u[3] = {P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2]};
        const double v[3] = {P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2]};
        const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        meas = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * 0.5;
      } else {
        double g[4][3];
        meas = tet_vol_grad(g, P[0], P[1], P[2], P[3]);
      }
      acc += meas / (double)NNODE * rho;
    }
    for (int d = 0; d < N; ++d) out[r * N + d] = acc;
  }
}

DFB_API dfb_status dfb_lumped_mass(dfb_plan *plan, const dfb_mesh *mesh, double rho, dfb_vec *out) {
  DFB_REQUIRE(plan && mesh && out, "dfb_lumped_mass: NULL argument");
  DFB_REQUIRE(plan->pat->convention == 0 && plan->flavour == 0, "dfb_lumped_mass: needs a convention-0 / flavour-0 plan");
  DFB_REQUIRE(mesh->num_elem == plan->num_elem && mesh->num_node == plan->num_node, "dfb_lumped_mass: mesh does not match the plan");
  DFB_REQUIRE(out->n_blk == plan->pat->num_blk, "dfb_lumped_mass: vector has %llu blocks, pattern %llu", (unsigned long long)out->n_blk,
              (unsigned long long)plan->pat->num_blk);
  dfb_ctx *c = plan->ctx;
  const uint64_t nr = plan->pat->num_blk;
  uint64_t g = (nr + 127) / 128;
  if (g < 1) g = 1;
  if (g > (uint64_t)c->sm_count * 16) g = (uint64_t)c->sm_count * 16;
#define LM_CASE(NNODE, NDIM)                                                                                                          \
  DFB_LAUNCH(c, (k_lumped_mass<NNODE, NDIM>), (unsigned)g, 128, 0, plan->d_nz2ptr, plan->d_contrib, plan->pat->nnz, nr, mesh->d_elem2vtx, \
             mesh->d_xyz, rho, out->blk_dim, out->d)
  if (mesh->num_node == 3 && mesh->ndim == 2) {
    LM_CASE(3, 2);
  } else if (mesh->num_node == 3 && mesh->ndim == 3) {
    LM_CASE(3, 3);
  } else if (mesh->num_node == 4 && mesh->ndim == 3) {
    LM_CASE(4, 3);
  } else {
    return dfb_fail(DFB_ERR_UNSUPPORTED, "dfb_lumped_mass: %d-node elements in %d-D are not supported", mesh->num_node, mesh->ndim);
  }
#undef LM_CASE
  DFB_CHECK_LAUNCH();
  return DFB_OK;
}

// one element on host arrays (tests / API parity with the reference's by-value element functions)
DFB_API dfb_status dfb_emat_one(dfb_ctx *ctx, int32_t kind, const double *node2xyz, double p0, double p1, double *emat) {
  DFB_REQUIRE(ctx && node2xyz && emat, "dfb_emat_one: NULL argument");
  const ElemInfo info = elem_info(kind);
  DFB_REQUIRE(info.n_node != 0 && (kind <= DFB_ELEM_SOLID_LINEAR_TET || kind == DFB_ELEM_MASS_TET), "dfb_emat_one: unsupported kind %d", kind);
  uint32_t conn[4] = {0, 1, 2, 3};
  dfb_mesh *mesh = nullptr;
  DFB_TRY(dfb_mesh_create_u32(ctx, conn, 1, info.n_node, node2xyz, info.n_node, info.ndim, &mesh));
  const size_t n = (size_t)info.n_node * info.n_node * info.N * info.N;
  double *d_e = nullptr;
  dfb_status st = dfb_dev_alloc_t(&d_e, n);
  if (st == DFB_OK) st = dfb_emat_batch(ctx, kind, mesh, p0, p1, nullptr, d_e, nullptr, nullptr);
  if (st == DFB_OK) {
    cudaError_t e = cudaMemcpyAsync(emat, d_e, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = dfb_fail(DFB_ERR_CUDA, "dfb_emat_one: %s", cudaGetErrorString(e));
  }
  dfb_dev_free(d_e);
  dfb_mesh_destroy(mesh);
  return st;
}
