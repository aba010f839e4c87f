/* delfem_b200.h — C ABI of libdelfem_b200.so: the B200-native (sm_100a CUDA, fp64) implementation of del-fem's
 * data-parallel hot path  element kernels -> merge into "CSR + separate diagonal"/BSR -> BC -> (P)CG.
 *
 * This is the drop-in boundary a Rust `-sys` crate binds.  Plain pointers and sizes only; every
 * function returns a dfb_status (0 = OK) and never unwinds; the Rust wrapper maps non-zero to the reference's
 * `panic!`/`Result`.  Each entry point cites the reference interface it replaces (paths relative to the reference
 * root, del-fem workspace 0.1.9).
 *
 * Conventions shared with the reference
 *   - indices on the host side are `usize` (uint64_t); the device narrows to u32 (DFB_ERR_INDEX_OVERFLOW if it cannot)
 *   - an NxN block is column-major: entry (a,b) at a + N*b   (del_geo_core::matn_col_major)
 *   - nested Rust arrays `[[[T;B];N];N]` are flat row-major `(i*N + j)*B + k`
 *   - convention 0 ("csrdia"): off-diagonal pattern + separate diagonal `row2val`   (del-fem-cpu/src/sparse_square.rs:75-81)
 *     convention 1 ("blkcsr"): diagonal inside the pattern, no `row2val`             (del-fem-cpu/src/merge.rs:51-88)
 * Threading: a dfb_ctx binds one device + one stream; calls on a ctx are stream-ordered and asynchronous until a
 * download / sync / history read.  A ctx is not thread-safe; distinct ctxs are independent.
 * Ownership: handles own their device buffers; host pointers are borrowed for the duration of the call only.
 */
#ifndef DELFEM_B200_H
#define DELFEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t dfb_status;
enum {
  DFB_OK = 0,
  DFB_ERR_BAD_ARG = 1,        /* reference: assert!/assert_eq! on sizes                           */
  DFB_ERR_PATTERN_MISS = 2,   /* reference: assert_ne!(idx0, usize::MAX)  sparse_square.rs:201     */
  DFB_ERR_SINGULAR_DIAG = 3,  /* reference: try_inverse().unwrap()        sparse_ilu.rs:164        */
  DFB_ERR_CUDA = 4,           /* reference: DriverError                   del-fem-cudarc/src/rod2.rs:14 */
  DFB_ERR_NCCL = 5,
  DFB_ERR_INDEX_OVERFLOW = 6, /* usize index does not fit the device's u32                          */
  DFB_ERR_NOT_POSITIVE = 7,   /* reference: assert!(pap >= 0)             del-fem-ls/src/solver_sparse.rs:53 */
  DFB_ERR_UNSUPPORTED = 8
};

typedef struct dfb_ctx dfb_ctx;
typedef struct dfb_vec dfb_vec;
typedef struct dfb_mesh dfb_mesh;
typedef struct dfb_pattern dfb_pattern;
typedef struct dfb_bsr dfb_bsr;
typedef struct dfb_plan dfb_plan;
typedef struct dfb_precond dfb_precond;
typedef struct dfb_halo dfb_halo;
typedef struct dfb_bcflag dfb_bcflag;
typedef struct dfb_solver dfb_solver;

/* element kinds (kernel = reference function it batches) */
enum {
  DFB_ELEM_LAPLACE_TRI2 = 0,      /* laplace_tri2::ddw_                     del-fem-cpu/src/laplace_tri2.rs:3      N=1, 3 nodes, xy  */
  DFB_ELEM_COT_LAPLACE_TRI3 = 1,  /* tri3::emat_cotangent_laplacian         call site laplace_tri3.rs:18           N=1, 3 nodes, xyz */
  DFB_ELEM_SOLID_LINEAR_TRI2 = 2, /* solid_linear_tri2::emat_tri2           del-fem-cpu/src/solid_linear_tri2.rs:23 N=2, 3 nodes, xy  */
  DFB_ELEM_LAPLACE_TET = 3,       /* DERIVED from laplace_tri2.rs:12-15                                            N=1, 4 nodes, xyz */
  DFB_ELEM_SOLID_LINEAR_TET = 4,  /* DERIVED from solid_linear_tri2.rs:11-19 + solid_linear_hex.rs:18-25           N=3, 4 nodes, xyz */
  DFB_ELEM_NEOHOOKEAN_TET = 5,    /* DERIVED from solid_incompressive_hex.rs:86-148 (chain rule), 4 nodes          N=3, 4 nodes, xyz */
  DFB_ELEM_SPRING3 = 6,           /* spring3::wdwddw_squared_length_difference  del-fem-cpu/src/spring3.rs:1       N=3, 2 nodes, xyz */
  DFB_ELEM_STVK_TRI3 = 7,         /* stvk_tri3::wdwddw_                     del-fem-cpu/src/stvk_tri3.rs:152       N=3, 3 nodes, xyz */
  DFB_ELEM_BENDING_QUAD = 8,      /* energy k/2 |w|^2 DERIVED from quadratic_bending::wdw  quadratic_bending.rs:1  N=3, 4 nodes, xyz */
  DFB_ELEM_MASS_TET = 9           /* DERIVED consistent mass rho V/20 (1+delta_ij) I3 (lumped precedent mitc_tri3.rs:304) N=3, 4 nodes */
};

/* preconditioner kinds */
enum {
  DFB_PRECOND_NONE = 0,
  DFB_PRECOND_JACOBI = 1,       /* DERIVED: z[i][d] = r[i][d] / D[i][d + N d]                                             */
  DFB_PRECOND_BLOCK_JACOBI = 2, /* sparse_ilu::{copy_value,decompose,solve_preconditioning_vec} with an empty off-diagonal
                                   pattern  del-fem-cpu/src/sparse_ilu.rs:87-219                                           */
  DFB_PRECOND_ILU0 = 3          /* the reference's own preconditioner: initialize_ilu0 :28, copy_value :87, decompose :127,
                                   solve_preconditioning_vec :183 — level-scheduled on the device, one GPU only             */
};

/* ------------------------------------------------------------------------------------------------ context */
const char *dfb_last_error_message(void); /* per-thread string describing the last non-zero status */
const char *dfb_version(void);
dfb_status dfb_ctx_create(int32_t device, dfb_ctx **out);
dfb_status dfb_ctx_destroy(dfb_ctx *ctx);
dfb_status dfb_ctx_sync(dfb_ctx *ctx);
/* number of CUDA kernels this ctx has launched so far (graph replays count their kernel nodes) */
dfb_status dfb_ctx_launch_count(dfb_ctx *ctx, uint64_t *out);
/* tuning knobs: "spmv_variant" (20 = tile-packed mirror, one TMA bulk copy per 8-row tile [default]; 4 = TMA ring over the
   canonical arrays, no extra memory; 0 = row per thread), "iters_per_sync" (PCG iterations per chunk = per CUDA-graph replay and
   per convergence poll), "use_graph" (1 [default]: replay chunks as a CUDA graph), "profile_spmv", "assemble_variant",
   "use_peer_allreduce", "use_peer_halo", "peer_halo_bytes" */
dfb_status dfb_ctx_set_option(dfb_ctx *ctx, const char *key, int64_t value);
dfb_status dfb_ctx_get_option(dfb_ctx *ctx, const char *key, int64_t *value);
/* CUDA-event timing on the ctx stream (events are recorded on the stream the kernels are launched on) */
dfb_status dfb_timer_start(dfb_ctx *ctx);
dfb_status dfb_timer_stop(dfb_ctx *ctx, float *elapsed_ms); /* records, synchronises, returns elapsed */
dfb_status dfb_flush_l2(dfb_ctx *ctx);                      /* overwrites a buffer larger than L2 */
dfb_status dfb_device_info(dfb_ctx *ctx, int32_t *sm_count, int64_t *l2_bytes, int64_t *total_mem_bytes);
/* page-lock a caller-owned host buffer so that the uploads/downloads that borrow it are true asynchronous DMA */
dfb_status dfb_host_register(dfb_ctx *ctx, void *host, uint64_t bytes);
dfb_status dfb_host_unregister(dfb_ctx *ctx, void *host);
/* with option "profile_spmv" = 1 the solvers bracket every SpMV launch (first 16384 per solve) with CUDA events on the
   launching stream; this returns the accumulated device time and launch count since the last reset */
dfb_status dfb_profile_spmv(dfb_ctx *ctx, double *total_ms, uint64_t *count, int32_t reset);
/* with option "profile_phases" = 1 the library brackets its own assembly / setup launches with CUDA events per phase:
   [0] element kernels (two-phase path)  [1] merge gather  [2] gradient gather + energy sum  [3] SpMV-mirror pack
   [4] fused row-tile assembly  [5] Dirichlet BC  [6] preconditioner update  [7] unused.  ms_out / count_out: 8 entries each. */
dfb_status dfb_ctx_phase_times(dfb_ctx *ctx, double *ms_out, uint64_t *count_out, int32_t reset);
/* build the SpMV mirror of the current values now (otherwise the first SpMV after a value change does it) */
dfb_status dfb_bsr_pack(dfb_bsr *m);

/* ------------------------------------------------------------------------------------------------ vectors
 * `[[T;N]]` slices of the reference (del-fem-cpu/src/slice_of_array.rs:1-47). n_blk rows are "owned"; n_ghost extra
 * blocks at the tail hold halo copies on multi-GPU runs (0 otherwise). */
dfb_status dfb_vec_create(dfb_ctx *ctx, uint64_t n_blk, uint64_t n_ghost, int32_t blk_dim, dfb_vec **out);
dfb_status dfb_vec_destroy(dfb_vec *v);
dfb_status dfb_vec_upload(dfb_vec *v, const double *host /* n_blk*blk_dim */);
dfb_status dfb_vec_download(dfb_vec *v, double *host /* n_blk*blk_dim */);
dfb_status dfb_vec_set_zero(dfb_vec *v);                                  /* slice_of_array::set_zero  :1  */
dfb_status dfb_vec_copy(dfb_vec *dst, const dfb_vec *src);                /* slice_of_array::copy      :18 */
dfb_status dfb_vec_dot(const dfb_vec *a, const dfb_vec *b, double *out);  /* slice_of_array::dot       :8  (deterministic tree; all-reduced over ranks) */
dfb_status dfb_vec_add_scaled(dfb_vec *u, double alpha, const dfb_vec *p);/* add_scaled_vector         :26 */
dfb_status dfb_vec_scale_and_add(dfb_vec *p, double beta, const dfb_vec *r); /* scale_and_add_vec      :38 */
dfb_status dfb_vec_scale(dfb_vec *v, double s);                           /* vecn::scale_in_place per block (sparse_square.rs:156) */
/* set_fix_dof_to_rhs_vector  del-fem-cpu/src/sparse_square.rs:57-70 */
dfb_status dfb_vec_set_fixed_dof_zero(dfb_vec *rhs, const int32_t *blk2isfix_host /* n_blk*blk_dim */);
/* out[i] = lo + (hi-lo)*splitmix64(seed, first_idx + i) — synthetic inputs for benches, bit-identical to the oracle's */
dfb_status dfb_vec_fill_splitmix(dfb_vec *v, uint64_t seed, uint64_t first_idx, double lo, double hi);
double *dfb_vec_device_ptr(dfb_vec *v);
/* device-resident Dirichlet flags `&[[i32; N]]` (0 = free), so that per-step BC application needs no host traffic */
dfb_status dfb_bcflag_create(dfb_ctx *ctx, const int32_t *blk2isfix_host, uint64_t n_blk_all, int32_t blk_dim, dfb_bcflag **out);
dfb_status dfb_bcflag_destroy(dfb_bcflag *f);
dfb_status dfb_vec_set_fixed_dof_zero_dev(dfb_vec *rhs, const dfb_bcflag *f);

/* ------------------------------------------------------------------------------------------------ mesh
 * device-resident `elem2vtx` (uniform mesh, num_node per element) + `vtx2xyz` — the arguments of the reference's
 * per-mesh loops (del-fem-cpu/src/laplace_tri3.rs:1-9, del-fem-numpy/src/laplacian.rs:28-36). */
dfb_status dfb_mesh_create(dfb_ctx *ctx, const uint64_t *elem2vtx, uint64_t num_elem, int32_t num_node,
                           const double *vtx2xyz, uint64_t num_vtx, int32_t ndim, dfb_mesh **out);
dfb_status dfb_mesh_create_u32(dfb_ctx *ctx, const uint32_t *elem2vtx, uint64_t num_elem, int32_t num_node,
                               const double *vtx2xyz, uint64_t num_vtx, int32_t ndim, dfb_mesh **out);
/* synthetic Kuhn-6 tet grid generated on the device: vertices k in [k_begin, k_end) planes of an
   nx*ny*nz grid, spacing h; vertex ids are local: i + nx*(j + ny*(k-k_begin)). */
dfb_status dfb_mesh_create_kuhn(dfb_ctx *ctx, uint64_t nx, uint64_t ny, uint64_t nz, double h, uint64_t k_begin,
                                uint64_t k_end, dfb_mesh **out);
/* renumber vertices: new_id = (id - shift) mod num_vtx (elem2vtx and vtx2xyz are permuted consistently). Used by the
   slab partition to move a leading ghost plane behind the owned vertices. */
dfb_status dfb_mesh_rotate_vertex_ids(dfb_mesh *m, uint64_t shift);
dfb_status dfb_mesh_destroy(dfb_mesh *m);
dfb_status dfb_mesh_update_xyz(dfb_mesh *m, const double *vtx2xyz);
dfb_status dfb_mesh_sizes(const dfb_mesh *m, uint64_t *num_elem, int32_t *num_node, uint64_t *num_vtx, int32_t *ndim);
dfb_status dfb_mesh_download(const dfb_mesh *m, uint64_t *elem2vtx, double *vtx2xyz);

/* ------------------------------------------------------------------------------------------------ pattern
 * `row2idx`/`idx2col` of sparse_square::Matrix (del-fem-cpu/src/sparse_square.rs:75-81, from_vtx2vtx :117;
 * del-fem-ls/src/sparse_square.rs:44 symbolic_initialization). num_own = rows stored (== num_blk on one GPU). */
dfb_status dfb_pattern_create(dfb_ctx *ctx, const uint64_t *row2idx, const uint64_t *idx2col, uint64_t num_blk,
                              int32_t convention, dfb_pattern **out);
dfb_status dfb_pattern_create_u32(dfb_ctx *ctx, const uint32_t *row2idx, const uint32_t *idx2col, uint64_t num_blk,
                                  int32_t convention, dfb_pattern **out);
/* del_msh_cpu::vtx2vtx::from_uniform_mesh(elem2vtx, num_node, num_vtx, is_self) built on the device
   (call sites del-fem-cpu/src/solid_linear_tri2.rs:105, del_fem_numpy/LaplacianMatrix.py:8).
   is_self = 0 -> convention 0, is_self = 1 -> convention 1.  Rows [0, num_rows) only (num_rows <= num_vtx). */
dfb_status dfb_pattern_from_uniform_mesh(dfb_ctx *ctx, const dfb_mesh *mesh, int32_t is_self, uint64_t num_rows,
                                         dfb_pattern **out);
dfb_status dfb_pattern_destroy(dfb_pattern *p);
dfb_status dfb_pattern_sizes(const dfb_pattern *p, uint64_t *num_blk, uint64_t *nnz, int32_t *convention);
dfb_status dfb_pattern_download(const dfb_pattern *p, uint64_t *row2idx, uint64_t *idx2col);

/* ------------------------------------------------------------------------------------------------ matrix
 * sparse_square::Matrix<[T;NN]> (del-fem-cpu/src/sparse_square.rs:75-214) and del-fem-ls Matrix<MAT>. */
dfb_status dfb_bsr_create(dfb_ctx *ctx, dfb_pattern *pattern, int32_t blk_dim, dfb_bsr **out);
dfb_status dfb_bsr_destroy(dfb_bsr *m);
dfb_pattern *dfb_bsr_pattern(const dfb_bsr *m); /* borrowed: the row2idx / idx2col the matrix was created on (pub fields :76-78) */
dfb_status dfb_bsr_set_zero(dfb_bsr *m);                                       /* Matrix::set_zero :174 */
dfb_status dfb_bsr_upload(dfb_bsr *m, const double *row2val, const double *idx2val);
dfb_status dfb_bsr_download(const dfb_bsr *m, double *row2val, double *idx2val);
/* Matrix::set_fixed_dof (del-fem-cpu/src/sparse_square.rs:129) / set_fixed_bc (sparse_square.rs:3-55) */
dfb_status dfb_bsr_set_fixed_dof(dfb_bsr *m, double val_dia, const int32_t *blk2isfix_host /* (n_blk+n_ghost)*N */);
dfb_status dfb_bsr_set_fixed_dof_dev(dfb_bsr *m, double val_dia, const dfb_bcflag *f);
/* Matrix::mult_vec  y <- alpha*A*x + beta*y   :143-171 / MatrixRef::mult_vec :419-447 */
dfb_status dfb_bsr_mult_vec(const dfb_bsr *m, dfb_vec *y, double beta, double alpha, dfb_vec *x);
/* mult_square_matrices  del-fem-ls/src/sparse_matrix_multiplication.rs:109-188 (pattern: symbolic_multiplication :7-107, discovery order,
   diagonal kept apart): C = M0 * M1 for scalar (blk_dim 1) matrices of the separate-diagonal convention. Creates the pattern and the
   matrix of C; the caller destroys both (matrix first). */
dfb_status dfb_bsr_mult_square_matrices(const dfb_bsr *m0, const dfb_bsr *m1, dfb_pattern **out_pattern, dfb_bsr **out);
/* del-fem-ls mult_mat  del-fem-ls/src/sparse_square.rs:134-164: scalar matrix (blk_dim 1) applied to the num_dim = x->blk_dim interleaved
   components of x: y <- alpha*A*x + beta*y per component */
dfb_status dfb_bsr_mult_mat(const dfb_bsr *m, dfb_vec *y, double beta, double alpha, const dfb_vec *x);
/* row2val[i][d + N d] += scale * v[i][d]  — inertia / damping on the diagonal (rod3_darboux.rs:1219-1226) */
dfb_status dfb_bsr_add_to_diagonal(dfb_bsr *m, double scale, const dfb_vec *v);
/* values(m) += alpha * values(other); both on the SAME pattern handle. Sums Hessians of different element families (the reference
   merges them into one matrix in one loop nest, rod3_darboux.rs:760-850) */
dfb_status dfb_bsr_add_scaled(dfb_bsr *m, double alpha, const dfb_bsr *other);
/* single-element merge shims (API parity with the reference's per-element calls; one tiny launch each):
   merge_for_array_blk sparse_square.rs:180-214 (flavour 0) / ls merge :66-104 (flavour 1) / csrdia merge.rs:3 /
   blkcsr merge.rs:51 (chosen by the pattern's convention). emat is [n_node][n_node][N*N] on the host. */
dfb_status dfb_bsr_merge_one(dfb_bsr *m, const double *emat, const uint64_t *node2vtx, int32_t n_node, int32_t flavour);

/* ------------------------------------------------------------------------------------------------ assembly
 * The merge as a deterministic, atomic-free gather through a precomputed element->nonzero map. */
/* flavour 0: node-index diagonal test (sparse_square.rs:194), 1: vertex-id test (merge.rs:31, ls :87) */
dfb_status dfb_plan_create(dfb_ctx *ctx, dfb_pattern *pattern, const dfb_mesh *mesh, int32_t flavour, dfb_plan **out);
dfb_status dfb_plan_destroy(dfb_plan *p);
/* elem2nz[(e*n_node + i)*n_node + j] = index into idx2val, or nnz + row for a diagonal block (convention 0) */
dfb_status dfb_plan_download_elem2nz(const dfb_plan *p, uint32_t *elem2nz);
dfb_status dfb_plan_sizes(const dfb_plan *p, uint64_t *num_slots, uint64_t *num_contrib);
/* fused element kernel + gather (no element-matrix scratch). kinds 0..4, 6. matrix values are overwritten
   (== set_zero + element loop + merge).  p0,p1: alpha | lambda,myu | stiffness. */
dfb_status dfb_assemble(dfb_plan *plan, int32_t kind, const dfb_mesh *mesh, double p0, double p1, dfb_bsr *m);
/* batched element kernels (two-phase path): emat_dev is [num_elem][n_node][n_node][N*N] device memory */
dfb_status dfb_emat_batch(dfb_ctx *ctx, int32_t kind, const dfb_mesh *mesh, double p0, double p1, const dfb_vec *disp,
                          double *emat_dev, double *evec_dev /* [num_elem][n_node][N] or NULL */, double *w_dev /* [num_elem] or NULL */);
dfb_status dfb_assemble_from_emat(dfb_plan *plan, const double *emat_dev, dfb_bsr *m);
/* neo-Hookean Newton step input: Hessian -> m, gradient -> dw (deterministic gather), energy -> *w */
dfb_status dfb_assemble_neohookean_tet(dfb_plan *plan, const dfb_mesh *mesh, const dfb_vec *disp, double myu,
                                       double lambda, dfb_bsr *m, dfb_vec *dw, double *w);
/* generic energy-model assembly (two-phase): kinds DFB_ELEM_NEOHOOKEAN_TET (p0 = myu, p1 = lambda) and DFB_ELEM_SPRING3 (p0 = stiffness,
   mesh = edges, rest positions = mesh coordinates; del-fem-cpu/src/spring3.rs:1 merged as in rod3_darboux.rs:775-792),
   DFB_ELEM_STVK_TRI3 (p0 = lambda, p1 = myu, mesh = surface triangles; stvk_tri3.rs:152-330) and DFB_ELEM_BENDING_QUAD
   (p0 = stiffness, mesh = 4-node hinges (wing, wing, edge, edge); quadratic_bending.rs:1-43) */
dfb_status dfb_assemble_energy(dfb_plan *plan, int32_t kind, const dfb_mesh *mesh, const dfb_vec *disp, double p0, double p1,
                               dfb_bsr *m, dfb_vec *dw, double *w);
/* lumped mass: out[v][0..N) = sum over elements at v of measure / n_node * rho (tri2 area, tri3 area, tet volume); deterministic
   gather in ascending element order. DERIVED; loop shape of mitc_tri3::mass_lumped_plate_bending, del-fem-cpu/src/mitc_tri3.rs:304-332 */
dfb_status dfb_lumped_mass(dfb_plan *plan, const dfb_mesh *mesh, double rho, dfb_vec *out);
/* one element on host arrays (tests): node2xyz [n_node][ndim], out emat [n_node][n_node][N*N] */
dfb_status dfb_emat_one(dfb_ctx *ctx, int32_t kind, const double *node2xyz, double p0, double p1, double *emat);
/* device scratch helpers for the two-phase path */
dfb_status dfb_malloc(dfb_ctx *ctx, uint64_t bytes, void **dev_ptr);
dfb_status dfb_free(dfb_ctx *ctx, void *dev_ptr);
dfb_status dfb_memcpy_d2h(dfb_ctx *ctx, void *host, const void *dev, uint64_t bytes);
dfb_status dfb_memcpy_h2d(dfb_ctx *ctx, void *dev, const void *host, uint64_t bytes);

/* ------------------------------------------------------------------------------------------------ solvers */
dfb_status dfb_precond_create(dfb_ctx *ctx, int32_t kind, const dfb_bsr *m, dfb_precond **out);
/* Preconditioner::initialize_iluk  del-fem-cpu/src/sparse_ilu.rs:58-64 (symbolic_iluk :221-312): ILU on the level-of-fill pattern;
   lev_fill = 0 is DFB_PRECOND_ILU0, a very large lev_fill the complete LU (linearsystem.rs:55). Single GPU, level-scheduled. */
dfb_status dfb_precond_create_iluk(dfb_ctx *ctx, const dfb_bsr *m, uint64_t lev_fill, dfb_precond **out);
/* copy_value + decompose  (del-fem-cpu/src/sparse_ilu.rs:87-181; Solver::end_merge del-fem-ls/src/linearsystem.rs:67) */
dfb_status dfb_precond_update(dfb_precond *pc, const dfb_bsr *m);
/* solve_preconditioning_vec  sparse_ilu.rs:183-219 (in place) */
dfb_status dfb_precond_apply(const dfb_precond *pc, dfb_vec *v);
dfb_status dfb_precond_destroy(dfb_precond *pc);
/* conjugate_gradient  del-fem-cpu/src/sparse_square.rs:218-261 (eps_sq = DBL_EPSILON) /
   del-fem-ls/src/solver_sparse.rs:5-70 (eps_sq = 1e-20f, check_pap = 1).
   History = ||r_k||/||r_0|| for k>=1; *n_hist is the reference's hist.len() even if it exceeds hist_cap. */
dfb_status dfb_cg(const dfb_bsr *m, dfb_vec *r, dfb_vec *u, dfb_vec *ap, dfb_vec *p, double tol, uint64_t max_it,
                  double eps_sq, int32_t check_pap, double *hist_out, uint64_t hist_cap, uint64_t *n_hist);
/* preconditioned_conjugate_gradient  del-fem-cpu/src/sparse_square.rs:264-347 generalised to N /
   del-fem-ls/src/solver_sparse.rs:73-175.  History = absolute ||r_k|| incl. k=0, +1 trailing entry if max_it exhausted. */
dfb_status dfb_pcg(const dfb_bsr *m, const dfb_precond *pc, dfb_vec *r, dfb_vec *x, dfb_vec *pr, dfb_vec *p, double tol,
                   uint64_t max_it, double eps_sq, double *hist_out, uint64_t hist_cap, uint64_t *n_hist);

/* ------------------------------------------------------------------------------------------------ linearsystem::Solver
 * del-fem-ls/src/linearsystem.rs:8-112, one entry point per method. The solver owns sparse / ilu / r_vec / u_vec / ap_vec / p_vec /
 * conv / conv_ratio_tol (1e-5f) / max_num_iteration (100); the accessors return BORROWED handles (valid until the next
 * dfb_solver_initialize / dfb_solver_destroy) on which the caller merges and applies BCs between begin_merge and end_merge, exactly
 * as reference code works on `solver.sparse` and `solver.r_vec`. Generalised from scalar T to blk_dim 1..4. */
dfb_status dfb_solver_new(dfb_ctx *ctx, int32_t blk_dim, dfb_solver **out);               /* Solver::new         :33 */
/* Solver::initialize(colind, rowptr) :48 — the reference's parameter NAMES are swapped relative to their meaning: `colind` is the
   row-pointer array (colind_len = num_blk + 1), `rowptr` the column indices. Sets up ILU(0) like :53 (initialize_ilu0). */
dfb_status dfb_solver_initialize(dfb_solver *s, const uint64_t *colind, uint64_t colind_len, const uint64_t *rowptr, uint64_t rowptr_len);
/* the alternatives the reference keeps commented out at :50,:54-56 (ILU(k): kind = DFB_PRECOND_ILU0 with lev_fill > 0) and the
   DERIVED Jacobi kinds; takes effect immediately if the solver is initialised, else at dfb_solver_initialize */
dfb_status dfb_solver_set_preconditioner(dfb_solver *s, int32_t kind, uint64_t lev_fill);
dfb_status dfb_solver_begin_merge(dfb_solver *s);                                          /* Solver::begin_merge :60 */
dfb_status dfb_solver_end_merge(dfb_solver *s);                                            /* Solver::end_merge   :67 */
dfb_status dfb_solver_solve_cg(dfb_solver *s);   /* Solver::solve_cg :73  (del-fem-ls/src/solver_sparse.rs:5-70; DFB_ERR_NOT_POSITIVE = assert!(pap >= 0)) */
dfb_status dfb_solver_solve_pcg(dfb_solver *s);  /* Solver::solve_pcg :85 (solver_sparse.rs:73-175 with self.ilu) */
/* Solver::clone :96-111 — deep copy of every field (decomposed ILU values included), max_num_iteration RESET to 100 like :109 */
dfb_status dfb_solver_clone(const dfb_solver *s, dfb_solver **out);
dfb_status dfb_solver_destroy(dfb_solver *s);
dfb_bsr *dfb_solver_sparse(dfb_solver *s);       /* pub sparse :9  */
dfb_precond *dfb_solver_ilu(dfb_solver *s);      /* pub ilu    :10 */
dfb_vec *dfb_solver_r_vec(dfb_solver *s);        /* pub r_vec  :12 */
dfb_vec *dfb_solver_u_vec(dfb_solver *s);        /* pub u_vec  :13 */
dfb_vec *dfb_solver_ap_vec(dfb_solver *s);       /* pub ap_vec :17 */
dfb_vec *dfb_solver_p_vec(dfb_solver *s);        /* pub p_vec  :18 */
dfb_status dfb_solver_set_params(dfb_solver *s, double conv_ratio_tol, uint64_t max_num_iteration); /* pub conv_ratio_tol :15, max_num_iteration :16 */
dfb_status dfb_solver_get_params(const dfb_solver *s, double *conv_ratio_tol, uint64_t *max_num_iteration);
dfb_status dfb_solver_conv(const dfb_solver *s, double *out, uint64_t cap, uint64_t *n);   /* pub conv :14 (*n = conv.len()) */

/* ||r_k||^2 for k = 0..iterations of the LAST dfb_cg / dfb_pcg on this ctx (host copy kept by the library; the device only holds a
   ring of the last two chunks). *n = iterations + 1 even if it exceeds cap. Lets a caller size its history buffer from the actual
   iteration count instead of from max_it (the reference's Vec grows per iteration, sparse_square.rs:254). */
dfb_status dfb_ctx_last_history_sq(dfb_ctx *ctx, double *out, uint64_t cap, uint64_t *n);

/* ------------------------------------------------------------------------------------------------ multi-GPU
 * Vertex-partitioned subdomains, one rank (process) per GPU.  No reference precedent: the contract is
 * that the distributed assembly + PCG equals the single-domain result. */
dfb_status dfb_comm_unique_id(uint8_t id_out[128]);
dfb_status dfb_ctx_comm_init(dfb_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t nranks);
/* Optional: NVLink peer "mailboxes" (one process per GPU on one node). Each rank exports a 64-byte CUDA IPC handle, all ranks
   exchange them (any transport) and import the full table; (P)CG then all-reduces its scalars inside its own kernels through peer
   memory instead of calling ncclAllReduce twice per iteration. Without this call the NCCL path is used. */
dfb_status dfb_ctx_peer_export(dfb_ctx *ctx, uint8_t handle_out[64]);
dfb_status dfb_ctx_peer_import(dfb_ctx *ctx, const uint8_t *handles /* nranks x 64 */, int32_t nranks);
/* halo of a vector: for each peer p, owned blocks send_idx[send_ptr[p]..send_ptr[p+1]) are sent to rank peers[p],
   and ghost blocks [recv_ptr[p], recv_ptr[p+1]) (offsets into the ghost tail) are received from it. */
dfb_status dfb_halo_create(dfb_ctx *ctx, int32_t n_peers, const int32_t *peers, const uint64_t *send_ptr,
                           const uint32_t *send_idx, const uint64_t *recv_ptr, dfb_halo **out);
dfb_status dfb_halo_destroy(dfb_halo *h);
dfb_status dfb_halo_exchange(dfb_halo *h, dfb_vec *v);
/* Optional (needs dfb_ctx_peer_import): remote_off[p] = recv_ptr entry, on rank peers[p], of the range that receives THIS rank's
   packet (in blocks). With it (P)CG pushes the boundary entries of its direction vector straight into the neighbours' landing zones
   over NVLink right after the direction kernel, and the SpMV becomes ONE launch whose boundary tiles wait for the neighbours' flags —
   instead of pack kernel + ncclSend/ncclRecv on a second stream and three SpMV launches. The landing zone is sized by the ctx option
   "peer_halo_bytes" (default 8 MiB per buffer, set before dfb_ctx_peer_export). Default on (ctx option "use_peer_halo" = 0 selects the
   NCCL halo). Standalone dfb_bsr_mult_vec / dfb_halo_exchange always use NCCL. */
dfb_status dfb_halo_set_remote(dfb_halo *h, const uint64_t *remote_off);
/* attach a halo to a matrix: mult_vec / cg / pcg then exchange x's ghosts (overlapped with interior rows) and
   all-reduce their dot products.  interior rows = rows [int_begin, int_end) reference no ghost column. */
dfb_status dfb_bsr_set_halo(dfb_bsr *m, dfb_halo *h, uint64_t int_begin, uint64_t int_end);

#ifdef __cplusplus
}
#endif
#endif /* DELFEM_B200_H */
