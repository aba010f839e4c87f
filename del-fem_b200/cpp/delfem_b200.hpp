// delfem_b200.hpp — header-only C++ mirror of the reference's Rust interface over the C ABI (include/delfem_b200.h).
//
// The reference is compiled code (Rust) and its toolchain is absent from the build image, so the host side above the C ABI is
// written in C++ with the reference's names, argument meaning and error behaviour:
//   del_fem_cpu::sparse_square::{Matrix, conjugate_gradient, preconditioned_conjugate_gradient, set_fix_dof_to_rhs_vector}
//       (del-fem-cpu/src/sparse_square.rs:57-347)
//   del_fem_cpu::sparse_ilu::{Preconditioner, copy_value, decompose, solve_preconditioning_vec}  (sparse_ilu.rs:3-219)
//   del_fem_ls::linearsystem::Solver                                                             (del-fem-ls/src/linearsystem.rs:8-112)
// A reference `assert!`/`panic!`/`unwrap()` becomes a thrown delfem_b200::Panic carrying the library's error string.
#pragma once
#include <array>
#include <cstdint>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/delfem_b200.h"

namespace delfem_b200 {

struct Panic : std::runtime_error {
  int status;
  Panic(int st, const std::string &msg) : std::runtime_error(msg), status(st) {}
};
inline void check(dfb_status st) {
  if (st != DFB_OK) throw Panic(st, dfb_last_error_message());
}

class Ctx {
 public:
  explicit Ctx(int device = 0) { check(dfb_ctx_create(device, &h_)); }
  ~Ctx() { dfb_ctx_destroy(h_); }
  Ctx(const Ctx &) = delete;
  Ctx &operator=(const Ctx &) = delete;
  dfb_ctx *get() const { return h_; }
  void sync() const { check(dfb_ctx_sync(h_)); }

 private:
  dfb_ctx *h_ = nullptr;
};

// `[[f64; N]]` on the device (slice_of_array.rs)
template <int N>
class DevVec {
 public:
  DevVec(const Ctx &ctx, std::size_t n_blk) : n_(n_blk) { check(dfb_vec_create(ctx.get(), n_blk, 0, N, &h_)); }
  DevVec(const Ctx &ctx, const std::vector<std::array<double, N>> &v) : DevVec(ctx, v.size()) { upload(v); }
  ~DevVec() { dfb_vec_destroy(h_); }
  DevVec(const DevVec &) = delete;
  DevVec &operator=(const DevVec &) = delete;
  void upload(const std::vector<std::array<double, N>> &v) {
    if (v.size() != n_) throw Panic(DFB_ERR_BAD_ARG, "DevVec::upload: length mismatch");
    check(dfb_vec_upload(h_, reinterpret_cast<const double *>(v.data())));
  }
  std::vector<std::array<double, N>> to_vec() const {
    std::vector<std::array<double, N>> out(n_);
    check(dfb_vec_download(h_, reinterpret_cast<double *>(out.data())));
    return out;
  }
  std::size_t len() const { return n_; }
  dfb_vec *get() const { return h_; }

 private:
  dfb_vec *h_ = nullptr;
  std::size_t n_;
};

namespace slice_of_array {  // del-fem-cpu/src/slice_of_array.rs:1-47
template <int N> void set_zero(DevVec<N> &p) { check(dfb_vec_set_zero(p.get())); }
template <int N> double dot(const DevVec<N> &a, const DevVec<N> &b) {
  double v;
  check(dfb_vec_dot(a.get(), b.get(), &v));
  return v;
}
template <int N> void copy(DevVec<N> &p, const DevVec<N> &u) { check(dfb_vec_copy(p.get(), u.get())); }
template <int N> void add_scaled_vector(DevVec<N> &u, double alpha, const DevVec<N> &p) { check(dfb_vec_add_scaled(u.get(), alpha, p.get())); }
template <int N> void scale_and_add_vec(DevVec<N> &p, double beta, const DevVec<N> &r) { check(dfb_vec_scale_and_add(p.get(), beta, r.get())); }
}  // namespace slice_of_array

// A uniform mesh on the device (flat elem2vtx with NNODE vertices per element, flat coordinates with NDIM per vertex) and the
// element -> nonzero map of (pattern, mesh): together they replace the reference's per-element loop
// `for node2vtx in elem2vtx.chunks(NNODE) { emat = ...; merge_for_array_blk(&emat, node2vtx, &mut col2idx) }` (solid_linear_tri2.rs:144-161).
class Mesh {
 public:
  Mesh(const Ctx &ctx, const std::vector<std::size_t> &elem2vtx, int num_node, const std::vector<double> &vtx2xyz, int num_dim) {
    static_assert(sizeof(std::size_t) == sizeof(uint64_t), "usize must be 64-bit");
    if (elem2vtx.size() % num_node != 0 || vtx2xyz.size() % num_dim != 0) throw Panic(DFB_ERR_BAD_ARG, "Mesh: array lengths");
    check(dfb_mesh_create(ctx.get(), reinterpret_cast<const uint64_t *>(elem2vtx.data()), elem2vtx.size() / num_node, num_node,
                          vtx2xyz.data(), vtx2xyz.size() / num_dim, num_dim, &h_));
  }
  ~Mesh() { dfb_mesh_destroy(h_); }
  Mesh(const Mesh &) = delete;
  Mesh &operator=(const Mesh &) = delete;
  dfb_mesh *get() const { return h_; }

 private:
  dfb_mesh *h_ = nullptr;
};

namespace sparse_square {

// sparse_square::Matrix<[f64; N*N]>  (sparse_square.rs:75-214)
template <int N>
class Matrix {
 public:
  static constexpr int NN = N * N;
  std::size_t num_blk = 0;

  // Matrix::from_vtx2vtx (:117)
  static Matrix from_vtx2vtx(const Ctx &ctx, const std::vector<std::size_t> &vtx2idx, const std::vector<std::size_t> &idx2vtx) {
    if (vtx2idx.empty() || vtx2idx.back() != idx2vtx.size()) throw Panic(DFB_ERR_BAD_ARG, "from_vtx2vtx: vtx2idx[num_blk] != idx2vtx.len()");
    Matrix m;
    m.num_blk = vtx2idx.size() - 1;
    static_assert(sizeof(std::size_t) == sizeof(uint64_t), "usize must be 64-bit");
    check(dfb_pattern_create(ctx.get(), reinterpret_cast<const uint64_t *>(vtx2idx.data()),
                             reinterpret_cast<const uint64_t *>(idx2vtx.data()), m.num_blk, 0, &m.pat_));
    check(dfb_bsr_create(ctx.get(), m.pat_, N, &m.h_));
    return m;
  }
  Matrix(Matrix &&o) noexcept : num_blk(o.num_blk), pat_(o.pat_), h_(o.h_), plan_(o.plan_), plan_mesh_(o.plan_mesh_) {
    o.pat_ = nullptr;
    o.h_ = nullptr;
    o.plan_ = nullptr;
  }
  Matrix(const Matrix &) = delete;
  ~Matrix() {
    if (plan_) dfb_plan_destroy(plan_);
    if (h_) dfb_bsr_destroy(h_);
    if (pat_) dfb_pattern_destroy(pat_);
  }
  void set_zero() { check(dfb_bsr_set_zero(h_)); }  // :174
  // merge_for_array_blk (:180-214); col2idx kept for signature parity (the device merge needs no scratch)
  template <int NNODE>
  void merge_for_array_blk(const std::array<std::array<std::array<double, NN>, NNODE>, NNODE> &emat,
                           const std::array<std::size_t, NNODE> &node2vtx, std::vector<std::size_t> & /*col2idx*/) {
    check(dfb_bsr_merge_one(h_, reinterpret_cast<const double *>(emat.data()), reinterpret_cast<const uint64_t *>(node2vtx.data()), NNODE, 0));
  }
  // set_fixed_dof (:129)
  void set_fixed_dof(double val_dia, const std::vector<std::array<int32_t, N>> &blk2isfix) {
    if (blk2isfix.size() != num_blk) throw Panic(DFB_ERR_BAD_ARG, "set_fixed_dof: assert_eq!(bc_flag.len(), row2val.len())");
    check(dfb_bsr_set_fixed_dof(h_, val_dia, reinterpret_cast<const int32_t *>(blk2isfix.data())));
  }
  // mult_vec (:143-171): y <- alpha*A*x + beta*y
  void mult_vec(DevVec<N> &y_vec, double beta, double alpha, DevVec<N> &x_vec) const {
    check(dfb_bsr_mult_vec(h_, y_vec.get(), beta, alpha, x_vec.get()));
  }
  // the whole element loop as one deterministic gather (kind = DFB_ELEM_*, e.g. DFB_ELEM_SOLID_LINEAR_TRI2 with p0 = lambda, p1 = myu);
  // the plan (element -> nonzero map) is built on first use for this mesh and kept
  void assemble(const Ctx &ctx, const Mesh &mesh, int kind, double p0, double p1) {
    if (!plan_ || plan_mesh_ != mesh.get()) {
      if (plan_) dfb_plan_destroy(plan_);
      plan_ = nullptr;
      check(dfb_plan_create(ctx.get(), pat_, mesh.get(), 0, &plan_));
      plan_mesh_ = mesh.get();
    }
    check(dfb_assemble(plan_, kind, mesh.get(), p0, p1, h_));
  }
  // values(self) += alpha * values(other): sums Hessians of element families assembled on the same pattern
  void add_scaled(double alpha, const Matrix &other) { check(dfb_bsr_add_scaled(h_, alpha, other.h_)); }
  // inertia / damping on the diagonal (rod3_darboux.rs:1219-1226)
  void add_to_diagonal(double scale, const DevVec<N> &v) { check(dfb_bsr_add_to_diagonal(h_, scale, v.get())); }
  void download(std::vector<std::array<double, NN>> &row2val, std::vector<std::array<double, NN>> &idx2val) const {
    uint64_t nb, nnz;
    int32_t conv;
    check(dfb_pattern_sizes(pat_, &nb, &nnz, &conv));
    row2val.resize(nb);
    idx2val.resize(nnz);
    check(dfb_bsr_download(h_, reinterpret_cast<double *>(row2val.data()), reinterpret_cast<double *>(idx2val.data())));
  }
  dfb_bsr *get() const { return h_; }
  dfb_pattern *pattern() const { return pat_; }

 private:
  Matrix() = default;
  dfb_pattern *pat_ = nullptr;
  dfb_bsr *h_ = nullptr;
  dfb_plan *plan_ = nullptr;
  const dfb_mesh *plan_mesh_ = nullptr;
};

// set_fix_dof_to_rhs_vector (:57-70)
template <int N>
void set_fix_dof_to_rhs_vector(DevVec<N> &blk2rhs, const std::vector<std::array<int32_t, N>> &blk2isfix) {
  check(dfb_vec_set_fixed_dof_zero(blk2rhs.get(), reinterpret_cast<const int32_t *>(blk2isfix.data())));
}

// conjugate_gradient (:218-261) -> convergence history (ratios, k = 0 excluded)
template <int N>
std::vector<double> conjugate_gradient(DevVec<N> &r_vec, DevVec<N> &u_vec, DevVec<N> &ap_vec, DevVec<N> &p_vec, double conv_ratio_tol,
                                       std::size_t max_iteration, const Matrix<N> &mat) {
  std::vector<double> hist(max_iteration + 2);
  uint64_t n = 0;
  check(dfb_cg(mat.get(), r_vec.get(), u_vec.get(), ap_vec.get(), p_vec.get(), conv_ratio_tol, max_iteration,
               std::numeric_limits<double>::epsilon(), 0, hist.data(), hist.size(), &n));
  hist.resize(n);
  return hist;
}

}  // namespace sparse_square

namespace sparse_ilu {
// Preconditioner on an empty off-diagonal pattern = block-Jacobi (sparse_ilu.rs:87-219), or point Jacobi
template <int N>
class Preconditioner {
 public:
  explicit Preconditioner(int kind = DFB_PRECOND_BLOCK_JACOBI) : kind_(kind) {}
  ~Preconditioner() { if (h_) dfb_precond_destroy(h_); }
  void initialize(const Ctx &ctx, const sparse_square::Matrix<N> &a) { reset(); check(dfb_precond_create(ctx.get(), kind_, a.get(), &h_)); }
  // Preconditioner::initialize_ilu0 (:28-56) / initialize_iluk (:58-64): the reference's own preconditioners, level-scheduled
  void initialize_ilu0(const Ctx &ctx, const sparse_square::Matrix<N> &a) { reset(); check(dfb_precond_create(ctx.get(), DFB_PRECOND_ILU0, a.get(), &h_)); }
  void initialize_iluk(const Ctx &ctx, const sparse_square::Matrix<N> &a, std::size_t lev_fill) {
    reset();
    check(dfb_precond_create_iluk(ctx.get(), a.get(), lev_fill, &h_));
  }
  dfb_precond *get() const { return h_; }
  const sparse_square::Matrix<N> *src = nullptr;

 private:
  void reset() {
    if (h_) dfb_precond_destroy(h_);
    h_ = nullptr;
  }
  int kind_;
  dfb_precond *h_ = nullptr;
};
template <int N> void copy_value(Preconditioner<N> &ilu, const sparse_square::Matrix<N> &a) { ilu.src = &a; }  // :87
template <int N> void decompose(Preconditioner<N> &ilu) { check(dfb_precond_update(ilu.get(), ilu.src->get())); }  // :127, unwrap() -> Panic
template <int N> void solve_preconditioning_vec(DevVec<N> &vec, const Preconditioner<N> &ilu) { check(dfb_precond_apply(ilu.get(), vec.get())); }  // :183
}  // namespace sparse_ilu

namespace sparse_square {
// preconditioned_conjugate_gradient (:264-347, generalised to N): absolute ||r_k|| incl. k = 0, trailing entry at max_nitr
template <int N>
std::vector<double> preconditioned_conjugate_gradient(DevVec<N> &r_vec, DevVec<N> &x_vec, DevVec<N> &pr_vec, DevVec<N> &p_vec,
                                                      double conv_ratio_tol, std::size_t max_nitr, const Matrix<N> &mat,
                                                      const sparse_ilu::Preconditioner<N> &ilu) {
  std::vector<double> hist(max_nitr + 3);
  uint64_t n = 0;
  check(dfb_pcg(mat.get(), ilu.get(), r_vec.get(), x_vec.get(), pr_vec.get(), p_vec.get(), conv_ratio_tol, max_nitr,
                std::numeric_limits<double>::epsilon(), hist.data(), hist.size(), &n));
  hist.resize(n);
  return hist;
}
}  // namespace sparse_square

namespace linearsystem {
// del_fem_ls::linearsystem::Solver (del-fem-ls/src/linearsystem.rs:8-112) over dfb_solver_*: same lifecycle (initialize ->
// begin_merge -> merges / BCs on sparse() and r_vec() -> end_merge -> solve_cg / solve_pcg), same defaults (1e-5f, 100, ILU(0)),
// clone() resets max_num_iteration to 100 like :109. sparse() .. p_vec() are the reference's pub fields as BORROWED C handles.
template <int N>
class Solver {
 public:
  explicit Solver(const Ctx &ctx) { check(dfb_solver_new(ctx.get(), N, &h_)); }  // Solver::new :33
  Solver(Solver &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  Solver(const Solver &) = delete;
  ~Solver() { if (h_) dfb_solver_destroy(h_); }
  // initialize(colind, rowptr) :48 — first array = row pointer, second = column indices (the reference's names are swapped)
  void initialize(const std::vector<std::size_t> &colind, const std::vector<std::size_t> &rowptr) {
    check(dfb_solver_initialize(h_, reinterpret_cast<const uint64_t *>(colind.data()), colind.size(),
                                reinterpret_cast<const uint64_t *>(rowptr.data()), rowptr.size()));
  }
  void begin_merge() { check(dfb_solver_begin_merge(h_)); }  // :60
  void end_merge() { check(dfb_solver_end_merge(h_)); }      // :67
  void solve_cg() { check(dfb_solver_solve_cg(h_)); }        // :73
  void solve_pcg() { check(dfb_solver_solve_pcg(h_)); }      // :85
  Solver clone() const {                                      // :96
    Solver t;
    check(dfb_solver_clone(h_, &t.h_));
    return t;
  }
  dfb_bsr *sparse() const { return dfb_solver_sparse(h_); }
  dfb_precond *ilu() const { return dfb_solver_ilu(h_); }
  dfb_vec *r_vec() const { return dfb_solver_r_vec(h_); }
  dfb_vec *u_vec() const { return dfb_solver_u_vec(h_); }
  std::vector<double> conv() const {
    uint64_t n = 0;
    check(dfb_solver_conv(h_, nullptr, 0, &n));
    std::vector<double> out(n);
    check(dfb_solver_conv(h_, out.data(), n, &n));
    return out;
  }
  double conv_ratio_tol() const { double t; uint64_t m; check(dfb_solver_get_params(h_, &t, &m)); return t; }
  std::size_t max_num_iteration() const { double t; uint64_t m; check(dfb_solver_get_params(h_, &t, &m)); return m; }
  void set_params(double conv_ratio_tol, std::size_t max_num_iteration) { check(dfb_solver_set_params(h_, conv_ratio_tol, max_num_iteration)); }
  dfb_solver *get() const { return h_; }

 private:
  Solver() = default;
  dfb_solver *h_ = nullptr;
};
}  // namespace linearsystem

}  // namespace delfem_b200
