// C++ twin of the reference's end-to-end test (del-fem-cpu/src/solid_linear_tri2.rs:91-326) written against the C++ mirror
// del-fem_b200/cpp/delfem_b200.hpp: per-element emat_tri2 -> Matrix::merge_for_array_blk -> energy identity (0.5 u^T A u ==
// sum of element energies) -> set_fixed_dof -> conjugate_gradient -> residual check. Runs on the GPU through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../del-fem_b200/cpp/delfem_b200.hpp"

using namespace delfem_b200;

static double splitmix(uint64_t idx) {  // same stand-in for the ChaCha stream as the oracle / bench
  uint64_t z = (idx + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

#define CHECK(cond)                                                     \
  do {                                                                  \
    if (!(cond)) {                                                      \
      std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #cond, __LINE__); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main() {
  const double lambda = 1.0, myu = 1.0;
  const int n = 14, m = n + 1;
  const std::size_t num_vtx = (std::size_t)m * m;
  std::vector<double> vtx2xy(num_vtx * 2);
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) {
      vtx2xy[(i + m * j) * 2 + 0] = (double)i / n;
      vtx2xy[(i + m * j) * 2 + 1] = (double)j / n;
    }
  std::vector<std::array<std::size_t, 3>> tri2vtx;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      const std::size_t v00 = i + m * j, v10 = v00 + 1, v01 = v00 + m, v11 = v01 + 1;
      tri2vtx.push_back({v00, v10, v11});
      tri2vtx.push_back({v00, v11, v01});
    }
  // vtx2vtx (first-seen order, is_self = false): del_msh_cpu::vtx2vtx::from_uniform_mesh restated
  std::vector<std::vector<std::size_t>> v2e(num_vtx);
  for (std::size_t e = 0; e < tri2vtx.size(); ++e)
    for (auto v : tri2vtx[e]) v2e[v].push_back(e);
  std::vector<std::size_t> row2idx(1, 0), idx2col;
  for (std::size_t v = 0; v < num_vtx; ++v) {
    std::vector<std::size_t> seen;
    for (auto e : v2e[v])
      for (auto w : tri2vtx[e]) {
        if (w == v) continue;
        bool dup = false;
        for (auto s : seen) dup |= (s == w);
        if (!dup) seen.push_back(w);
      }
    idx2col.insert(idx2col.end(), seen.begin(), seen.end());
    row2idx.push_back(idx2col.size());
  }

  Ctx ctx(0);
  auto bsm = sparse_square::Matrix<2>::from_vtx2vtx(ctx, row2idx, idx2col);
  bsm.set_zero();
  std::vector<std::array<double, 2>> u(num_vtx), r_vec(num_vtx, {0.0, 0.0});
  std::vector<std::array<int32_t, 2>> flg(num_vtx, {0, 0});
  for (std::size_t v = 0; v < num_vtx; ++v) {
    if (vtx2xy[v * 2 + 1] < 0.001) flg[v] = {1, 1};
    for (int d = 0; d < 2; ++d) u[v][d] = flg[v][d] ? 0.0 : 0.1 * (2.0 * splitmix(v * 2 + d) - 1.0);
  }
  double eng = 0.0;
  std::vector<std::size_t> col2idx;
  for (auto &t : tri2vtx) {
    double xy[6];
    for (int k = 0; k < 3; ++k) {
      xy[k * 2] = vtx2xy[t[k] * 2];
      xy[k * 2 + 1] = vtx2xy[t[k] * 2 + 1];
    }
    std::array<std::array<std::array<double, 4>, 3>, 3> emat;
    check(dfb_emat_one(ctx.get(), DFB_ELEM_SOLID_LINEAR_TRI2, xy, lambda, myu, &emat[0][0][0]));  // emat_tri2 (:23)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {  // er = emat * u (column-major 2x2), r -= er, energy
        const auto &b = emat[i][j];
        const double e0 = b[0] * u[t[j]][0] + b[2] * u[t[j]][1], e1 = b[1] * u[t[j]][0] + b[3] * u[t[j]][1];
        r_vec[t[i]][0] -= e0;
        r_vec[t[i]][1] -= e1;
        eng += 0.5 * (u[t[i]][0] * e0 + u[t[i]][1] * e1);
      }
    bsm.merge_for_array_blk<3>(emat, t, col2idx);  // (:161)
  }
  {  // 0.5 u^T A u == sum of element energies (:164-176)
    DevVec<2> ud(ctx, u), au(ctx, num_vtx);
    bsm.mult_vec(au, 0.0, 1.0, ud);
    const double pap = 0.5 * slice_of_array::dot(ud, au);
    std::printf("energy identity: %.15e vs %.15e\n", pap, eng);
    CHECK(std::fabs(pap - eng) < 1.0e-10);
  }
  for (std::size_t v = 0; v < num_vtx; ++v)
    for (int d = 0; d < 2; ++d)
      if (flg[v][d]) r_vec[v][d] = 0.0;
  bsm.set_fixed_dof(1.0, flg);  // (:188)
  DevVec<2> r(ctx, r_vec), uu(ctx, num_vtx), ap(ctx, num_vtx), p(ctx, num_vtx);
  auto hist = sparse_square::conjugate_gradient(r, uu, ap, p, 1.0e-5, 300, bsm);  // (:217)
  std::printf("CG iterations: %zu, last ratio %.3e\n", hist.size(), hist.back());
  CHECK(hist.size() > 10 && hist.size() < 300 && hist.back() < 1.0e-5);
  {  // residual r - A u (:279-289)
    DevVec<2> r1(ctx, num_vtx);
    bsm.mult_vec(r1, 0.0, 1.0, uu);
    auto a = r1.to_vec();
    double worst = 0.0;
    for (std::size_t v = 0; v < num_vtx; ++v) worst = std::fmax(worst, std::hypot(r_vec[v][0] - a[v][0], r_vec[v][1] - a[v][1]));
    std::printf("max |r - A u| = %.3e\n", worst);
    CHECK(worst < 1.0e-4);
    auto sol = uu.to_vec();  // K u_sol = -K u  =>  u_sol = -u on the free dofs
    double err = 0.0;
    for (std::size_t v = 0; v < num_vtx; ++v) err = std::fmax(err, std::hypot(sol[v][0] + u[v][0], sol[v][1] + u[v][1]));
    CHECK(err < 1.0e-3);
  }
  // PCG with the reference's ILU code path on an empty pattern (block-Jacobi)
  sparse_ilu::Preconditioner<2> ilu;
  ilu.initialize(ctx, bsm);
  sparse_ilu::copy_value(ilu, bsm);
  sparse_ilu::decompose(ilu);
  DevVec<2> r2(ctx, r_vec), x(ctx, num_vtx), pr(ctx, num_vtx), p2(ctx, num_vtx);
  auto hist2 = sparse_square::preconditioned_conjugate_gradient(r2, x, pr, p2, 1.0e-5, 300, bsm, ilu);
  std::printf("block-Jacobi PCG history length: %zu\n", hist2.size());
  CHECK(hist2.size() > 2 && hist2.back() / hist2.front() < 1.0e-5);
  // a merge outside the pattern must fail like the reference's assert_ne!(idx0, usize::MAX)
  bool threw = false;
  try {
    std::array<std::array<std::array<double, 4>, 2>, 2> bad{};
    std::array<std::size_t, 2> far{0, num_vtx - 1};
    bsm.merge_for_array_blk<2>(bad, far, col2idx);
  } catch (const Panic &e) {
    threw = e.status == DFB_ERR_PATTERN_MISS;
  }
  CHECK(threw);
  std::printf("OK\n");
  return 0;
}
