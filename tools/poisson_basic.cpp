// poisson_basic on the C ABI: afivo's main multigrid example (afivo/examples/poisson_basic.f90) written against the
// C++ mirror of its interface (include/afmg.hpp).  Two Gaussians (sigma 0.04 at 0.1^3 and 0.75^3, m_gaussians.f90)
// as manufactured solution on the domain 3 x 1 x 1, boxes of 16^3 cells, refinement where dr^2 |rhs| > 1e-3 up to
// level 7 - NDIM (:143-165), Dirichlet conditions from the analytic solution (:219-235), ten FMG cycles with the
// maximum residual and the maximum error printed after each (:104-119).  Known behaviour: the residual falls
// monotonically, the error settles at the discretisation level after two or three cycles.
//
//     ./poisson_basic_3d            Build: make -C tools
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "afmg.hpp"

namespace {

// m_gaussians.f90:54-105
struct gauss_t {
  int n_gauss = 2;
  double ampl[2] = {1.0, 1.0}, sigma[2] = {0.04, 0.04};
  double r0[2][3] = {{0.1, 0.1, 0.1}, {0.75, 0.75, 0.75}};
  double single(const double* r, int n) const {
    double s = 0;
    for (int d = 0; d < 3; ++d) {
      const double x = (r[d] - r0[n][d]) / sigma[n];
      s += x * x;
    }
    return ampl[n] * std::exp(-s);
  }
  double value(const double* r) const {
    double v = 0;
    for (int n = 0; n < n_gauss; ++n) v += single(r, n);
    return v;
  }
  double laplacian(const double* r) const {
    double v = 0;
    for (int n = 0; n < n_gauss; ++n) {
      double s = 0;
      for (int d = 0; d < 3; ++d) {
        const double x = (r[d] - r0[n][d]) / sigma[n];
        s += x * x;
      }
      v += 4 / (sigma[n] * sigma[n]) * (s - 0.5 * 3) * single(r, n);
    }
    return v;
  }
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

int main() {
  using namespace afmg;
  const int box_size = 16, n_iterations = 10, max_lvl = 7 - 3;
  const gauss_t gs;
  std::printf(" Running poisson_basic_3d\n");
  const int domain_size[3] = {3 * box_size, box_size, box_size};
  const double lo[3] = {0, 0, 0}, domain_len[3] = {3.0, 1.0, 1.0};
  try {
    double t0 = now();
    // refine_routine: any cell of the box with dr^2 |rhs| > 1e-3, below the maximum level
    af_t tree = af_build_tree(box_size, domain_size, max_lvl, [&](int lvl, const int*, const double* centre) {
      const double dr = domain_len[0] / domain_size[0] * std::pow(0.5, lvl - 1);
      for (int k = 0; k < box_size; ++k)
        for (int j = 0; j < box_size; ++j)
          for (int i = 0; i < box_size; ++i) {
            const double r[3] = {centre[0] + (i + 0.5 - 0.5 * box_size) * dr, centre[1] + (j + 0.5 - 0.5 * box_size) * dr,
                                 centre[2] + (k + 0.5 - 0.5 * box_size) * dr};
            if (std::fabs(dr * dr * gs.laplacian(r)) > 1e-3) return true;
          }
      return false;
    }, lo, domain_len);
    std::printf(" Wall-clock time generating AMR grid: %10.3E seconds\n", now() - t0);
    const std::vector<int32_t> leaves = tree.ids(true);
    const size_t nin = (size_t)box_size * box_size * box_size;
    std::printf(" Number of boxes used:   %d\n Highest level:          %d\n Number of leaf cells:   %zu\n", tree.highest_id,
                tree.highest_lvl, leaves.size() * nin);

    mg_t mg;
    mg.sides_bc_coords = [&](int, int, const std::vector<double>& coords, std::vector<double>& bc_val, int& bc_type) {
      bc_type = AFMG_BC_DIRICHLET;
      for (size_t n = 0; n < bc_val.size(); ++n) bc_val[n] = gs.value(&coords[3 * n]);
    };
    mg_init(tree, mg);

    // set_initial_condition: rhs = Laplacian of the Gaussians, sol = their value, on the cells of the leaves
    std::vector<double> rhs(leaves.size() * nin), sol(leaves.size() * nin);
    for (size_t b = 0; b < leaves.size(); ++b)
      for (int k = 1; k <= box_size; ++k)
        for (int j = 1; j <= box_size; ++j)
          for (int i = 1; i <= box_size; ++i) {
            const int ijk[3] = {i, j, k};
            double r[3];
            tree.r_cc(leaves[b], ijk, r);
            const size_t q = b * nin + (size_t)(i - 1) + box_size * ((j - 1) + (size_t)box_size * (k - 1));
            rhs[q] = gs.laplacian(r);
            sol[q] = gs.value(r);
          }
    mg.set_cc_interior(AFMG_RHS, leaves, rhs.data());

    std::printf(" Multigrid iteration | max residual | max error\n");
    const int n2 = box_size + 2;
    std::vector<double> phi(leaves.size() * tree.box_len());
    t0 = now();
    for (int mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) {
      mg_fas_fmg(tree, mg, true, mg_iter > 1);
      const double residu = af_tree_maxabs_cc(tree, mg, AFMG_TMP);
      mg.get_cc(AFMG_PHI, leaves, phi.data());  // set_error: err = phi - solution
      double err = 0;
      for (size_t b = 0; b < leaves.size(); ++b)
        for (int k = 1; k <= box_size; ++k)
          for (int j = 1; j <= box_size; ++j)
            for (int i = 1; i <= box_size; ++i) {
              const double p = phi[b * tree.box_len() + (size_t)i + n2 * (j + (size_t)n2 * k)];
              const double s = sol[b * nin + (size_t)(i - 1) + box_size * ((j - 1) + (size_t)box_size * (k - 1))];
              err = std::fmax(err, std::fabs(p - s));
            }
      std::printf("%8d             %14.5E%14.5E\n", mg_iter, residu, err);
    }
    std::printf(" Wall-clock time after %d iterations: %10.3E seconds\n", n_iterations, now() - t0);
    mg_destroy(mg);
  } catch (const afmg::error& e) {
    std::fprintf(stderr, "error stop: %s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
