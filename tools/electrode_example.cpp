// electrode_example on the C ABI: afivo/examples/electrode_example.f90 in 2D (Cartesian; with an argument: cylindrical,
// as the reference's command line) written against the C++ mirror (include/afmg.hpp).  A rod electrode
// (0.4, 0.4)-(0.6, 0.6) of radius 0.02 at potential 1 in a grounded unit box; coarse grid 4 x 4 boxes of 8^2 cells,
// refined while lvl < 9 - 2 NDIM and r_min(1) < 0.5 (:84-93); mg%lsf = the rod level set (:109-139, here the
// library's built-in rod shape), mg%lsf_boundary_value = 1, af_bc_dirichlet_zero; ten times mg_fas_fmg +
// mg_compute_phi_gradient with the residual printed as the reference does (:60-66).
//
//     ./electrode_example_2d [cyl] [--dry-run]       Build: make -C tools
//
// --dry-run stops after the host-side set-up (tree, level-set distances, stencils) and prints its sizes and a
// checksum: that half needs no GPU and is compared with the Python mirror in tests/test_cpp_host.py.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "afmg.hpp"

int main(int argc, char** argv) {
  using namespace afmg;
  bool cyl = false, dry = false;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "--dry-run")) dry = true;
    else cyl = true;
  }
  const int box_size = 8, n_iterations = 10, ndim = 2;
  const int cgs[3] = {4 * box_size, 4 * box_size, 1};
  try {
    af_t tree = af_build_tree_nd(
        ndim, box_size, cgs, 9 - 2 * ndim,
        [&](int lvl, const int* ix, const double*) { return lvl < 9 - 2 * ndim && (ix[0] - 1) * (0.25 / (1 << (lvl - 1))) < 0.5; },
        nullptr, nullptr, nullptr, cyl ? AFMG_CYL : AFMG_XYZ);
    afmg_electrode rod{};
    rod.type = AFMG_ELECTRODE_ROD;
    rod.ndim = ndim;
    rod.rod_r0[0] = rod.rod_r0[1] = 0.4;
    rod.rod_r1[0] = rod.rod_r1[1] = 0.6;
    rod.rod_radius = 0.02;
    const lsf_t get_lsf = electrode_lsf(rod);

    mg_t mg;
    mg.sides_bc = af_bc_dirichlet_zero;
    mg.lsf_boundary_value = 1.0;
    if (dry) {
      const stencil_set_t st = mg_build_stencils(tree, mg, nullptr, get_lsf);
      double sum = 0.0;
      for (double v : st.blob) sum += v;
      std::printf("%d %d %zu %zu %zu %.17g\n", tree.highest_lvl, tree.highest_id, st.desc.size(), st.lsf_ids.size(),
                  st.blob.size(), sum);
      return 0;
    }
    mg_init(tree, mg);
    mg_set_operators_tree(tree, mg, nullptr, get_lsf);
    for (int mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) {
      mg_fas_fmg(tree, mg, true, mg_iter > 1);
      mg_compute_phi_gradient(tree, mg, 1.0, true);
      std::printf("%8d%14.5E\n", mg_iter, af_tree_maxabs_cc(tree, mg, AFMG_TMP));
    }
    // beyond the reference's output: extrema of the potential on the leaves and the largest field norm
    const std::vector<int32_t> leaves = tree.ids(true);
    std::vector<double> phi(leaves.size() * tree.box_len()), fld(phi.size());
    mg.get_cc(AFMG_PHI, leaves, phi.data());
    mg.get_cc(AFMG_FLD, leaves, fld.data());
    double lo = 1e300, hi = -1e300, fmax = 0;
    const int n2 = box_size + 2;
    for (size_t b = 0; b < leaves.size(); ++b)
      for (int j = 1; j <= box_size; ++j)
        for (int i = 1; i <= box_size; ++i) {
          const size_t q = b * tree.box_len() + i + (size_t)n2 * j;
          lo = std::fmin(lo, phi[q]);
          hi = std::fmax(hi, phi[q]);
          fmax = std::fmax(fmax, fld[q]);
        }
    std::printf(" min / max potential: %12.5E %12.5E   max field norm: %12.5E\n", lo, hi, fmax);
    mg_destroy(mg);
  } catch (const afmg::error& e) {
    std::fprintf(stderr, "error stop: %s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
