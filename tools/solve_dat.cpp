// Solve the Poisson problem stored in an afivo .dat file (af_write_tree, version 3) on the GPU, natively: the C++ twin
// of tools/solve_dat.py on include/afmg_dat.hpp.
//     ./solve_dat sim_000100.dat [ndim] [eps-variable] [cycles = 5] [lsf_boundary_value = 0]
// Prints the residual max-norm per cycle the way field_compute tests it (src/m_field.f90:491-524) in the same format
// as the Python tool.  (Not yet run on a device: written after round 1's GPU minutes were spent; the reader and the
// stencil hand-over are checked against the Python side on the CPU, tests/test_cpp_host.py.)
#include <cstdio>
#include <cstdlib>
#include <string>

#include "afmg_dat.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: solve_dat file.dat [ndim] [eps-variable|-] [cycles] [lsf_boundary_value]\n");
    return 2;
  }
  try {
    const int ndim = argc > 2 ? std::atoi(argv[2]) : 0;
    std::string eps = argc > 3 ? argv[3] : "";
    if (eps == "-") eps.clear();
    const int cycles = argc > 4 ? std::atoi(argv[4]) : 5;
    const afmg::dat_t dat = afmg::read_tree(argv[1], ndim);
    const afmg::af_t& t = dat.tree;
    std::printf("%s: NDIM=%d n_cell=%d levels=%d boxes=%zu variables=", argv[1], t.ndim, t.n_cell, t.highest_lvl, t.ids(false).size());
    for (const std::string& n : dat.cc_names) std::printf("%s ", n.c_str());
    std::printf("\n");
    afmg::mg_t mg;
    mg.lsf_boundary_value = argc > 5 ? std::atof(argv[5]) : 0.0;
    afmg::mg_from_dat(dat, mg, "phi", "rhs", eps);
    afmg::mg_fas_fmg(t, mg, true, true);
    std::printf("FMG      residual %.6e\n", afmg::af_tree_maxabs_cc(t, mg, AFMG_TMP));
    for (int i = 0; i < cycles; ++i) {
      afmg::mg_fas_vcycle(t, mg, true);
      std::printf("V-cycle %d residual %.6e\n", i + 1, afmg::af_tree_maxabs_cc(t, mg, AFMG_TMP));
    }
    afmg::field_from_potential(t, mg, -1.0);
    const std::vector<int32_t> leaves = t.ids(true);
    std::vector<double> fld(leaves.size() * t.box_len());
    mg.get_cc(AFMG_FLD, leaves, fld.data());
    double emax = 0;
    for (double v : fld) emax = v > emax ? v : emax;  // ghost cells included: they mirror interior values or are interpolated
    std::printf("max |E| on leaves %.6e\n", emax);
    afmg::mg_destroy(mg);
  } catch (const afmg::error& e) {
    std::fprintf(stderr, "error stop: %s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
