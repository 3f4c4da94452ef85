// Builds the explicit stencils of a permittivity + electrode problem with the C++ mirror (afmg::mg_build_stencils on
// a tree from afmg::af_build_tree; no device involved) and dumps them for tests/test_cpp_host.py to compare with the
// Python mirror's (afivo_streamer_b200/stencils.py) on the same tree and functions.
//   cpp_stencil_dump <out.bin> <custom_prolongation 0|1> <dist_method 0|1>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "afmg.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const int nc = 8, cgs[3] = {8, 8, 8};
  afmg::af_t t = afmg::af_build_tree(nc, cgs, 4, [](int, const int* ix, const double*) { return ix[0] == 1 && ix[1] == 1 && ix[2] == 1; });
  // permittivity 1 + x y / 2 + z / 4 at the cell centres (ghost cells included), indexed by box id
  const size_t blen = t.box_len();
  std::vector<double> eps((size_t)(t.highest_id + 1) * blen, 1.0);
  for (int id = 1; id <= t.highest_id; ++id)
    for (int k = 0; k <= nc + 1; ++k)
      for (int j = 0; j <= nc + 1; ++j)
        for (int i = 0; i <= nc + 1; ++i) {
          const int ijk[3] = {i, j, k};
          double r[3];
          t.r_cc(id, ijk, r);
          eps[(size_t)id * blen + i + (nc + 2) * (j + (size_t)(nc + 2) * k)] = 1.0 + 0.5 * (r[0] * r[1]) + 0.25 * r[2];
        }
  afmg::lsf_t lsf = [](const double* r) {
    const double dx = r[0] - 0.2, dy = r[1] - 0.25, dz = r[2] - 0.15;
    return std::sqrt(dx * dx + dy * dy + dz * dz) - 0.11;
  };
  afmg::mg_t mg;
  afmg_lsf_opts o;
  afmg_lsf_opts_default(&o);
  o.dist_method = std::atoi(argv[3]);
  afmg::stencil_set_t st = afmg::mg_build_stencils(t, mg, eps.data(), lsf, &o, std::atoi(argv[2]) != 0);
  std::printf("%zu %zu %zu\n", st.desc.size(), st.blob.size(), st.lsf_ids.size());
  for (const auto& d : st.desc)
    std::printf("D %d %d %d %d %d %d %lld %lld %lld\n", d.box_id, d.tag, d.op_stype, d.cylindrical_gradient, d.prolong_stype,
                d.prolong_shape, (long long)d.op_offset, (long long)d.f_offset, (long long)d.prolong_offset);
  for (size_t b = 0; b < st.lsf_ids.size(); ++b) std::printf("L %d\n", st.lsf_ids[b]);
  FILE* f = std::fopen(argv[1], "wb");
  if (!f) return 3;
  std::fwrite(st.blob.data(), sizeof(double), st.blob.size(), f);
  for (size_t b = 0; b < st.lsf_ids.size(); ++b) {
    std::fwrite(st.lsf_dd[b].data(), sizeof(double), st.lsf_dd[b].size(), f);
    std::fwrite(st.lsf_cells[b].data(), sizeof(double), st.lsf_cells[b].size(), f);
  }
  std::fclose(f);
  return 0;
}
