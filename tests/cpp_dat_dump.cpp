// Reads an afivo .dat file with the C++ reader (include/afmg_dat.hpp) and dumps what tests/test_cpp_host.py compares
// with the Python reader: header, per-variable checksums, topology checksums, boundary-condition rows, and the
// stencil set as afmg_set_stencils would receive it.
//   cpp_dat_dump <file.dat> [ndim]
#include <cstdio>
#include <cstdlib>

#include "afmg_dat.hpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    const afmg::dat_t d = afmg::read_tree(argv[1], argc > 2 ? std::atoi(argv[2]) : 0);
    const afmg::af_t& t = d.tree;
    std::printf("H %d %d %d %d %d %d\n", t.ndim, t.highest_lvl, t.highest_id, t.n_cell, t.coord_t, (int)d.ready);
    std::printf("G %d %d %d %d %d %d %.17g %.17g\n", t.coarse_grid_size[0], t.coarse_grid_size[1], t.coarse_grid_size[2],
                (int)t.periodic[0], (int)t.periodic[1], (int)t.periodic[2], t.dr_base[0], t.r_base[t.ndim - 1]);
    for (size_t i = 0; i < d.cc_names.size(); ++i) {
      double s = 0;
      if (d.cc.count((int)i + 1))
        for (double v : d.cc.at((int)i + 1)) s += v;
      std::printf("V %s %d %.17g\n", d.cc_names[i].c_str(), (int)d.cc.count((int)i + 1), s);
    }
    long long topo = 0;
    for (int id = 1; id <= t.highest_id; ++id) {
      topo += (long long)t.lvl[id] * 3 + t.parent[id] * 5 + d.tag[id] * 7;
      for (int q = 0; q < t.ndim; ++q) topo += t.ix[(size_t)id * t.ndim + q] * (11 + q);
      for (int q = 0; q < t.num_children(); ++q) topo += (long long)t.children[(size_t)id * t.num_children() + q] * (q + 1);
      for (int q = 0; q < t.num_neighbors(); ++q) topo += (long long)t.neighbors[(size_t)id * t.num_neighbors() + q] * (q + 2);
    }
    std::printf("T %lld\n", topo);
    double bsum = 0;
    long btypes = 0;
    for (const auto& kv : d.bc) {
      for (double v : kv.second.bc_val) bsum += v;
      for (int32_t v : kv.second.bc_type) btypes += v;
    }
    std::printf("B %zu %ld %.17g\n", d.bc.size(), btypes, bsum);
    const afmg::stencil_set_t st = afmg::dat_stencil_set(d);
    double blob = 0;
    for (double v : st.blob) blob += v;
    std::printf("S %zu %zu %.17g\n", st.desc.size(), st.blob.size(), blob);
    for (const auto& e : st.desc)
      std::printf("D %d %d %d %d %d %d %lld %lld %lld\n", e.box_id, e.tag, e.op_stype, e.cylindrical_gradient, e.prolong_stype,
                  e.prolong_shape, (long long)e.op_offset, (long long)e.f_offset, (long long)e.prolong_offset);
  } catch (const afmg::error& e) {
    std::printf("error %d %s\n", e.code, e.what());
    return 1;
  }
  return 0;
}
