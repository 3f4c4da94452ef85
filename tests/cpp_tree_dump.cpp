// Dumps the tree afmg.hpp builds (af_build_tree) so that tests/test_cpp_host.py can compare it, array by array,
// with the Python builder that follows the reference's conventions (afivo_streamer_b200/tree.py).
//   cpp_tree_dump <kind: uniform|corner|sphere> n_cell cx cy cz max_lvl [ndim = 3 [cylindrical = 0]]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "afmg.hpp"

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const char* kind = argv[1];
  const int nc = std::atoi(argv[2]);
  const int cgs[3] = {std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5])};
  const int lvl = std::atoi(argv[6]);
  const int nd = argc > 7 ? std::atoi(argv[7]) : 3;
  const bool cyl = argc > 8 && std::atoi(argv[8]) != 0;
  afmg::refine_t fn = nullptr;
  if (!std::strcmp(kind, "corner")) fn = [](int, const int* ix, const double*) { return ix[0] == 1 && ix[1] == 1 && ix[2] == 1; };
  if (!std::strcmp(kind, "sphere"))
    fn = [nd](int, const int*, const double* c) {
      double s = 0;
      for (int d = 0; d < nd; ++d) s += (c[d] - 0.4) * (c[d] - 0.4);
      return std::sqrt(s) < 0.45;
    };
  if (nd == 2) {  // the 2D / cylindrical builder: same dump with 4 children, 4 neighbours, a 3 x 3 neighbour matrix
    afmg::af_t t2 = afmg::af_build_tree_nd(2, nc, cgs, lvl, fn, nullptr, nullptr, nullptr, cyl ? AFMG_CYL : AFMG_XYZ);
    std::printf("%d %d\n", t2.highest_lvl, t2.highest_id);
    for (int l = 1; l <= t2.highest_lvl; ++l) {
      std::printf("L %zu", t2.lvl_ids[l].size());
      for (int id : t2.lvl_ids[l]) std::printf(" %d", id);
      std::printf("\n");
    }
    for (int id = 1; id <= t2.highest_id; ++id) {
      std::printf("B %d %d %d %d", t2.lvl[id], t2.ix[id * 2], t2.ix[id * 2 + 1], t2.parent[id]);
      for (int c = 0; c < 4; ++c) std::printf(" %d", t2.children[id * 4 + c]);
      for (int c = 0; c < 4; ++c) std::printf(" %d", t2.neighbors[id * 4 + c]);
      for (int c = 0; c < 9; ++c) std::printf(" %d", t2.neighbor_mat[id * 9 + c]);
      std::printf("\n");
    }
    for (int id = 1; id <= t2.highest_id; ++id)
      std::printf("G %.17g %.17g %.17g %.17g\n", t2.r_min[id * 2], t2.r_min[id * 2 + 1], t2.dr[id * 2], t2.dr[id * 2 + 1]);
    std::printf("coord %d\n", t2.coord_t);
    return 0;
  }
  afmg::af_t t = afmg::af_build_tree(nc, cgs, lvl, fn);
  std::printf("%d %d\n", t.highest_lvl, t.highest_id);
  for (int l = 1; l <= t.highest_lvl; ++l) {
    std::printf("L %zu", t.lvl_ids[l].size());
    for (int id : t.lvl_ids[l]) std::printf(" %d", id);
    std::printf("\n");
  }
  for (int id = 1; id <= t.highest_id; ++id) {
    std::printf("B %d %d %d %d %d", t.lvl[id], t.ix[id * 3], t.ix[id * 3 + 1], t.ix[id * 3 + 2], t.parent[id]);
    for (int c = 0; c < 8; ++c) std::printf(" %d", t.children[id * 8 + c]);
    for (int c = 0; c < 6; ++c) std::printf(" %d", t.neighbors[id * 6 + c]);
    for (int c = 0; c < 27; ++c) std::printf(" %d", t.neighbor_mat[id * 27 + c]);
    std::printf("\n");
  }
  for (int id = 1; id <= t.highest_id; ++id)  // geometry: r_min, dr and the first / last face coordinates of side 3
  {
    std::vector<double> fc;
    t.face_coords(id, 3, fc);
    std::printf("G %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", t.r_min[id * 3], t.r_min[id * 3 + 1],
                t.r_min[id * 3 + 2], t.dr[id * 3], t.dr[id * 3 + 1], t.dr[id * 3 + 2], fc[0], fc[1], fc[2], fc[fc.size() - 3],
                fc[fc.size() - 2], fc[fc.size() - 1]);
  }
  // the mirror refuses to run without a device and reports it the way the reference reports errors
  try {
    afmg::mg_t mg;
    mg.sides_bc = afmg::af_bc_dirichlet_zero;
    afmg::mg_init(t, mg);
    std::printf("device present\n");
  } catch (const afmg::error& e) {
    std::printf("error %d\n", e.code);
  }
  return 0;
}
