// Dumps the tree afmg.hpp builds (af_build_tree) so that tests/test_cpp_host.py can compare it, array by array,
// with the Python builder that follows the reference's conventions (afivo_streamer_b200/tree.py).
//   cpp_tree_dump <kind: uniform|corner|sphere> n_cell cx cy cz max_lvl
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
  afmg::refine_t fn = nullptr;
  if (!std::strcmp(kind, "corner")) fn = [](int, const int* ix, const double*) { return ix[0] == 1 && ix[1] == 1 && ix[2] == 1; };
  if (!std::strcmp(kind, "sphere"))
    fn = [](int, const int*, const double* c) {
      return std::sqrt((c[0] - 0.4) * (c[0] - 0.4) + (c[1] - 0.4) * (c[1] - 0.4) + (c[2] - 0.4) * (c[2] - 0.4)) < 0.45;
    };
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
