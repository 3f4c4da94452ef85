// Dumps the tree afmg.hpp builds (af_init_fully_refined) so that tests/test_cpp_host.py can compare it, array by
// array, with the Python builder that follows the reference's conventions (afivo_streamer_b200/tree.py).
#include <cstdio>
#include <cstdlib>

#include "afmg.hpp"

int main(int argc, char** argv) {
  const int nc = std::atoi(argv[1]), coarse = std::atoi(argv[2]), lvl = std::atoi(argv[3]);
  (void)argc;
  afmg::af_t t = afmg::af_init_fully_refined(nc, coarse, lvl);
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
