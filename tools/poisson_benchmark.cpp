// poisson_benchmark through the C ABI: the reference's own multigrid benchmark program
// (afivo/examples/poisson_benchmark.f90) with the solver calls going to libafmg.so, as a Fortran caller would
// make them through the shim (fortran/m_af_multigrid_gpu.f90).  Same command line, same measurement loop
// (:119-150), same output lines:
//
//     ./poisson_benchmark_3d n_cell coarse_grid_size max_ref_lvl runtime(s)        (defaults 16 16 2 0.2)
//
// The tree is the one af_init + af_adjust_refinement build for "fully refine up to max_ref_lvl" (:155-165), in the
// reference's conventions: level-1 ids i + (j-1) nx + (k-1) nx ny (m_af_core.f90:436-501), children appended parent
// by parent in af_child_dix order (:1187-1254), neighbours / neighbor_mat with af_phys_boundary = -1 outside the
// unit cube (:595-661).  Build: make -C tools   (g++ -std=c++17 ... -lafmg).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "afmg.h"

namespace {

struct Tree {
  int nc = 0, L = 0, n = 0;
  std::vector<int32_t> lvl_counts, lvl_ids, lvl, ix, parent, children, neighbors, nmat;
  std::vector<std::vector<int32_t>> ids_of_level;  // [L+1]
};

Tree build_uniform(int nc, int coarse, int max_lvl) {
  Tree t;
  t.nc = nc;
  t.L = max_lvl;
  const int nb1 = coarse / nc;
  long total = 0;
  for (int l = 1; l <= max_lvl; ++l) total += (long)nb1 * nb1 * nb1 * (1L << (3 * (l - 1)));
  t.n = (int)total;
  const size_t N = (size_t)t.n + 1;
  t.lvl.assign(N, 0);
  t.ix.assign(N * 3, 0);
  t.parent.assign(N, 0);
  t.children.assign(N * 8, 0);
  t.neighbors.assign(N * 6, 0);
  t.nmat.assign(N * 27, 0);
  t.ids_of_level.assign(max_lvl + 1, {});
  // dense (level, ix) -> id maps: the refinement is full, every position exists
  std::vector<std::vector<int32_t>> at(max_lvl + 1);
  int next = 1;
  {
    at[1].assign((size_t)nb1 * nb1 * nb1, 0);
    for (int k = 1; k <= nb1; ++k)
      for (int j = 1; j <= nb1; ++j)
        for (int i = 1; i <= nb1; ++i) {
          const int id = next++;
          t.lvl[id] = 1;
          t.ix[(size_t)id * 3 + 0] = i;
          t.ix[(size_t)id * 3 + 1] = j;
          t.ix[(size_t)id * 3 + 2] = k;
          at[1][(size_t)(i - 1) + nb1 * ((j - 1) + (size_t)nb1 * (k - 1))] = id;
          t.ids_of_level[1].push_back(id);
        }
  }
  for (int l = 1; l < max_lvl; ++l) {
    const int nbl = nb1 << l;  // boxes per dimension on level l + 1
    at[l + 1].assign((size_t)nbl * nbl * nbl, 0);
    for (int p : t.ids_of_level[l])
      for (int c = 0; c < 8; ++c) {
        const int id = next++;
        t.lvl[id] = l + 1;
        t.parent[id] = p;
        t.children[(size_t)p * 8 + c] = id;
        int q[3];
        for (int d = 0; d < 3; ++d) {
          q[d] = 2 * t.ix[(size_t)p * 3 + d] - 1 + ((c >> d) & 1);  // af_child_dix
          t.ix[(size_t)id * 3 + d] = q[d];
        }
        at[l + 1][(size_t)(q[0] - 1) + nbl * ((q[1] - 1) + (size_t)nbl * (q[2] - 1))] = id;
        t.ids_of_level[l + 1].push_back(id);
      }
  }
  for (int l = 1; l <= max_lvl; ++l) {
    const int nbl = nb1 << (l - 1);
    for (int id : t.ids_of_level[l]) {
      const int32_t* q = &t.ix[(size_t)id * 3];
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int x = q[0] + dx, y = q[1] + dy, z = q[2] + dz;
            const bool out = x < 1 || x > nbl || y < 1 || y > nbl || z < 1 || z > nbl;
            const int v = out ? -1 : at[l][(size_t)(x - 1) + nbl * ((y - 1) + (size_t)nbl * (z - 1))];
            t.nmat[(size_t)id * 27 + (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)] = v;
          }
      for (int nb = 0; nb < 6; ++nb) {
        int d[3] = {0, 0, 0};
        d[nb >> 1] = (nb & 1) ? 1 : -1;
        t.neighbors[(size_t)id * 6 + nb] = t.nmat[(size_t)id * 27 + (d[0] + 1) + 3 * (d[1] + 1) + 9 * (d[2] + 1)];
      }
    }
    t.lvl_counts.push_back((int32_t)t.ids_of_level[l].size());
    t.lvl_ids.insert(t.lvl_ids.end(), t.ids_of_level[l].begin(), t.ids_of_level[l].end());
  }
  return t;
}

void check(int rc, afmg_handle* h, const char* what) {
  if (rc == AFMG_OK) return;
  std::fprintf(stderr, "libafmg error %d in %s: %s\n", rc, what, afmg_last_error(h));
  std::exit(1);  // the reference: error stop
}

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
  int n_cell = 16, coarse_grid_size = 16, max_ref_lvl = 2;
  double runtime = 0.2;
  std::printf(" Running poisson_benchmark_3d\n");
  if (argc >= 2) n_cell = std::atoi(argv[1]);
  else std::printf(" No arguments specified, using default values\n Usage: ./poisson_benchmark_3d n_cell coarse_grid_size max_ref_lvl runtime(s)\n");
  if (argc >= 3) coarse_grid_size = std::atoi(argv[2]);
  if (argc >= 4) max_ref_lvl = std::atoi(argv[3]);
  if (argc >= 5) {
    runtime = std::atof(argv[4]);
    if (runtime < 1e-3) {
      std::printf("Run time should be > 1e-3 seconds\n");
      return 1;
    }
  }
  std::printf(" Box size:            %d\n Coarse grid size:    %d\n Max refinement lvl:  %d\n Run time (s):        %g\n", n_cell,
              coarse_grid_size, max_ref_lvl, runtime);
  if (n_cell < 2 || coarse_grid_size % n_cell != 0 || max_ref_lvl < 1) {
    std::printf("invalid arguments\n");
    return 1;
  }

  double t0 = now();
  Tree t = build_uniform(n_cell, coarse_grid_size, max_ref_lvl);
  std::printf(" Wall-clock time generating AMR grid: %10.3E seconds\n", now() - t0);
  const long cells_finest = (long)t.ids_of_level[max_ref_lvl].size() * n_cell * n_cell * n_cell;
  std::printf(" Number of boxes used:   %d\n Highest level:          %d\n Cells on the finest level: %ld\n", t.n, t.L, cells_finest);

  // mg%sides_bc => af_bc_dirichlet_zero; mg_init
  afmg_opts o{};
  o.ndim = 3;
  o.n_cell = n_cell;
  o.coord_t = AFMG_XYZ;
  o.n_cycle_down = 2;
  o.n_cycle_up = 2;
  o.prolongation_type = AFMG_PROLONG_AUTO;
  o.operator_mask = -1;
  o.device = -1;
  for (int d = 0; d < 3; ++d) {
    o.coarse_grid_size[d] = coarse_grid_size;
    o.dr_base[d] = 1.0 / coarse_grid_size;  // unit cube
  }
  afmg_handle* h = nullptr;
  int rc = afmg_create(&h, &o);
  if (rc != AFMG_OK) {
    std::fprintf(stderr, "afmg_create failed (%d): %s\n", rc, afmg_last_error(nullptr));
    return 1;
  }
  afmg_tree td{};
  td.highest_lvl = t.L;
  td.highest_id = t.n;
  td.lvl_counts = t.lvl_counts.data();
  td.lvl_ids = t.lvl_ids.data();
  td.lvl = t.lvl.data();
  td.ix = t.ix.data();
  td.parent = t.parent.data();
  td.children = t.children.data();
  td.neighbors = t.neighbors.data();
  td.neighbor_mat = t.nmat.data();
  td.r_min = nullptr;
  check(afmg_set_tree(h, &td), h, "afmg_set_tree");
  {
    std::vector<int32_t> bid, bnb, bty;
    for (int id = 1; id <= t.n; ++id)
      for (int nb = 0; nb < 6; ++nb)
        if (t.neighbors[(size_t)id * 6 + nb] == -1) {
          bid.push_back(id);
          bnb.push_back(nb + 1);
          bty.push_back(AFMG_BC_DIRICHLET);
        }
    std::vector<double> bval(bid.size() * (size_t)n_cell * n_cell, 0.0);
    check(afmg_set_bc(h, (int32_t)bid.size(), bid.data(), bnb.data(), bty.data(), bval.data()), h, "afmg_set_bc");
  }
  {  // set_init_cond: rhs = 1 on the interior of every box (the solver reads it on the leaves)
    const auto& leaves = t.ids_of_level[max_ref_lvl];
    std::vector<double> ones(leaves.size() * (size_t)n_cell * n_cell * n_cell, 1.0);
    check(afmg_upload_interior(h, AFMG_RHS, (int32_t)leaves.size(), leaves.data(), ones.data()), h, "afmg_upload_interior");
  }

  check(afmg_fas_fmg(h, 0, 0), h, "afmg_fas_fmg");  // warm-up call

  int n_iterations = 1000, mg_iter = 1;
  double time = 0;
  t0 = now();
  for (mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) {
    check(afmg_fas_fmg(h, 0, mg_iter > 1), h, "afmg_fas_fmg");
    time = now() - t0;
    if (time > 0.2 * runtime) break;
  }
  if (mg_iter > n_iterations) mg_iter = n_iterations;
  n_iterations = (int)std::ceil((runtime / time) * mg_iter);

  t0 = now();
  for (mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) check(afmg_fas_fmg(h, 0, mg_iter > 1), h, "afmg_fas_fmg");
  time = now() - t0;
  std::printf(" Wall-clock time after %d iterations: %10.3E seconds\n", n_iterations, time);
  std::printf(" Per iteration: %10.3E seconds\n", time / n_iterations);

  // beyond the reference's output: the figures of BASELINE.json's metric and the state of the solution
  double cu = 0, res = 0;
  check(afmg_cell_updates(h, 0, 1, &cu), h, "afmg_cell_updates");
  check(afmg_fas_vcycle(h, 1, 0, 1), h, "afmg_fas_vcycle");
  check(afmg_max_abs(h, AFMG_TMP, &res), h, "afmg_max_abs");
  std::printf(" Cell-updates per second (FMG): %10.3E\n Residual max-norm after one more V-cycle: %10.3E\n",
              cu * n_iterations / time, res);
  afmg_destroy(h);
  return 0;
}
