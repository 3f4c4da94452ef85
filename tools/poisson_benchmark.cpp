// poisson_benchmark on the C ABI: the reference's own multigrid benchmark program
// (afivo/examples/poisson_benchmark.f90) written against the C++ mirror of its interface (include/afmg.hpp), the
// solver calls going to libafmg.so as a Fortran caller's would through the shim (fortran/m_af_multigrid_gpu.f90).
// Same command line, same measurement loop (:119-150), same output lines:
//
//     ./poisson_benchmark_3d n_cell coarse_grid_size max_ref_lvl runtime(s)        (defaults 16 16 2 0.2)
//
// Build: make -C tools   (g++ -std=c++17 ... -lafmg).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "afmg.hpp"

namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

int main(int argc, char** argv) {
  using namespace afmg;
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
  try {
    // call af_init(...); do; call af_adjust_refinement(tree, ref_routine, ref_info); ...; end do   (:72-90)
    double t0 = now();
    af_t tree = af_init_fully_refined(n_cell, coarse_grid_size, max_ref_lvl);
    std::printf(" Wall-clock time generating AMR grid: %10.3E seconds\n", now() - t0);
    const std::vector<int32_t> leaves = tree.ids(true);
    const long cells_finest = (long)leaves.size() * n_cell * n_cell * n_cell;
    std::printf(" Number of boxes used:   %d\n Highest level:          %d\n Cells on the finest level: %ld\n", tree.highest_id,
                tree.highest_lvl, cells_finest);

    mg_t mg;                             // mg%i_phi, i_rhs, i_tmp are the library's own variables
    mg.sides_bc = af_bc_dirichlet_zero;  // mg%sides_bc => af_bc_dirichlet_zero
    mg_init(tree, mg);                   // call mg_init(tree, mg)
    {                                    // set_init_cond: box%cc(1:nc, 1:nc, 1:nc, i_rhs) = 1
      std::vector<double> ones(leaves.size() * (size_t)n_cell * n_cell * n_cell, 1.0);
      mg.set_cc_interior(AFMG_RHS, leaves, ones.data());
    }

    mg_fas_fmg(tree, mg, false, false);  // warm-up call

    // test how long cycles take to determine the number of cycles (:119-129)
    int n_iterations = 1000, mg_iter = 1;
    double time = 0;
    t0 = now();
    for (mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) {
      mg_fas_fmg(tree, mg, false, mg_iter > 1);
      time = now() - t0;
      if (time > 0.2 * runtime) break;
    }
    if (mg_iter > n_iterations) mg_iter = n_iterations;
    n_iterations = (int)std::ceil((runtime / time) * mg_iter);

    // the actual benchmarking (:131-150)
    t0 = now();
    for (mg_iter = 1; mg_iter <= n_iterations; ++mg_iter) mg_fas_fmg(tree, mg, false, mg_iter > 1);
    time = now() - t0;
    std::printf(" Wall-clock time after %d iterations: %10.3E seconds\n", n_iterations, time);
    std::printf(" Per iteration: %10.3E seconds\n", time / n_iterations);

    // beyond the reference's output: BASELINE.json's metric and the state of the solution
    double cu = 0;
    mg.check(afmg_cell_updates(mg.h, 0, 1, &cu), "afmg_cell_updates");
    mg_fas_vcycle(tree, mg, true);
    const double res = af_tree_maxabs_cc(tree, mg, AFMG_TMP);
    std::printf(" Cell-updates per second (FMG): %10.3E\n Residual max-norm after one more V-cycle: %10.3E\n",
                cu * n_iterations / time, res);
    mg_destroy(mg);
  } catch (const afmg::error& e) {
    std::fprintf(stderr, "error stop: %s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
