/* Plain C99 consumer of include/afmg.h: proves that the header is C (not C++), that every type is complete, and
 * that the pure host entry points work without a GPU.  Built and run by tests/test_c_abi.py. */
#include <stdio.h>
#include <string.h>

#include "afmg.h"

int main(void) {
  int fails = 0;
  /* layout helpers: a bijection onto 0 .. (nc+2)^3 - 1 */
  const int nc = 8, n2 = nc + 2, len = afmg_layout_box_len(3, nc);
  if (len != n2 * n2 * n2) { printf("box_len %d\n", len); ++fails; }
  static unsigned char seen[10 * 10 * 10];
  memset(seen, 0, sizeof seen);
  for (int k = 0; k < n2; ++k)
    for (int j = 0; j < n2; ++j)
      for (int i = 0; i < n2; ++i) {
        const int q = afmg_layout_offset(3, nc, i, j, k);
        if (q < 0 || q >= len || seen[q]) { printf("offset (%d,%d,%d) -> %d\n", i, j, k, q); ++fails; }
        else seen[q] = 1;
      }
  /* the partition rule: 3 levels with 1, 8, 64 boxes over 2 ranks */
  const int32_t counts[3] = {1, 8, 64};
  int32_t cuts[3 * 3];
  if (afmg_partition(2, 3, counts, cuts) != AFMG_OK) { printf("partition failed\n"); ++fails; }
  if (cuts[0] != 0 || cuts[2] != 1 || cuts[3] != 0 || cuts[5] != 8 || cuts[6] != 0 || cuts[7] != 32 || cuts[8] != 64) {
    printf("cuts %d %d %d | %d %d %d | %d %d %d\n", cuts[0], cuts[1], cuts[2], cuts[3], cuts[4], cuts[5], cuts[6], cuts[7], cuts[8]);
    ++fails;
  }
  /* no CPU fallback: without a device afmg_create reports AFMG_ERR_CUDA (with one it succeeds) */
  afmg_opts o;
  memset(&o, 0, sizeof o);
  o.ndim = 3; o.n_cell = nc; o.coord_t = AFMG_XYZ; o.n_cycle_down = 2; o.n_cycle_up = 2;
  o.prolongation_type = AFMG_PROLONG_AUTO; o.operator_mask = -1; o.device = -1;
  for (int d = 0; d < 3; ++d) { o.coarse_grid_size[d] = nc; o.dr_base[d] = 1.0 / nc; }
  afmg_handle* h = NULL;
  const int rc = afmg_create(&h, &o);
  if (rc == AFMG_OK) {
    printf("device present: handle created\n");
    /* call order violations are codes, not aborts */
    if (afmg_fas_vcycle(h, 1, 0, 1) != AFMG_ERR_STATE) { printf("vcycle before set_tree\n"); ++fails; }
    if (afmg_compute_phi_gradient(h, -1.0, 1) != AFMG_ERR_STATE) { printf("gradient before set_tree\n"); ++fails; }
    afmg_destroy(h);
  } else if (rc == AFMG_ERR_CUDA) {
    printf("no device: %s\n", afmg_last_error(NULL));
  } else {
    printf("afmg_create returned %d\n", rc);
    ++fails;
  }
  /* host-side builders from C: a rod electrode, one box of distances, its tag and operator (no device involved) */
  {
    afmg_electrode el;
    memset(&el, 0, sizeof el);
    el.type = AFMG_ELECTRODE_ROD; el.ndim = 3; el.rod_radius = 0.1;
    el.rod_r0[0] = el.rod_r0[1] = el.rod_r1[0] = el.rod_r1[1] = 0.5; el.rod_r0[2] = 0.0; el.rod_r1[2] = 0.45;
    if (afmg_electrode_prepare(&el) != AFMG_OK) { printf("electrode_prepare\n"); ++fails; }
    const double p[3] = {0.5, 0.8, 0.2};
    const double lsf = afmg_electrode_lsf(p, &el);
    if (lsf < 0.2 - 1e-14 || lsf > 0.2 + 1e-14) { printf("rod lsf %g\n", lsf); ++fails; }
    afmg_lsf_opts lo;
    afmg_lsf_opts_default(&lo);
    if (lo.dist_method != AFMG_LSF_DIST_LINEAR || lo.tol != 1e-8) { printf("lsf opts\n"); ++fails; }
    static unsigned char mask[8 * 8 * 8];
    static double dd[6 * 8 * 8 * 8], v[7 * 8 * 8 * 8], f[8 * 8 * 8];
    const double r_min[3] = {0.0, 0.0, 0.0}, dr[3] = {0.125, 0.125, 0.125};
    int32_t n_boundary = 0, stype = 0, has_f = 0, cyl = 0;
    if (afmg_build_box_lsf_distances(3, 8, r_min, dr, afmg_electrode_lsf, &el, &lo, NULL, mask, dd, &n_boundary) != AFMG_OK ||
        n_boundary <= 0) { printf("lsf distances: %d boundary cells\n", n_boundary); ++fails; }
    const int32_t tag = afmg_build_box_tag(3, 8, NULL, n_boundary > 0);
    if (tag != AFMG_TAG_LSF_BOX) { printf("tag %d\n", tag); ++fails; }
    if (afmg_build_box_operator(3, 8, AFMG_XYZ, tag, dr, r_min, NULL, dd, v, f, &stype, &has_f, &cyl) != AFMG_OK ||
        stype != 2 || !has_f || cyl) { printf("operator stype %d has_f %d\n", stype, has_f); ++fails; }
    /* every row sums to f = -(the weights moved to the right-hand side): what is missing from the stencil */
    for (int c = 0; c < 512; ++c) {
      double s = 0.0, scale = 0.0;
      for (int m = 0; m < 7; ++m) { s += v[7 * c + m]; scale += v[7 * c + m] < 0 ? -v[7 * c + m] : v[7 * c + m]; }
      const double e = s - f[c];
      if (e > 1e-12 * scale || e < -1e-12 * scale) { printf("row %d: sum %g f %g\n", c, s, f[c]); ++fails; break; }
    }
  }
  /* null handles are argument errors everywhere */
  if (afmg_fas_fmg(NULL, 1, 0) != AFMG_ERR_ARG || afmg_field_from_potential(NULL, -1.0) != AFMG_ERR_ARG ||
      afmg_helmholtz_compute(NULL, 0, NULL, 0, 0.0, NULL, NULL) != AFMG_ERR_ARG ||
      afmg_set_lsf_boundary_values(NULL, 0, NULL, NULL) != AFMG_ERR_ARG) { printf("null handle\n"); ++fails; }
  printf(fails ? "FAILED (%d)\n" : "c abi ok\n", fails);
  return fails ? 1 : 0;
}
