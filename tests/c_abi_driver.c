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
  /* null handles are argument errors everywhere */
  if (afmg_fas_fmg(NULL, 1, 0) != AFMG_ERR_ARG || afmg_field_from_potential(NULL, -1.0) != AFMG_ERR_ARG ||
      afmg_helmholtz_compute(NULL, 0, NULL, 0, 0.0, NULL, NULL) != AFMG_ERR_ARG ||
      afmg_set_lsf_boundary_values(NULL, 0, NULL, NULL) != AFMG_ERR_ARG) { printf("null handle\n"); ++fails; }
  printf(fails ? "FAILED (%d)\n" : "c abi ok\n", fails);
  return fails ? 1 : 0;
}
