// Host-only checks of the C++ mirror's caller-side helpers against the values the Python mirror's tests pin
// (src/m_photoi_helmh.f90:80-136, src/m_field.f90:467-480).  Returns 0 when everything matches.
#include <cmath>
#include <cstdio>

#include "afmg.hpp"

int main() {
  int fails = 0;
  auto near = [&](double a, double b, const char* what) {
    if (std::fabs(a - b) > 1e-14 * std::fabs(b)) {
      std::printf("%s: %.17g vs %.17g\n", what, a, b);
      ++fails;
    }
  };
  const afmg::helmh_params_t b3 = afmg::photoi_helmh_parameters();
  near(b3.lambdas[2], 66755.67 * 0.2, "Bourdon-3 lambda");
  near(b3.coeffs[0], 1117314.935 * 0.04, "Bourdon-3 coeff");
  const afmg::helmh_params_t lu = afmg::photoi_helmh_parameters("Luque", 0.2, 2.0);
  near(lu.lambdas[0], 4425.38 * 2, "Luque lambda");
  near(lu.coeffs[1], 19972.14 * 4, "Luque coeff");
  int caught = 0;
  try { afmg::photoi_helmh_parameters("Luque", 0.2, 1.0, 0.5); } catch (const afmg::error&) { ++caught; }
  try { afmg::photoi_helmh_parameters("Bourdon-2", 0.0); } catch (const afmg::error&) { ++caught; }
  try { afmg::photoi_helmh_parameters("nobody"); } catch (const afmg::error&) { ++caught; }
  if (caught != 3) { std::printf("error stops: %d of 3\n", caught); ++fails; }
  const int cgs[3] = {8, 8, 16};
  const double lo[3] = {0, 0, 0}, hi[3] = {1e-2, 1e-2, 2e-2};
  const afmg::af_t t = afmg::af_build_tree(8, cgs, 3, nullptr, lo, hi);
  const double min_dr = 1.25e-3 / 4;
  near(afmg::field_residual_threshold(t, 0.0, 0.0), 1e-6, "min_residual");
  near(afmg::field_residual_threshold(t, 1e10, 0.0), 1e6, "rhs term");
  near(afmg::field_residual_threshold(t, 0.0, -4e4), 1e-10 * 4e4 / (2e-2 * min_dr), "round-off term");
  near(afmg::field_residual_threshold(t, 0.0, 4e4, true), 1e-8 * 4e4 / (2e-2 * min_dr), "electrode term");
  std::printf(fails ? "FAILED (%d)\n" : "params ok\n", fails);
  return fails ? 1 : 0;
}
