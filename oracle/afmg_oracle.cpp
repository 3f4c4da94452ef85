// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import, link or call
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / CPU baseline.
//
// CPU (C++17 + OpenMP) restatement, routine by routine, of the afivo FAS multigrid path of
// MD-CWI/afivo-streamer.  Every function cites the reference file:line it follows (paths are
// relative to /root/reference).  Expression order follows the Fortran source so that, built
// with -O2 -ffp-contract=off (gfortran -O2 without -march emits no FMA on x86-64), results are
// the ones the reference computes in exact IEEE fp64 dataflow.
//
// PARITY PINNING: the reference cannot be built here (no Fortran compiler; Hypre 2.31.0 is not
// vendored), so this oracle is pinned against the known answers the reference's own tests hold
// for this path (see tests/test_oracle_*.py): 4681 boxes for 5 full 3D levels
// (afivo/tests/answers/test_refinement_3d), zero Neumann field keeps zero ghost cells on a
// corner-refined tree (afivo/tests/test_ghostcell.f90), exactness of ghost cells and
// prolongation for linear fields (afivo/examples/check_ghostcells.f90, check_prolongation.f90),
// and the manufactured Gaussian solution of afivo/examples/poisson_basic.f90.  The coarse-grid
// solve is the one documented deviation: the reference calls Hypre PFMG (iterative, tol 1e-6,
// afivo/src/m_coarse_solver.f90:421-439); here the same BC-folded matrix
// (stencil_handle_boundaries, :442-491) is solved directly (banded LU), i.e. the limit
// tolerance -> 0 of the reference.  "parity unpinned" at the Hypre boundary.
//
// Dimension (2 or 3) is a run-time property of the tree; routines are templated on ND.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---- constants (afivo/src/m_af_types.f90:38-69, 523-544; m_af_stencil.f90:19-36) ----------
constexpr int af_no_box = 0;
constexpr int af_bc_dirichlet = -10, af_bc_neumann = -11, af_bc_continuous = -12,
              af_bc_dirichlet_copy = -13;
constexpr int af_cyl = 2;
constexpr int mg_normal_box = 0, mg_lsf_box = 1, mg_veps_box = 2, mg_ceps_box = 4;
constexpr int mg_cycle_down = 1, mg_cycle_up = 3;
constexpr int stencil_constant = 1, stencil_variable = 2;
constexpr int af_stencil_357 = 1, af_stencil_p234 = 2, af_stencil_p248 = 3;
constexpr int mg_prolong_linear = 17, mg_prolong_sparse = 18, mg_prolong_auto = 19;

// I_FLD: the cell-centred field norm of the callers (i_electric_fld, src/m_streamer.f90:302)
enum { I_PHI = 0, I_RHS = 1, I_TMP = 2, I_EPS = 3, I_FLD = 4, N_VAR = 5 };

// stencil_t (afivo/src/m_af_types.f90:260-282), without the sparse variant
struct Stencil {
  int shape = 0;
  int stype = -1;
  bool cylindrical_gradient = false;
  std::vector<double> c;              // (n_coeff)
  std::vector<double> v;              // (n_coeff, IJK) first index fastest
  std::vector<double> f;              // (IJK) optional
  std::vector<double> bc_correction;  // (IJK) optional
};

// box_t (afivo/src/m_af_types.f90:286-322), the parts the multigrid path touches
struct Box {
  int lvl = 0, tag = 0, parent = 0;
  int ix[3] = {1, 1, 1};
  int children[8] = {0};
  int neighbors[6] = {0};
  int neighbor_mat[27] = {0};
  double dr[3] = {0, 0, 0}, r_min[3] = {0, 0, 0};
  std::vector<double> cc[N_VAR];
  Stencil op, prolong;
  bool has_op = false, has_prolong = false;
  std::vector<double> lsf_dd;  // all_distances(2*ND, IJK); empty = no level-set boundary in box
  // distances from each fine cell centre to its ND+1 coarse prolongation points (mg%lsf_dist evaluated by the
  // caller) for mg_box_prolong_lsf_stencil; dd(1) < 0 marks a cell outside the root mask; empty = not given
  std::vector<double> lsf_pdd;
  // boundary conditions as data: type and values per physical face (sides_bc callbacks in the
  // reference never depend on phi: m_af_ghostcell.f90:615-652, src/m_field.f90:590-670)
  int bc_type[6] = {0};
  std::vector<double> bc_val[6];
  // face-centred field box%fc(nc+1, nc+1[, nc+1], NDIM) of one variable (m_af_core.f90:552), first index
  // fastest; allocated by the first gradient call
  std::vector<double> fc;
  // mg_lsf_boundary_value(box, mg) with mg%lsf_boundary_function evaluated by the caller (m_coarse_solver.f90:
  // 493-510): per interior cell; empty = the constant mg%lsf_boundary_value
  std::vector<double> lsf_bv;
  std::vector<double> lsf_cc;  // cc(IJK, mg%i_lsf) on the interior (sign test of mg_box_lpllsf_gradient); empty: >= 0
  // boundary condition of I_FLD (cc_methods(iv)%bc): default af_bc_neumann_zero (src/m_field.f90:392-393)
  int fld_bc_type[6] = {0};
  std::vector<double> fld_bc_val[6];
};

struct Tree {
  int ndim = 3, nc = 0, coord_t = 1, highest_lvl = 0, highest_id = 0;
  int coarse_grid_size[3] = {1, 1, 1};
  double dr_base[3] = {0, 0, 0}, r_base[3] = {0, 0, 0};
  std::vector<std::vector<int>> ids, leaves, parents;  // index lvl (1-based; [0] unused)
  std::vector<Box> boxes;                              // index id (1-based; [0] unused)
  bool has_eps = false;
  // mg_t (afivo/src/m_af_types.f90:572-665)
  int n_cycle_down = 2, n_cycle_up = 2;
  bool use_corners = false, subtract_mean = false;
  double helmholtz_lambda = 0.0, lsf_boundary_value = 0.0;
  int operator_mask = -1, prolongation_type = mg_prolong_auto;
  bool lsf_use_custom_prolongation = false;
  // coarse solver (replaces Hypre): banded LU of the BC-folded level-1 matrix
  int cs_n = 0, cs_bw = 0;
  int cs_nx[3] = {1, 1, 1};
  std::vector<double> cs_lu;         // (2*bw+1) x n band storage
  std::vector<double> cs_bc_to_rhs;  // (nc^(D-1), 2D, n_boxes1)
  std::vector<double> cs_lsf_fac;    // (nc^D, n_boxes1)
  bool cs_ready = false;
  int n2() const { return nc + 2; }
};

template <int ND>
struct G {  // index helpers for a box with one ghost layer, first index fastest
  int nc, n2, s[3];
  explicit G(int nc_) : nc(nc_), n2(nc_ + 2) {
    s[0] = 1;
    s[1] = n2;
    s[2] = (ND == 3) ? n2 * n2 : 0;
  }
  int size() const { return ND == 3 ? n2 * n2 * n2 : n2 * n2; }
  int at(int i, int j, int k) const { return i + n2 * j + s[2] * k; }
  int klo() const { return ND == 3 ? 1 : 0; }
  int khi() const { return ND == 3 ? nc : 0; }
  int ncell() const { return ND == 3 ? nc * nc * nc : nc * nc; }
  // interior linear index (IJK order, i fastest), 0-based
  int lin(int i, int j, int k) const { return (i - 1) + nc * ((j - 1) + (ND == 3 ? nc * (k - 1) : 0)); }
};

inline bool af_has_children(const Box& b) { return b.children[0] != af_no_box; }

// af_get_child_offset (afivo/src/m_af_types.f90:903-910)
template <int ND>
inline void child_offset(const Box& b, int nc, int* off) {
  for (int d = 0; d < 3; ++d) off[d] = (d < ND) ? ((b.ix[d] - 1) & 1) * (nc >> 1) : 0;
}

// af_cyl_radius_cc, af_cyl_flux_factors, af_cyl_child_weights (m_af_types.f90:1166-1212)
inline double cyl_radius_cc(const Box& b, int i) { return b.r_min[0] + (i - 0.5) * b.dr[0]; }
inline void cyl_flux_factors(const Box& b, int nc, double* rfac /* (2, nc) */) {
  for (int i = 1; i <= nc; ++i) {
    double r = cyl_radius_cc(b, i);
    double inv_r = 1 / r;
    rfac[2 * (i - 1) + 0] = (r - 0.5 * b.dr[0]) * inv_r;
    rfac[2 * (i - 1) + 1] = (r + 0.5 * b.dr[0]) * inv_r;
  }
}

// ---------------------------------------------------------------------------------------------
// Stencil kernels (afivo/src/m_af_stencil.f90)
// ---------------------------------------------------------------------------------------------

// stencil_gsrb_357 (afivo/src/m_af_stencil.f90:838-998)
template <int ND>
void stencil_gsrb_357(Box& box, const Stencil& st, int redblack, int iv, int i_rhs, int nc) {
  G<ND> g(nc);
  double* cc = box.cc[iv].data();
  double* rhs = box.cc[i_rhs].data();
  constexpr int NCF = 2 * ND + 1;
  const int off[6] = {-1, 1, -g.s[1], g.s[1], -g.s[2], g.s[2]};

  if (!st.bc_correction.empty()) {  // :856-859
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) rhs[g.at(i, j, k)] = rhs[g.at(i, j, k)] + st.bc_correction[g.lin(i, j, k)];
  }

  if (ND == 2 && st.cylindrical_gradient) {  // :886-925
    std::vector<double> rfac(2 * nc);
    cyl_flux_factors(box, nc, rfac.data());
    if (st.stype == stencil_constant) {
      const double* c = st.c.data();
      std::vector<double> cc_cyl(NCF * nc), inv_cc1(nc);
      for (int i = 1; i <= nc; ++i) {
        double* q = &cc_cyl[NCF * (i - 1)];
        q[1] = rfac[2 * (i - 1) + 0] * c[1];
        q[2] = rfac[2 * (i - 1) + 1] * c[2];
        q[0] = c[0] - (q[1] - c[1]) - (q[2] - c[2]);
        for (int m = 3; m < NCF; ++m) q[m] = c[m];
        inv_cc1[i - 1] = 1 / q[0];
      }
      for (int j = 1; j <= nc; ++j) {
        int i0 = 2 - ((redblack ^ j) & 1);
        for (int i = i0; i <= nc; i += 2) {
          const double* q = &cc_cyl[NCF * (i - 1)];
          int n = g.at(i, j, 0);
          cc[n] = (rhs[n] - q[1] * cc[n - 1] - q[2] * cc[n + 1] - q[3] * cc[n - g.n2] - q[4] * cc[n + g.n2]) *
                  inv_cc1[i - 1];
        }
      }
    } else {
      for (int j = 1; j <= nc; ++j) {
        int i0 = 2 - ((redblack ^ j) & 1);
        for (int i = i0; i <= nc; i += 2) {
          const double* c = &st.v[NCF * g.lin(i, j, 0)];
          double q[NCF];
          q[1] = rfac[2 * (i - 1) + 0] * c[1];
          q[2] = rfac[2 * (i - 1) + 1] * c[2];
          q[0] = c[0] - (q[1] - c[1]) - (q[2] - c[2]);
          for (int m = 3; m < NCF; ++m) q[m] = c[m];
          int n = g.at(i, j, 0);
          cc[n] = (rhs[n] - q[1] * cc[n - 1] - q[2] * cc[n + 1] - q[3] * cc[n - g.n2] - q[4] * cc[n + g.n2]) / q[0];
        }
      }
    }
  } else if (st.stype == stencil_constant) {  // :927-940 (2D), :956-973 (3D)
    const double* c = st.c.data();
    const double inv_c1 = 1 / c[0];
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j) {
        int i0 = 2 - ((redblack ^ (k + j)) & 1);
        for (int i = i0; i <= nc; i += 2) {
          int n = g.at(i, j, k);
          double acc = rhs[n];
          for (int m = 0; m < 2 * ND; ++m) acc = acc - c[m + 1] * cc[n + off[m]];
          cc[n] = acc * inv_c1;
        }
      }
  } else {  // variable: :942-952 (2D), :974-990 (3D); note the division
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j) {
        int i0 = 2 - ((redblack ^ (k + j)) & 1);
        for (int i = i0; i <= nc; i += 2) {
          const double* c = &st.v[NCF * g.lin(i, j, k)];
          int n = g.at(i, j, k);
          double acc = rhs[n];
          for (int m = 0; m < 2 * ND; ++m) acc = acc - c[m + 1] * cc[n + off[m]];
          cc[n] = acc / c[0];
        }
      }
  }

  if (!st.bc_correction.empty()) {  // :993-996
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) rhs[g.at(i, j, k)] = rhs[g.at(i, j, k)] - st.bc_correction[g.lin(i, j, k)];
  }
}

// stencil_apply_357 (afivo/src/m_af_stencil.f90:367-496)
template <int ND>
void stencil_apply_357(Box& box, const Stencil& st, int iv, int i_out, int nc) {
  G<ND> g(nc);
  const double* cc = box.cc[iv].data();
  double* out = box.cc[i_out].data();
  constexpr int NCF = 2 * ND + 1;
  const int off[6] = {-1, 1, -g.s[1], g.s[1], -g.s[2], g.s[2]};

  if (ND == 2 && st.cylindrical_gradient) {
    std::vector<double> rfac(2 * nc);
    cyl_flux_factors(box, nc, rfac.data());
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        const double* c = (st.stype == stencil_constant) ? st.c.data() : &st.v[NCF * g.lin(i, j, 0)];
        double q[NCF];
        q[1] = rfac[2 * (i - 1) + 0] * c[1];
        q[2] = rfac[2 * (i - 1) + 1] * c[2];
        q[0] = c[0] - (q[1] - c[1]) - (q[2] - c[2]);
        for (int m = 3; m < NCF; ++m) q[m] = c[m];
        int n = g.at(i, j, 0);
        out[n] = q[0] * cc[n] + q[1] * cc[n - 1] + q[2] * cc[n + 1] + q[3] * cc[n - g.n2] + q[4] * cc[n + g.n2];
      }
  } else {
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          const double* c = (st.stype == stencil_constant) ? st.c.data() : &st.v[NCF * g.lin(i, j, k)];
          int n = g.at(i, j, k);
          double acc = c[0] * cc[n];
          for (int m = 0; m < 2 * ND; ++m) acc = acc + c[m + 1] * cc[n + off[m]];
          out[n] = acc;
        }
  }
  if (!st.bc_correction.empty()) {  // :490-493
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) out[g.at(i, j, k)] = out[g.at(i, j, k)] - st.bc_correction[g.lin(i, j, k)];
  }
}

// stencil_prolong_248 / stencil_prolong_234 (afivo/src/m_af_stencil.f90:582-815), add = .true.
template <int ND>
void stencil_prolong(const Box& box_p, Box& box_c, const Stencil& st, int iv, int iv_to, int nc) {
  G<ND> g(nc);
  int ofs[3];
  child_offset<ND>(box_c, nc, ofs);
  const double* P = box_p.cc[iv].data();
  double* C = box_c.cc[iv_to].data();
  const int ncf = (st.shape == af_stencil_p248) ? (1 << ND) : (ND + 1);
  for (int k = g.klo(); k <= g.khi(); ++k) {
    int k_c1 = (ND == 3) ? ofs[2] + ((k + 1) >> 1) : 0;
    int k_c2 = (ND == 3) ? k_c1 + 1 - 2 * (k & 1) : 0;
    for (int j = 1; j <= nc; ++j) {
      int j_c1 = ofs[1] + ((j + 1) >> 1);
      int j_c2 = j_c1 + 1 - 2 * (j & 1);
      for (int i = 1; i <= nc; ++i) {
        int i_c1 = ofs[0] + ((i + 1) >> 1);
        int i_c2 = i_c1 + 1 - 2 * (i & 1);
        const double* c = (st.stype == stencil_constant) ? st.c.data() : &st.v[ncf * g.lin(i, j, k)];
        int n = g.at(i, j, k);
        double acc = C[n];
        if (st.shape == af_stencil_p248) {
          acc = acc + c[0] * P[g.at(i_c1, j_c1, k_c1)];
          acc = acc + c[1] * P[g.at(i_c2, j_c1, k_c1)];
          acc = acc + c[2] * P[g.at(i_c1, j_c2, k_c1)];
          acc = acc + c[3] * P[g.at(i_c2, j_c2, k_c1)];
          if (ND == 3) {
            acc = acc + c[4] * P[g.at(i_c1, j_c1, k_c2)];
            acc = acc + c[5] * P[g.at(i_c2, j_c1, k_c2)];
            acc = acc + c[6] * P[g.at(i_c1, j_c2, k_c2)];
            acc = acc + c[7] * P[g.at(i_c2, j_c2, k_c2)];
          }
        } else {
          acc = acc + c[0] * P[g.at(i_c1, j_c1, k_c1)];
          acc = acc + c[1] * P[g.at(i_c2, j_c1, k_c1)];
          acc = acc + c[2] * P[g.at(i_c1, j_c2, k_c1)];
          if (ND == 3) acc = acc + c[3] * P[g.at(i_c1, j_c1, k_c2)];
        }
        C[n] = acc;
      }
    }
  }
}

// af_restrict_box (afivo/src/m_af_restrict.f90:62-136)
template <int ND>
void af_restrict_box(const Box& box_c, Box& box_p, int iv, bool use_geometry, int nc, int coord_t) {
  G<ND> g(nc);
  const int hnc = nc >> 1;
  int ofs[3];
  child_offset<ND>(box_c, nc, ofs);
  const double* C = box_c.cc[iv].data();
  double* P = box_p.cc[iv].data();
  if (ND == 2) {
    if (coord_t == af_cyl && use_geometry) {
      for (int j = 1; j <= hnc; ++j) {
        int j_c = ofs[1] + j, j_f = 2 * j - 1;
        for (int i = 1; i <= hnc; ++i) {
          int i_c = ofs[0] + i, i_f = 2 * i - 1;
          double tmp = 0.25 * box_p.dr[0] / cyl_radius_cc(box_p, i_c);  // af_cyl_child_weights
          double w1 = 1 - tmp, w2 = 1 + tmp;
          double s1 = 0.0, s2 = 0.0;
          s1 = s1 + C[g.at(i_f, j_f, 0)];
          s1 = s1 + C[g.at(i_f, j_f + 1, 0)];
          s2 = s2 + C[g.at(i_f + 1, j_f, 0)];
          s2 = s2 + C[g.at(i_f + 1, j_f + 1, 0)];
          P[g.at(i_c, j_c, 0)] = 0.25 * (w1 * s1 + w2 * s2);
        }
      }
    } else {
      for (int j = 1; j <= hnc; ++j) {
        int j_c = ofs[1] + j, j_f = 2 * j - 1;
        for (int i = 1; i <= hnc; ++i) {
          int i_c = ofs[0] + i, i_f = 2 * i - 1;
          double s = 0.0;  // sum() in array element order
          for (int dj = 0; dj < 2; ++dj)
            for (int di = 0; di < 2; ++di) s = s + C[g.at(i_f + di, j_f + dj, 0)];
          P[g.at(i_c, j_c, 0)] = 0.25 * s;
        }
      }
    }
  } else {
    for (int k = 1; k <= hnc; ++k) {
      int k_c = ofs[2] + k, k_f = 2 * k - 1;
      for (int j = 1; j <= hnc; ++j) {
        int j_c = ofs[1] + j, j_f = 2 * j - 1;
        for (int i = 1; i <= hnc; ++i) {
          int i_c = ofs[0] + i, i_f = 2 * i - 1;
          double s = 0.0;
          for (int dk = 0; dk < 2; ++dk)
            for (int dj = 0; dj < 2; ++dj)
              for (int di = 0; di < 2; ++di) s = s + C[g.at(i_f + di, j_f + dj, k_f + dk)];
          P[g.at(i_c, j_c, k_c)] = 0.125 * s;
        }
      }
    }
  }
}

// mg_box_rstr_lpl (afivo/src/m_af_multigrid.f90:1227-1242)
template <int ND>
void mg_box_rstr_lpl(Tree& t, const Box& box_c, Box& box_p, int iv) {
  af_restrict_box<ND>(box_c, box_p, iv, iv != I_PHI, t.nc, t.coord_t);
}

// ---------------------------------------------------------------------------------------------
// Ghost cells (afivo/src/m_af_ghostcell.f90, afivo/src/m_af_multigrid.f90:294-621)
// ---------------------------------------------------------------------------------------------
inline int nb_dim(int nb) { return (nb - 1) >> 1; }         // af_neighb_dim - 1
inline bool nb_low(int nb) { return (nb & 1) == 1; }        // af_neighb_low
inline int nb_high_pm(int nb) { return nb_low(nb) ? -1 : 1; }  // af_neighb_high_pm

// af_get_index_bc_outside (afivo/src/m_af_types.f90:967-986) for n_gc = 1
template <int ND>
inline void index_bc_outside(int nb, int nc, int* lo, int* hi) {
  for (int d = 0; d < 3; ++d) {
    lo[d] = (d < ND) ? 1 : 0;
    hi[d] = (d < ND) ? nc : 0;
  }
  int d = nb_dim(nb);
  if (nb_low(nb)) lo[d] = hi[d] = 0;
  else lo[d] = hi[d] = nc + 1;
}

// copy_from_nb (afivo/src/m_af_ghostcell.f90:654-669)
template <int ND>
void copy_from_nb(Box& box, const Box& box_nb, const int* dnb, const int* lo, const int* hi, int iv, int nc) {
  G<ND> g(nc);
  double* A = box.cc[iv].data();
  const double* B = box_nb.cc[iv].data();
  for (int k = lo[2]; k <= hi[2]; ++k)
    for (int j = lo[1]; j <= hi[1]; ++j)
      for (int i = lo[0]; i <= hi[0]; ++i)
        A[g.at(i, j, k)] = B[g.at(i - dnb[0] * nc, j - dnb[1] * nc, (ND == 3) ? k - dnb[2] * nc : 0)];
}

// bc_to_gc (afivo/src/m_af_ghostcell.f90:173-279)
template <int ND>
void bc_to_gc(Box& box, int nb, int iv, const double* bc_val, int bc_type, int nc) {
  G<ND> g(nc);
  double c0, c1, c2;
  switch (bc_type) {
    case af_bc_dirichlet: c0 = 2; c1 = -1; c2 = 0; break;
    case af_bc_neumann: c0 = box.dr[nb_dim(nb)] * nb_high_pm(nb); c1 = 1; c2 = 0; break;
    case af_bc_continuous: c0 = 0; c1 = 2; c2 = -1; break;
    case af_bc_dirichlet_copy: c0 = 1; c1 = 0; c2 = 0; break;
    default: std::fprintf(stderr, "fill_bc: unknown boundary condition %d\n", bc_type); std::abort();
  }
  double* cc = box.cc[iv].data();
  const int d = nb_dim(nb);
  const int sd = g.s[d];
  const int i_gc = nb_low(nb) ? 0 : nc + 1;
  const int i_1 = nb_low(nb) ? 1 : nc;
  const int i_2 = nb_low(nb) ? 2 : nc - 1;
  // transverse dims in increasing order; bc_val(a + (b-1)*nc)
  int td[2], ntd = 0;
  for (int q = 0; q < ND; ++q)
    if (q != d) td[ntd++] = q;
  const int nb_b = (ND == 3) ? nc : 1;
  for (int b = 1; b <= nb_b; ++b)
    for (int a = 1; a <= nc; ++a) {
      int base = a * g.s[td[0]] + ((ND == 3) ? b * g.s[td[1]] : 0);
      double bv = bc_val[(a - 1) + (b - 1) * nc];
      cc[base + i_gc * sd] = c0 * bv + c1 * cc[base + i_1 * sd] + c2 * cc[base + i_2 * sd];
    }
}

// mg_sides_rb (afivo/src/m_af_multigrid.f90:294-461)
template <int ND>
void mg_sides_rb(Tree& t, int id, int nb, int iv) {
  const int nc = t.nc, hnc = nc / 2;
  G<ND> g(nc);
  Box& box = t.boxes[id];
  const int p_id = box.parent;
  const int p_nb_id = t.boxes[p_id].neighbors[nb - 1];
  int co[3];
  child_offset<ND>(box, nc, co);
  const Box& cb = t.boxes[p_nb_id];
  const double* Q = cb.cc[iv].data();
  const int d = nb_dim(nb);
  int td[2], ntd = 0;
  for (int q = 0; q < ND; ++q)
    if (q != d) td[ntd++] = q;
  const int layer = nb_low(nb) ? nc : 1;  // :326-354
  const int w = hnc + 2;
  // tmp(0:hnc+1 [, 0:hnc+1])
  std::vector<double> tmp(ND == 3 ? w * w : w);
  for (int b = 0; b < (ND == 3 ? w : 1); ++b)
    for (int a = 0; a < w; ++a) {
      int n = layer * g.s[d] + (co[td[0]] + a) * g.s[td[0]];
      if (ND == 3) n += (co[td[1]] + b) * g.s[td[1]];
      tmp[a + w * b] = Q[n];
    }
  std::vector<double> gc(ND == 3 ? nc * nc : nc);
  if (ND == 2) {  // :365-369
    for (int i = 1; i <= hnc; ++i) {
      double grad1 = 0.125 * (tmp[i + 1] - tmp[i - 1]);
      gc[2 * i - 2] = tmp[i] - grad1;
      gc[2 * i - 1] = tmp[i] + grad1;
    }
  } else {  // :371-380
    auto T = [&](int i, int j) { return tmp[i + w * j]; };
    auto GC = [&](int i, int j) -> double& { return gc[(i - 1) + nc * (j - 1)]; };
    for (int j = 1; j <= hnc; ++j)
      for (int i = 1; i <= hnc; ++i) {
        double grad1 = 0.125 * (T(i + 1, j) - T(i - 1, j));
        double grad2 = 0.125 * (T(i, j + 1) - T(i, j - 1));
        GC(2 * i - 1, 2 * j - 1) = T(i, j) - grad1 - grad2;
        GC(2 * i, 2 * j - 1) = T(i, j) + grad1 - grad2;
        GC(2 * i - 1, 2 * j) = T(i, j) - grad1 + grad2;
        GC(2 * i, 2 * j) = T(i, j) + grad1 + grad2;
      }
  }
  // :383-459
  const int ix = nb_low(nb) ? 1 : nc;
  const int dix = nb_low(nb) ? 1 : -1;
  double* cc = box.cc[iv].data();
  for (int b = 1; b <= (ND == 3 ? nc : 1); ++b)
    for (int a = 1; a <= nc; ++a) {
      int base = a * g.s[td[0]] + ((ND == 3) ? b * g.s[td[1]] : 0);
      double gcv = gc[(a - 1) + nc * (b - 1)];
      cc[base + (ix - dix) * g.s[d]] =
          0.5 * gcv + 0.75 * cc[base + ix * g.s[d]] - 0.25 * cc[base + (ix + dix) * g.s[d]];
    }
}

// af_gc_prolong_copy + af_prolong_copy (m_af_ghostcell.f90:378-390, m_af_prolong.f90:41-118),
// then mg_sides_rb_extrap (afivo/src/m_af_multigrid.f90:468-621)
template <int ND>
void mg_sides_rb_extrap(Tree& t, int id, int nb, int iv) {
  const int nc = t.nc;
  G<ND> g(nc);
  Box& box = t.boxes[id];
  const Box& box_p = t.boxes[box.parent];
  int lo[3], hi[3], ofs[3];
  index_bc_outside<ND>(nb, nc, lo, hi);
  child_offset<ND>(box, nc, ofs);
  double* cc = box.cc[iv].data();
  const double* P = box_p.cc[iv].data();
  for (int k = lo[2]; k <= hi[2]; ++k) {
    int k_c1 = (ND == 3) ? ofs[2] + ((k + 1) >> 1) : 0;
    for (int j = lo[1]; j <= hi[1]; ++j) {
      int j_c1 = ofs[1] + ((j + 1) >> 1);
      for (int i = lo[0]; i <= hi[0]; ++i) {
        int i_c1 = ofs[0] + ((i + 1) >> 1);
        cc[g.at(i, j, k)] = P[g.at(i_c1, j_c1, k_c1)];
      }
    }
  }
  const int d = nb_dim(nb);
  const int ixn = nb_low(nb) ? 1 : nc;
  const int dixn = nb_low(nb) ? 1 : -1;
  int td[2], ntd = 0;
  for (int q = 0; q < ND; ++q)
    if (q != d) td[ntd++] = q;
  for (int b = 1; b <= (ND == 3 ? nc : 1); ++b) {
    int db = -1 + 2 * (b & 1);
    for (int a = 1; a <= nc; ++a) {
      int da = -1 + 2 * (a & 1);
      int base = a * g.s[td[0]] + ((ND == 3) ? b * g.s[td[1]] : 0);
      int n_gc = base + (ixn - dixn) * g.s[d];
      int n_0 = base + ixn * g.s[d];
      if (ND == 3) {  // :562-565 etc: extrapolation using 2 points
        int n_diag = n_0 + dixn * g.s[d] + da * g.s[td[0]] + db * g.s[td[1]];
        cc[n_gc] = 0.5 * cc[n_gc] + 0.75 * cc[n_0] - 0.25 * cc[n_diag];
      } else {  // :509-512 bilinear extrapolation using 4 points
        int n_n = n_0 + dixn * g.s[d];       // cc(i+di, j)
        int n_t = n_0 + da * g.s[td[0]];     // cc(i, j+dj)
        int n_d = n_n + da * g.s[td[0]];     // cc(i+di, j+dj)
        // the Fortran sums (cc(i+di,j) + cc(i,j+dj)) for nb in x and (cc(i+di,j) + cc(i,j+dj))
        // for nb in y, with (di,dj) the index steps in x and y: keep that operand order
        double s = (d == 0) ? (cc[n_n] + cc[n_t]) : (cc[n_t] + cc[n_n]);
        cc[n_gc] = 0.5 * cc[n_gc] + 1.125 * cc[n_0] - 0.375 * s + 0.125 * cc[n_d];
      }
    }
  }
}

// mg_auto_rb (afivo/src/m_af_multigrid.f90:926-940)
template <int ND>
void mg_auto_rb(Tree& t, int id, int nb, int iv, int op_mask) {
  if ((t.boxes[id].tag & op_mask) == mg_veps_box) mg_sides_rb_extrap<ND>(t, id, nb, iv);
  else mg_sides_rb<ND>(t, id, nb, iv);
}

// af_edge_gc_extrap (afivo/src/m_af_ghostcell.f90:885-924), 3D only
void af_edge_gc_extrap(Box& box, const int* lo, int dim /*0-based*/, int iv, int nc) {
  G<3> g(nc);
  int o1 = (dim + 1) % 3, o2 = (dim + 2) % 3;
  int di[3];
  for (int d = 0; d < 3; ++d) di[d] = 1 - 2 * (lo[d] & 1);
  di[dim] = 0;
  int ia[3] = {lo[0], lo[1], lo[2]}, ib[3] = {lo[0], lo[1], lo[2]}, ic[3], ixx[3] = {lo[0], lo[1], lo[2]};
  ia[o1] += di[o1];
  ib[o2] += di[o2];
  for (int d = 0; d < 3; ++d) ic[d] = lo[d] + di[d];
  double* cc = box.cc[iv].data();
  for (int n = 1; n <= nc; ++n) {
    ia[dim] = ib[dim] = ic[dim] = ixx[dim] = n;
    cc[g.at(ixx[0], ixx[1], ixx[2])] =
        cc[g.at(ia[0], ia[1], ia[2])] + cc[g.at(ib[0], ib[1], ib[2])] - cc[g.at(ic[0], ic[1], ic[2])];
  }
}

// af_corner_gc_extrap (afivo/src/m_af_ghostcell.f90:860-879)
template <int ND>
void af_corner_gc_extrap(Box& box, const int* ix, int iv, int nc) {
  G<ND> g(nc);
  int di[3] = {0, 0, 0};
  for (int d = 0; d < ND; ++d) di[d] = 1 - 2 * (ix[d] & 1);
  double* cc = box.cc[iv].data();
  if (ND == 2) {
    cc[g.at(ix[0], ix[1], 0)] =
        cc[g.at(ix[0] + di[0], ix[1], 0)] + cc[g.at(ix[0], ix[1] + di[1], 0)] - cc[g.at(ix[0] + di[0], ix[1] + di[1], 0)];
  } else {
    cc[g.at(ix[0], ix[1], ix[2])] = cc[g.at(ix[0], ix[1] + di[1], ix[2] + di[2])] +
                                    cc[g.at(ix[0] + di[0], ix[1], ix[2] + di[2])] +
                                    cc[g.at(ix[0] + di[0], ix[1] + di[1], ix[2])] -
                                    2 * cc[g.at(ix[0] + di[0], ix[1] + di[1], ix[2] + di[2])];
  }
}

// tables afivo/src/m_af_types.f90:217-235
const int af_edge_dim[12] = {1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3};
const int af_edge_dir[12][3] = {{0, -1, -1}, {0, 1, -1}, {0, -1, 1}, {0, 1, 1},  {-1, 0, -1}, {1, 0, -1},
                                {-1, 0, 1},  {1, 0, 1},  {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0},  {1, 1, 0}};
const int af_edge_min_ix[12][3] = {{0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {0, 0, 0}, {1, 0, 0},
                                   {0, 0, 1}, {1, 0, 1}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}};

inline int nmat_index(int ndim, const int* d) {
  int m = 0, p = 1;
  for (int q = 0; q < ndim; ++q) {
    m += (d[q] + 1) * p;
    p *= 3;
  }
  return m;
}

// af_gc_box_corner (afivo/src/m_af_ghostcell.f90:125-170)
template <int ND>
void af_gc_box_corner(Tree& t, int id, int iv) {
  const int nc = t.nc;
  Box& box = t.boxes[id];
  if (ND == 3) {
    for (int n = 0; n < 12; ++n) {
      int dim = af_edge_dim[n] - 1;
      int nb_id = box.neighbor_mat[nmat_index(3, af_edge_dir[n])];
      int lo[3];
      for (int d = 0; d < 3; ++d) lo[d] = af_edge_min_ix[n][d] * (nc + 1);
      lo[dim] = 1;
      if (nb_id > af_no_box) {
        int hi[3] = {lo[0], lo[1], lo[2]};
        hi[dim] = nc;
        // dnb = af_neighb_offset(af_nb_adj_edge(:, n)) equals the edge direction
        copy_from_nb<ND>(box, t.boxes[nb_id], af_edge_dir[n], lo, hi, iv, nc);
      } else {
        af_edge_gc_extrap(box, lo, dim, iv, nc);
      }
    }
  }
  for (int n = 0; n < (1 << ND); ++n) {
    int dnb[3] = {0, 0, 0}, lo[3] = {0, 0, 0};
    for (int d = 0; d < ND; ++d) {
      int cd = (n >> d) & 1;
      dnb[d] = 2 * cd - 1;
      lo[d] = cd * (nc + 1);
    }
    int nb_id = box.neighbor_mat[nmat_index(ND, dnb)];
    if (nb_id > af_no_box) copy_from_nb<ND>(box, t.boxes[nb_id], dnb, lo, lo, iv, nc);
    else af_corner_gc_extrap<ND>(box, lo, iv, nc);
  }
}

// af_gc_interp (afivo/src/m_af_ghostcell.f90:394-498): refinement-boundary ghost cells from the coarse
// neighbour of the parent (2 or 3 coarse values) and the box's own first interior layer
template <int ND>
void af_gc_interp(Tree& t, int id, int nb, int iv) {
  const int nc = t.nc;
  G<ND> g(nc);
  Box& box = t.boxes[id];
  const Box& pn = t.boxes[t.boxes[box.parent].neighbors[nb - 1]];
  int off[3];
  child_offset<ND>(box, nc, off);  // af_get_child_offset(box, nb): the entry of dim(nb) is unused here
  const double third = 1 / 3.0, sixth = 1 / 6.0;
  const int d = nb_dim(nb);
  const int ix = nb_low(nb) ? 0 : nc + 1, ix_f = nb_low(nb) ? 1 : nc, ix_c = nb_low(nb) ? nc : 1;
  double* A = box.cc[iv].data();
  const double* P = pn.cc[iv].data();
  int td[2], ntd = 0;
  for (int q = 0; q < ND; ++q)
    if (q != d) td[ntd++] = q;
  if (ND == 2) {
    const int ta = td[0];
    for (int a = 1; a <= nc; ++a) {
      int a_c1 = off[ta] + ((a + 1) >> 1);
      int a_c2 = a_c1 + 1 - 2 * (a & 1);
      int q[3] = {0, 0, 0}, c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, f[3] = {0, 0, 0};
      q[d] = ix; q[ta] = a;
      f[d] = ix_f; f[ta] = a;
      c1[d] = ix_c; c1[ta] = a_c1;
      c2[d] = ix_c; c2[ta] = a_c2;
      A[g.at(q[0], q[1], 0)] = 0.5 * P[g.at(c1[0], c1[1], 0)] + sixth * P[g.at(c2[0], c2[1], 0)] +
                               third * A[g.at(f[0], f[1], 0)];
    }
  } else {
    // transverse dims (ta, tb) in increasing order; order of the coarse terms per case (:448-494):
    //   dim 1: (j_c1,k_c1), (j_c2,k_c1), (j_c1,k_c2);  dim 2: (i_c1,k_c1), (i_c2,k_c1), (i_c1,k_c2);
    //   dim 3: (i_c1,j_c1), (i_c1,j_c2), (i_c2,j_c1)
    const int ta = td[0], tb = td[1];
    for (int b = 1; b <= nc; ++b) {
      int b_c1 = off[tb] + ((b + 1) >> 1);
      int b_c2 = b_c1 + 1 - 2 * (b & 1);
      for (int a = 1; a <= nc; ++a) {
        int a_c1 = off[ta] + ((a + 1) >> 1);
        int a_c2 = a_c1 + 1 - 2 * (a & 1);
        int q[3], f[3], c11[3], c21[3], c12[3];
        q[d] = ix; q[ta] = a; q[tb] = b;
        f[d] = ix_f; f[ta] = a; f[tb] = b;
        c11[d] = ix_c; c11[ta] = a_c1; c11[tb] = b_c1;
        c21[d] = ix_c; c21[ta] = a_c2; c21[tb] = b_c1;
        c12[d] = ix_c; c12[ta] = a_c1; c12[tb] = b_c2;
        const double v11 = P[g.at(c11[0], c11[1], c11[2])];
        const double v21 = P[g.at(c21[0], c21[1], c21[2])];
        const double v12 = P[g.at(c12[0], c12[1], c12[2])];
        const double vf = A[g.at(f[0], f[1], f[2])];
        double r;
        if (d == 2) r = third * v11 + sixth * v12 + sixth * v21 + third * vf;
        else r = third * v11 + sixth * v21 + sixth * v12 + third * vf;
        A[g.at(q[0], q[1], q[2])] = r;
      }
    }
  }
}

// af_gc_box (afivo/src/m_af_ghostcell.f90:64-120)
template <int ND>
void af_gc_box(Tree& t, int id, int iv, bool corners) {
  const int nc = t.nc;
  Box& box = t.boxes[id];
  for (int nb = 1; nb <= 2 * ND; ++nb) {
    int nb_id = box.neighbors[nb - 1];
    if (nb_id > af_no_box) {
      int lo[3], hi[3], dnb[3] = {0, 0, 0};
      index_bc_outside<ND>(nb, nc, lo, hi);
      dnb[nb_dim(nb)] = nb_high_pm(nb);
      copy_from_nb<ND>(box, t.boxes[nb_id], dnb, lo, hi, iv, nc);
    } else if (nb_id == af_no_box) {
      if (iv == I_FLD) af_gc_interp<ND>(t, id, nb, iv);  // cc_methods(i_electric_fld)%rb (src/m_field.f90:392-393)
      else mg_auto_rb<ND>(t, id, nb, iv, t.operator_mask);
    } else if (iv == I_FLD) {
      if (box.fld_bc_val[nb - 1].empty()) {  // af_bc_neumann_zero (m_af_ghostcell.f90:615-625)
        std::vector<double> zero((ND == 3) ? nc * nc : nc, 0.0);
        bc_to_gc<ND>(box, nb, iv, zero.data(), af_bc_neumann, nc);
      } else {
        bc_to_gc<ND>(box, nb, iv, box.fld_bc_val[nb - 1].data(), box.fld_bc_type[nb - 1], nc);
      }
    } else {
      bc_to_gc<ND>(box, nb, iv, box.bc_val[nb - 1].data(), box.bc_type[nb - 1], nc);
    }
  }
  if (corners) af_gc_box_corner<ND>(t, id, iv);
}

// af_gc_lvl (afivo/src/m_af_ghostcell.f90:49-61)
template <int ND>
void af_gc_lvl(Tree& t, int lvl, int iv, bool corners = true) {
  const auto& ids = t.ids[lvl];
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) af_gc_box<ND>(t, ids[i], iv, corners);
}

// af_gc_tree (afivo/src/m_af_ghostcell.f90:25-46)
template <int ND>
void af_gc_tree(Tree& t, int iv, bool corners = true) {
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) af_gc_lvl<ND>(t, lvl, iv, corners);
}

// ---------------------------------------------------------------------------------------------
// Field from potential (afivo/src/m_af_multigrid.f90:1857-2140; caller src/m_field.f90:531-548)
// ---------------------------------------------------------------------------------------------
template <int ND>
struct F {  // face-centred array fc(nc+1, nc+1[, nc+1], ND), 1-based, first index fastest
  int n1;
  explicit F(int nc) : n1(nc + 1) {}
  size_t size() const { return (size_t)ND * (ND == 3 ? n1 * n1 * n1 : n1 * n1); }
  size_t at(int i, int j, int k, int dim /*0-based*/) const {
    size_t per = (ND == 3) ? (size_t)n1 * n1 * n1 : (size_t)n1 * n1;
    return dim * per + (size_t)(i - 1) + (size_t)n1 * ((j - 1) + (ND == 3 ? (size_t)n1 * (k - 1) : 0));
  }
};

// mg_box_lpl_gradient (afivo/src/m_af_multigrid.f90:1901-1999)
template <int ND>
void mg_box_lpl_gradient(Tree& t, Box& box, double fac) {
  const int nc = t.nc;
  G<ND> g(nc);
  F<ND> fi(nc);
  if (box.fc.size() != fi.size()) box.fc.assign(fi.size(), 0.0);
  const double* cc = box.cc[I_PHI].data();
  double inv_dr[3] = {0, 0, 0};
  for (int d = 0; d < ND; ++d) inv_dr[d] = fac / box.dr[d];
  for (int d = 0; d < ND; ++d) {
    int hi[3] = {nc, nc, (ND == 3) ? nc : 0};
    hi[d] = nc + 1;
    const int klo = (ND == 3) ? 1 : 0;
    for (int k = klo; k <= hi[2]; ++k)
      for (int j = 1; j <= hi[1]; ++j)
        for (int i = 1; i <= hi[0]; ++i)
          box.fc[fi.at(i, j, (ND == 3) ? k : 1, d)] = inv_dr[d] * (cc[g.at(i, j, k)] - cc[g.at(i, j, k) - g.s[d]]);
  }
  if ((box.tag & t.operator_mask) == mg_veps_box) {
    // fields at the box boundaries, where eps can change (:1938-1997)
    const double* eps = box.cc[I_EPS].data();
    for (int d = 0; d < ND; ++d) {
      int td[2], ntd = 0;
      for (int q = 0; q < ND; ++q)
        if (q != d) td[ntd++] = q;
      const int nb_b = (ND == 3) ? nc : 1;
      for (int side = 0; side < 2; ++side) {
        const int f = side ? nc + 1 : 1;       // face index
        const int hi_c = side ? nc + 1 : 1;    // cell on the high side of the face
        const int gc = side ? nc + 1 : 0;      // the ghost cell (its eps multiplies)
        for (int b = 1; b <= nb_b; ++b)
          for (int a = 1; a <= nc; ++a) {
            int q[3] = {0, 0, 0};
            q[td[0]] = a;
            if (ND == 3) q[td[1]] = b;
            int qh[3] = {q[0], q[1], q[2]}, ql[3] = {q[0], q[1], q[2]}, qg[3] = {q[0], q[1], q[2]}, qf[3] = {q[0], q[1], q[2]};
            qh[d] = hi_c;
            ql[d] = hi_c - 1;
            qg[d] = gc;
            qf[d] = f;
            const int ih = g.at(qh[0], qh[1], qh[2]), il = g.at(ql[0], ql[1], ql[2]), ig = g.at(qg[0], qg[1], qg[2]);
            // low side : 2*inv_dr*(cc(1)-cc(0))*eps(0)/(eps(1)+eps(0));  high: .. *eps(nc+1)/(eps(nc+1)+eps(nc))
            box.fc[fi.at(qf[0], qf[1], (ND == 3) ? qf[2] : 1, d)] =
                2 * inv_dr[d] * (cc[ih] - cc[il]) * eps[ig] / (eps[ih] + eps[il]);
          }
      }
    }
  }
}

// mg_lsf_boundary_value (afivo/src/m_coarse_solver.f90:493-510) at interior cell L (IJK order)
inline double lsf_boundary_value_at(const Tree& t, const Box& box, int L) {
  return box.lsf_bv.empty() ? t.lsf_boundary_value : box.lsf_bv[L];
}

// mg_box_lpllsf_gradient (afivo/src/m_af_multigrid.f90:2055-2137).  The sparse distance stencil lists, in
// IJK order, the cells with any(dd < 1) (store_lsf_distance_matrix :1075-1080); the boundary value is the
// constant mg%lsf_boundary_value (mg_lsf_boundary_value, m_coarse_solver.f90:493-510).
template <int ND>
void mg_box_lpllsf_gradient(Tree& t, Box& box, double fac) {
  mg_box_lpl_gradient<ND>(t, box, fac);
  const int nc = t.nc;
  G<ND> g(nc);
  F<ND> fi(nc);
  const double* cc = box.cc[I_PHI].data();
  double inv_dr[3] = {0, 0, 0};
  for (int d = 0; d < ND; ++d) inv_dr[d] = fac / box.dr[d];
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        const int L = g.lin(i, j, k);
        const double bc = lsf_boundary_value_at(t, box, L);
        const double* dd = &box.lsf_dd[(size_t)2 * ND * L];
        bool any = false;
        for (int n = 0; n < 2 * ND; ++n) any = any || dd[n] < 1.0;
        if (!any) continue;
        const bool pos = box.lsf_cc.empty() || box.lsf_cc[L] >= 0;
        const double phi = cc[g.at(i, j, k)];
        for (int d = 0; d < ND; ++d) {
          int q[3] = {i, j, (ND == 3) ? k : 1};
          if (dd[2 * d] < 1 && pos) box.fc[fi.at(q[0], q[1], q[2], d)] = inv_dr[d] * (phi - bc) / dd[2 * d];
          q[d] += 1;
          if (dd[2 * d + 1] < 1 && pos) box.fc[fi.at(q[0], q[1], q[2], d)] = inv_dr[d] * (bc - phi) / dd[2 * d + 1];
        }
      }
}

// mg_box_field_norm (afivo/src/m_af_multigrid.f90:2023-2051)
template <int ND>
void mg_box_field_norm(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  F<ND> fi(nc);
  double* out = box.cc[I_FLD].data();
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        const int kk = (ND == 3) ? k : 1;
        double a = box.fc[fi.at(i, j, kk, 0)] + box.fc[fi.at(i + 1, j, kk, 0)];
        double b = box.fc[fi.at(i, j, kk, 1)] + box.fc[fi.at(i, j + 1, kk, 1)];
        double s = a * a + b * b;
        if (ND == 3) {
          double c = box.fc[fi.at(i, j, kk, 2)] + box.fc[fi.at(i, j, kk + 1, 2)];
          s = s + c * c;
        }
        out[g.at(i, j, k)] = 0.5 * std::sqrt(s);
      }
}

// mg_compute_phi_gradient (afivo/src/m_af_multigrid.f90:1857-1898)
template <int ND>
void mg_compute_phi_gradient(Tree& t, double fac, bool with_norm) {
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) {
    const auto& ids = t.ids[lvl];
#pragma omp parallel for
    for (int i = 0; i < (int)ids.size(); ++i) {
      Box& box = t.boxes[ids[i]];
      const int tag = box.tag & t.operator_mask;
      if ((tag & mg_lsf_box) && tag != mg_veps_box && !af_has_children(box) && !box.lsf_dd.empty())
        mg_box_lpllsf_gradient<ND>(t, box, fac);
      else
        mg_box_lpl_gradient<ND>(t, box, fac);
      if (with_norm) mg_box_field_norm<ND>(t, box);
    }
  }
}

// mg_compute_field_norm (afivo/src/m_af_multigrid.f90:2002-2020)
template <int ND>
void mg_compute_field_norm(Tree& t) {
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) {
    const auto& ids = t.ids[lvl];
#pragma omp parallel for
    for (int i = 0; i < (int)ids.size(); ++i) mg_box_field_norm<ND>(t, t.boxes[ids[i]]);
  }
}

// ---------------------------------------------------------------------------------------------
// Stencil builders (afivo/src/m_af_multigrid.f90:1100-1185, 1246-1388, 1493-1532, 1782-1854)
// ---------------------------------------------------------------------------------------------

// af_stencil_try_constant (afivo/src/m_af_stencil.f90:1001-1028)
void stencil_try_constant(Stencil& st, int ncf, int ncell, double abs_tol) {
  for (int n = 0; n < ncell; ++n)
    for (int m = 0; m < ncf; ++m)
      if (std::fabs(st.v[ncf * n + m] - st.v[m]) > abs_tol) return;
  st.stype = stencil_constant;
  st.c.assign(st.v.begin(), st.v.begin() + ncf);
  st.v.clear();
}

// mg_box_lpl_stencil (afivo/src/m_af_multigrid.f90:1246-1264)
template <int ND>
void mg_box_lpl_stencil(Tree& t, Box& box) {
  Stencil& st = box.op;
  st = Stencil();
  st.shape = af_stencil_357;
  st.stype = stencil_constant;
  st.cylindrical_gradient = (t.coord_t == af_cyl);
  st.c.assign(2 * ND + 1, 0.0);
  for (int d = 0; d < ND; ++d) {
    double inv_dr2 = 1 / (box.dr[d] * box.dr[d]);
    st.c[2 * d + 1] = inv_dr2;
    st.c[2 * d + 2] = inv_dr2;
  }
  double s = 0.0;
  for (int m = 1; m < 2 * ND + 1; ++m) s = s + st.c[m];
  st.c[0] = -s - t.helmholtz_lambda;
}

// mg_box_lpld_stencil (afivo/src/m_af_multigrid.f90:1493-1532)
template <int ND>
void mg_box_lpld_stencil(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  Stencil& st = box.op;
  st = Stencil();
  st.shape = af_stencil_357;
  st.stype = stencil_variable;
  st.cylindrical_gradient = (t.coord_t == af_cyl);
  constexpr int NCF = 2 * ND + 1;
  st.v.assign((size_t)NCF * g.ncell(), 0.0);
  double idr2[6];
  for (int d = 0; d < ND; ++d) idr2[2 * d] = idr2[2 * d + 1] = 1 / (box.dr[d] * box.dr[d]);
  const double* eps = box.cc[I_EPS].data();
  const int off[6] = {-1, 1, -g.s[1], g.s[1], -g.s[2], g.s[2]};
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        int n = g.at(i, j, k);
        double a0 = eps[n];
        double* v = &st.v[(size_t)NCF * g.lin(i, j, k)];
        double s = 0.0;
        for (int m = 0; m < 2 * ND; ++m) {
          double a = eps[n + off[m]];
          v[m + 1] = idr2[m] * 2 * a0 * a / (a0 + a);
          s = s + v[m + 1];
        }
        v[0] = -s;
      }
  stencil_try_constant(st, NCF, g.ncell(), 2.220446049250313e-16);
}

// mg_box_lpld_lsf_stencil (afivo/src/m_af_multigrid.f90:1535-1623): variable permittivity AND a level-set
// boundary in the box; distances given as data (all_distances(2*ND, IJK), 1 = no boundary)
template <int ND>
void mg_box_lpld_lsf_stencil(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  Stencil& st = box.op;
  st = Stencil();
  st.shape = af_stencil_357;
  st.stype = stencil_variable;
  st.cylindrical_gradient = (t.coord_t == af_cyl);
  constexpr int NCF = 2 * ND + 1;
  st.v.assign((size_t)NCF * g.ncell(), 0.0);
  st.f.assign(g.ncell(), 0.0);
  double dr2[3];
  for (int d = 0; d < ND; ++d) dr2[d] = box.dr[d] * box.dr[d];
  const double* eps = box.cc[I_EPS].data();
  const int off[6] = {-1, 1, -g.s[1], g.s[1], -g.s[2], g.s[2]};
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        const int n = g.at(i, j, k), L = g.lin(i, j, k);
        const double* dd = &box.lsf_dd[(size_t)2 * ND * L];
        double* v = &st.v[(size_t)NCF * L];
        const double a0 = eps[n];
        // generalized Laplacian for neighbours at distance dd * dx (:1594-1601)
        for (int d = 0; d < ND; ++d) {
          v[1 + 2 * d] = 1 / (0.5 * dr2[d] * (dd[2 * d] + dd[2 * d + 1]) * dd[2 * d]);
          v[2 + 2 * d] = 1 / (0.5 * dr2[d] * (dd[2 * d] + dd[2 * d + 1]) * dd[2 * d + 1]);
        }
        // permittivity: constant up to an electrode boundary (:1603-1608)
        double s = 0.0;
        for (int m = 0; m < 2 * ND; ++m) {
          const double a = (dd[m] < 1.0) ? a0 : eps[n + off[m]];
          v[m + 1] = v[m + 1] * 2 * a0 * a / (a0 + a);
          s = s + v[m + 1];
        }
        v[0] = -s;
        for (int m = 0; m < 2 * ND; ++m)  // internal boundaries move to the right-hand side (:1612-1618)
          if (dd[m] < 1.0) {
            st.f[L] = st.f[L] - v[m + 1];
            v[m + 1] = 0.0;
          }
      }
  stencil_try_constant(st, NCF, g.ncell(), 2.220446049250313e-16);
}

// mg_box_lsf_stencil (afivo/src/m_af_multigrid.f90:1782-1854), distances given as data
template <int ND>
void mg_box_lsf_stencil(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  Stencil& st = box.op;
  st = Stencil();
  st.shape = af_stencil_357;
  st.stype = stencil_variable;
  st.cylindrical_gradient = false;
  constexpr int NCF = 2 * ND + 1;
  st.v.assign((size_t)NCF * g.ncell(), 0.0);
  st.f.assign(g.ncell(), 0.0);
  double dr2[3];
  for (int d = 0; d < ND; ++d) dr2[d] = box.dr[d] * box.dr[d];
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        int L = g.lin(i, j, k);
        const double* dd = &box.lsf_dd[(size_t)2 * ND * L];
        double* v = &st.v[(size_t)NCF * L];
        for (int d = 0; d < ND; ++d) {
          v[1 + 2 * d] = 1 / (0.5 * dr2[d] * (dd[2 * d] + dd[2 * d + 1]) * dd[2 * d]);
          v[2 + 2 * d] = 1 / (0.5 * dr2[d] * (dd[2 * d] + dd[2 * d + 1]) * dd[2 * d + 1]);
        }
        if (ND == 2 && t.coord_t == af_cyl) {
          double tmp = 1 / (box.dr[0] * (dd[0] + dd[1]) * cyl_radius_cc(box, i));
          v[1] = v[1] - tmp;
          v[2] = v[2] + tmp;
        }
        double s = 0.0;
        for (int m = 1; m < NCF; ++m) s = s + v[m];
        v[0] = -s;
        for (int n = 0; n < 2 * ND; ++n)
          if (dd[n] < 1.0) {
            st.f[L] = st.f[L] - v[n + 1];
            v[n + 1] = 0.0;
          }
      }
}

// mg_box_prolong_linear_stencil / _sparse_stencil (afivo/src/m_af_multigrid.f90:1267-1304)
template <int ND>
void mg_box_prolong_const(Box& box, bool sparse) {
  Stencil& st = box.prolong;
  st = Stencil();
  st.stype = stencil_constant;
  if (!sparse) {
    st.shape = af_stencil_p248;
    if (ND == 2) st.c = {9 / 16.0, 3 / 16.0, 3 / 16.0, 1 / 16.0};
    else st.c = {27 / 64.0, 9 / 64.0, 9 / 64.0, 3 / 64.0, 9 / 64.0, 3 / 64.0, 3 / 64.0, 1 / 64.0};
  } else {
    st.shape = af_stencil_p234;
    if (ND == 2) st.c = {0.5, 0.25, 0.25};
    else st.c = {0.25, 0.25, 0.25, 0.25};
  }
}

// mg_box_prolong_eps_stencil (afivo/src/m_af_multigrid.f90:1308-1388)
template <int ND>
void mg_box_prolong_eps_stencil(Tree& t, Box& box, const Box& box_p) {
  const int nc = t.nc;
  G<ND> g(nc);
  Stencil& st = box.prolong;
  st = Stencil();
  st.shape = af_stencil_p234;
  st.stype = stencil_variable;
  const int ncf = ND + 1;
  st.v.assign((size_t)ncf * g.ncell(), 0.0);
  int ofs[3];
  child_offset<ND>(box, nc, ofs);
  const double* E = box_p.cc[I_EPS].data();
  const double third = 1 / 3.0;
  for (int k = g.klo(); k <= g.khi(); ++k) {
    int k_c1 = (ND == 3) ? ofs[2] + ((k + 1) >> 1) : 0;
    int k_c2 = (ND == 3) ? k_c1 + 1 - 2 * (k & 1) : 0;
    for (int j = 1; j <= nc; ++j) {
      int j_c1 = ofs[1] + ((j + 1) >> 1);
      int j_c2 = j_c1 + 1 - 2 * (j & 1);
      for (int i = 1; i <= nc; ++i) {
        int i_c1 = ofs[0] + ((i + 1) >> 1);
        int i_c2 = i_c1 + 1 - 2 * (i & 1);
        double a0 = E[g.at(i_c1, j_c1, k_c1)];
        double a[3];
        a[0] = E[g.at(i_c2, j_c1, k_c1)];
        a[1] = E[g.at(i_c1, j_c2, k_c1)];
        if (ND == 3) a[2] = E[g.at(i_c1, j_c1, k_c2)];
        double* v = &st.v[(size_t)ncf * g.lin(i, j, k)];
        double s = 0.0;
        if (ND == 2) {
          for (int m = 0; m < 2; ++m) s = s + a0 / (a0 + a[m]);
          v[0] = 0.5 * s;
        } else {
          for (int m = 0; m < 3; ++m) s = s + (a0 - 0.5 * a[m]) / (a0 + a[m]);
          v[0] = third * s;
        }
        for (int m = 0; m < ND; ++m) v[1 + m] = 0.5 * a[m] / (a0 + a[m]);
      }
    }
  }
  stencil_try_constant(st, ncf, g.ncell(), 2.220446049250313e-16);
}

// mg_box_prolong_lsf_stencil (afivo/src/m_af_multigrid.f90:1392-1482): prolongation weights that shrink towards
// an electrode surface; the distances to the ND+1 coarse points are given as data (lsf_pdd)
template <int ND>
void mg_box_prolong_lsf_stencil(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  Stencil& st = box.prolong;
  st = Stencil();
  st.shape = af_stencil_p234;
  st.stype = stencil_variable;
  const int ncf = ND + 1;
  st.v.assign((size_t)ncf * g.ncell(), 0.0);
  for (int L = 0; L < g.ncell(); ++L) {
    double* v = &st.v[(size_t)ncf * L];
    if (ND == 2) { v[0] = 0.5; v[1] = 0.25; v[2] = 0.25; }
    else { v[0] = v[1] = v[2] = v[3] = 0.25; }
    const double* dd = &box.lsf_pdd[(size_t)ncf * L];
    if (dd[0] < 0) continue;  // not in the root mask
    if (ND == 2) {
      v[0] = 2 * dd[1] * dd[2];
      v[1] = dd[0] * dd[2];
      v[2] = dd[0] * dd[1];
    } else {
      v[0] = dd[1] * dd[2] * dd[3];
      v[1] = dd[0] * dd[2] * dd[3];
      v[2] = dd[0] * dd[1] * dd[3];
      v[3] = dd[0] * dd[1] * dd[2];
    }
    double s = 0.0;
    for (int m = 0; m < ncf; ++m) s = s + v[m];
    for (int m = 0; m < ncf; ++m) v[m] = v[m] / s;
    for (int m = 0; m < ncf; ++m)
      if (dd[m] < 1) v[m] = 0.0;
  }
  stencil_try_constant(st, ncf, g.ncell(), 2.220446049250313e-16);
}

// mg_set_box_tag (afivo/src/m_af_multigrid.f90:1100-1145); lsf presence given by lsf_dd data
template <int ND>
void mg_set_box_tag(Tree& t, Box& box) {
  box.tag = mg_normal_box;
  if (!box.lsf_dd.empty()) box.tag += mg_lsf_box;
  if (t.has_eps) {
    const auto& e = box.cc[I_EPS];
    double a = *std::min_element(e.begin(), e.end());
    double b = *std::max_element(e.begin(), e.end());
    if (b > a) box.tag += mg_veps_box;
    else if (std::max(std::fabs(a - 1), std::fabs(b - 1)) > 1e-8) box.tag += mg_ceps_box;
  }
}

// mg_lsf_boundary_value (afivo/src/m_coarse_solver.f90:494-511): constant value variant
// mg_set_operators_lvl (afivo/src/m_af_multigrid.f90:1147-1185), mg_store_operator_stencil
// (:823-859) with operator_type = mg_auto_operator, mg_store_prolongation_stencil (:862-903)
template <int ND>
void mg_set_operators_lvl(Tree& t, int lvl, bool force) {
  const auto& ids = t.ids[lvl];
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) {
    Box& box = t.boxes[ids[i]];
    if (force || !box.has_op) {
      mg_set_box_tag<ND>(t, box);
      switch (box.tag & t.operator_mask) {
        case mg_normal_box: mg_box_lpl_stencil<ND>(t, box); break;
        case mg_lsf_box: mg_box_lsf_stencil<ND>(t, box); break;
        case mg_veps_box:
        case mg_ceps_box: mg_box_lpld_stencil<ND>(t, box); break;
        case mg_veps_box + mg_lsf_box:
        case mg_ceps_box + mg_lsf_box: mg_box_lpld_lsf_stencil<ND>(t, box); break;
        default: std::fprintf(stderr, "mg_store_operator_stencil: unknown box tag\n"); std::abort();
      }
      box.has_op = true;
    }
    if (!box.op.f.empty()) {  // :1171-1174
      box.op.bc_correction.resize(box.op.f.size());
      for (size_t n = 0; n < box.op.f.size(); ++n) box.op.bc_correction[n] = box.op.f[n] * lsf_boundary_value_at(t, box, (int)n);
    }
    if (lvl > 1 && (force || !box.has_prolong)) {
      const Box& box_p = t.boxes[box.parent];
      switch (t.prolongation_type) {
        case mg_prolong_linear: mg_box_prolong_const<ND>(box, false); break;
        case mg_prolong_sparse: mg_box_prolong_const<ND>(box, true); break;
        default:
          switch (box.tag & t.operator_mask) {
            case mg_normal_box:
            case mg_ceps_box: mg_box_prolong_const<ND>(box, false); break;
            case mg_lsf_box:
            case mg_ceps_box + mg_lsf_box:
              if (t.lsf_use_custom_prolongation && !box.lsf_pdd.empty()) mg_box_prolong_lsf_stencil<ND>(t, box);
              else mg_box_prolong_const<ND>(box, false);
              break;
            case mg_veps_box:
            case mg_veps_box + mg_lsf_box: mg_box_prolong_eps_stencil<ND>(t, box, box_p); break;
            default: std::fprintf(stderr, "mg_store_prolongation_stencil: unknown box tag\n"); std::abort();
          }
      }
      box.has_prolong = true;
    }
  }
}

template <int ND>
void mg_set_operators_tree(Tree& t, bool force = false) {
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) mg_set_operators_lvl<ND>(t, lvl, force);
}

// ---------------------------------------------------------------------------------------------
// Coarse grid solver (afivo/src/m_coarse_solver.f90), Hypre replaced by a banded direct solve
// ---------------------------------------------------------------------------------------------

// stencil_get_357 (afivo/src/m_af_stencil.f90:1051-1087)
template <int ND>
void stencil_get_357(const Box& box, const Stencil& st, int nc, std::vector<double>& v) {
  G<ND> g(nc);
  constexpr int NCF = 2 * ND + 1;
  v.resize((size_t)NCF * g.ncell());
  if (st.stype == stencil_constant) {
    for (int n = 0; n < g.ncell(); ++n)
      for (int m = 0; m < NCF; ++m) v[NCF * n + m] = st.c[m];
  } else {
    v = st.v;
  }
  if (ND == 2 && st.cylindrical_gradient) {
    std::vector<double> rfac(2 * nc);
    cyl_flux_factors(box, nc, rfac.data());
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        double* c = &v[NCF * g.lin(i, j, 0)];
        double q[NCF];
        q[1] = rfac[2 * (i - 1) + 0] * c[1];
        q[2] = rfac[2 * (i - 1) + 1] * c[2];
        q[0] = c[0] - (q[1] - c[1]) - (q[2] - c[2]);
        for (int m = 3; m < NCF; ++m) q[m] = c[m];
        for (int m = 0; m < NCF; ++m) c[m] = q[m];
      }
  }
}

// coarse_solver_initialize + hypre_set_matrix + stencil_handle_boundaries
// (afivo/src/m_coarse_solver.f90:71-194, 442-491)
template <int ND>
int coarse_solver_initialize(Tree& t) {
  const int nc = t.nc;
  G<ND> g(nc);
  constexpr int NCF = 2 * ND + 1;
  const auto& ids1 = t.ids[1];
  const int nb1 = (int)ids1.size();
  int nx[3] = {1, 1, 1};
  for (int d = 0; d < ND; ++d) nx[d] = t.coarse_grid_size[d];
  const int n = nx[0] * nx[1] * nx[2];
  // periodic dimensions (the level-1 boxes at the domain edge then have a real neighbour there) couple
  // the first and last cells: the band becomes the whole matrix
  bool any_periodic = false;
  for (int ib = 0; ib < nb1; ++ib)
    for (int nb = 1; nb <= 2 * ND; ++nb) {
      const Box& b = t.boxes[ids1[ib]];
      const int d = nb_dim(nb);
      const bool at_edge = nb_low(nb) ? (b.ix[d] == 1) : (b.ix[d] * nc == nx[d]);
      if (at_edge && b.neighbors[nb - 1] > af_no_box) any_periodic = true;
    }
  const int bw = any_periodic ? n - 1 : ((ND == 3) ? nx[0] * nx[1] : nx[0]);
  const int nface = (ND == 3) ? nc * nc : nc;
  t.cs_n = n;
  t.cs_bw = bw;
  for (int d = 0; d < 3; ++d) t.cs_nx[d] = nx[d];
  t.cs_bc_to_rhs.assign((size_t)nface * 2 * ND * nb1, 0.0);
  t.cs_lsf_fac.assign((size_t)g.ncell() * nb1, 0.0);
  const int ldab = 2 * bw + 1;
  std::vector<double>& ab = t.cs_lu;
  ab.assign((size_t)ldab * n, 0.0);  // ab[(bw + r - c) + ldab * c] = A(r, c)
  const int gstride[3] = {1, nx[0], nx[0] * nx[1]};

  for (int ib = 0; ib < nb1; ++ib) {
    Box& box = t.boxes[ids1[ib]];
    std::vector<double> full;
    stencil_get_357<ND>(box, box.op, nc, full);
    double* bc_to_rhs = &t.cs_bc_to_rhs[(size_t)nface * 2 * ND * ib];
    // stencil_handle_boundaries
    for (int nb = 1; nb <= 2 * ND; ++nb) {
      if (box.neighbors[nb - 1] >= af_no_box) continue;
      const int d = nb_dim(nb);
      const int bc_type = box.bc_type[nb - 1];
      int td[2], ntd = 0;
      for (int q = 0; q < ND; ++q)
        if (q != d) td[ntd++] = q;
      const int layer = nb_low(nb) ? 1 : nc;
      for (int b = 1; b <= (ND == 3 ? nc : 1); ++b)
        for (int a = 1; a <= nc; ++a) {
          int ijk[3] = {1, 1, 1};
          ijk[d] = layer;
          ijk[td[0]] = a;
          if (ND == 3) ijk[td[1]] = b;
          double* s = &full[(size_t)NCF * g.lin(ijk[0], ijk[1], ND == 3 ? ijk[2] : 0)];
          double* b2r = &bc_to_rhs[(a - 1) + nc * (b - 1) + nface * (nb - 1)];
          if (bc_type == af_bc_dirichlet) {
            s[0] = s[0] - s[nb];
            *b2r = -2 * s[nb];
            s[nb] = 0.0;
          } else if (bc_type == af_bc_neumann) {
            s[0] = s[0] + s[nb];
            *b2r = -(s[nb] * box.dr[d]) * nb_high_pm(nb);
            s[nb] = 0.0;
          } else {
            std::fprintf(stderr, "mg_box_lpl_stencil: unsupported boundary condition\n");
            return 1;
          }
        }
    }
    if (!box.op.f.empty())
      for (int m = 0; m < g.ncell(); ++m) t.cs_lsf_fac[(size_t)g.ncell() * ib + m] = box.op.f[m];
    // global rows: ilo = (ix - 1) * nc + 1 (:185-186)
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          int gi[3] = {(box.ix[0] - 1) * nc + i - 1, (box.ix[1] - 1) * nc + j - 1,
                       ND == 3 ? (box.ix[2] - 1) * nc + k - 1 : 0};
          int r = gi[0] + nx[0] * (gi[1] + nx[1] * gi[2]);
          const double* s = &full[(size_t)NCF * g.lin(i, j, k)];
          ab[(size_t)bw + (size_t)ldab * r] += s[0];
          for (int m = 0; m < 2 * ND; ++m) {
            if (s[m + 1] == 0.0) continue;
            int d = m >> 1, sgn = (m & 1) ? 1 : -1;
            int q = gi[d] + sgn;
            int c = r + sgn * gstride[d];
            if (q < 0 || q >= nx[d]) {
              if (!any_periodic) {
                std::fprintf(stderr, "coarse matrix: coupling outside the grid\n");
                return 2;
              }
              c = r - sgn * (nx[d] - 1) * gstride[d];  // periodic wrap (Hypre: HYPRE_StructGridSetPeriodic)
            }
            ab[(size_t)(bw + r - c) + (size_t)ldab * c] += s[m + 1];
          }
        }
  }
  // banded LU without pivoting (the matrix is a diagonally dominant M-matrix up to sign)
  for (int c = 0; c < n; ++c) {
    double piv = ab[(size_t)bw + (size_t)ldab * c];
    if (piv == 0.0 || !std::isfinite(piv)) return 3;
    int rmax = std::min(n - 1, c + bw);
    for (int r = c + 1; r <= rmax; ++r) ab[(size_t)(bw + r - c) + (size_t)ldab * c] /= piv;
    int cmax = std::min(n - 1, c + bw);
#pragma omp parallel for if (bw > 64)
    for (int c2 = c + 1; c2 <= cmax; ++c2) {
      double u = ab[(size_t)(bw + c - c2) + (size_t)ldab * c2];
      if (u == 0.0) continue;
      for (int r = c + 1; r <= rmax; ++r)
        ab[(size_t)(bw + r - c2) + (size_t)ldab * c2] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * u;
    }
  }
  // a singular system (all-Neumann, no Helmholtz term) shows up as a tiny last pivot
  double last = std::fabs(ab[(size_t)bw + (size_t)ldab * (n - 1)]);
  double first = std::fabs(ab[(size_t)bw]);
  if (last < 1e-10 * first) return 4;
  t.cs_ready = true;
  return 0;
}

// solve_coarse_grid: coarse_solver_set_rhs_phi + solve + coarse_solver_get_phi + af_gc_lvl(1)
// (afivo/src/m_af_multigrid.f90:266-291, afivo/src/m_coarse_solver.f90:286-358)
template <int ND>
void solve_coarse_grid(Tree& t) {
  const int nc = t.nc;
  G<ND> g(nc);
  const auto& ids1 = t.ids[1];
  const int n = t.cs_n, bw = t.cs_bw, ldab = 2 * bw + 1;
  const int* nx = t.cs_nx;
  const int nface = (ND == 3) ? nc * nc : nc;
  std::vector<double> x(n, 0.0);
  for (int ib = 0; ib < (int)ids1.size(); ++ib) {
    Box& box = t.boxes[ids1[ib]];
    std::vector<double> tmp(g.ncell());
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) tmp[g.lin(i, j, k)] = box.cc[I_RHS][g.at(i, j, k)];
    for (int nb = 1; nb <= 2 * ND; ++nb) {
      if (box.neighbors[nb - 1] >= af_no_box) continue;
      const int d = nb_dim(nb);
      int td[2], ntd = 0;
      for (int q = 0; q < ND; ++q)
        if (q != d) td[ntd++] = q;
      const int layer = nb_low(nb) ? 1 : nc;
      const double* b2r = &t.cs_bc_to_rhs[(size_t)nface * (2 * ND * ib + (nb - 1))];
      for (int b = 1; b <= (ND == 3 ? nc : 1); ++b)
        for (int a = 1; a <= nc; ++a) {
          int ijk[3] = {1, 1, 1};
          ijk[d] = layer;
          ijk[td[0]] = a;
          if (ND == 3) ijk[td[1]] = b;
          int m = (a - 1) + nc * (b - 1);
          int L = g.lin(ijk[0], ijk[1], ND == 3 ? ijk[2] : 0);
          tmp[L] = tmp[L] + b2r[m] * box.bc_val[nb - 1][m];
        }
    }
    bool any_lsf = false;
    for (auto& b2 : t.boxes)
      if (!b2.lsf_dd.empty()) { any_lsf = true; break; }
    if (any_lsf)
      for (int m = 0; m < g.ncell(); ++m)
        tmp[m] = tmp[m] + t.cs_lsf_fac[(size_t)g.ncell() * ib + m] * lsf_boundary_value_at(t, box, m);
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          int gi[3] = {(box.ix[0] - 1) * nc + i - 1, (box.ix[1] - 1) * nc + j - 1,
                       ND == 3 ? (box.ix[2] - 1) * nc + k - 1 : 0};
          x[gi[0] + nx[0] * (gi[1] + nx[1] * gi[2])] = tmp[g.lin(i, j, k)];
        }
  }
  const std::vector<double>& ab = t.cs_lu;
  for (int c = 0; c < n; ++c) {  // L y = b (unit lower)
    double xc = x[c];
    if (xc == 0.0) continue;
    int rmax = std::min(n - 1, c + bw);
    for (int r = c + 1; r <= rmax; ++r) x[r] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * xc;
  }
  for (int c = n - 1; c >= 0; --c) {  // U x = y
    x[c] /= ab[(size_t)bw + (size_t)ldab * c];
    double xc = x[c];
    int rmin = std::max(0, c - bw);
    for (int r = rmin; r < c; ++r) x[r] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * xc;
  }
  for (int ib = 0; ib < (int)ids1.size(); ++ib) {
    Box& box = t.boxes[ids1[ib]];
    for (int k = g.klo(); k <= g.khi(); ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          int gi[3] = {(box.ix[0] - 1) * nc + i - 1, (box.ix[1] - 1) * nc + j - 1,
                       ND == 3 ? (box.ix[2] - 1) * nc + k - 1 : 0};
          box.cc[I_PHI][g.at(i, j, k)] = x[gi[0] + nx[0] * (gi[1] + nx[1] * gi[2])];
        }
  }
  af_gc_lvl<ND>(t, 1, I_PHI);
}

// ---------------------------------------------------------------------------------------------
// Multigrid driver (afivo/src/m_af_multigrid.f90:137-264, 624-810)
// ---------------------------------------------------------------------------------------------

// mg_auto_op / residual_box (:801-810, :906-912)
template <int ND>
void residual_box(Tree& t, Box& box) {
  const int nc = t.nc;
  G<ND> g(nc);
  stencil_apply_357<ND>(box, box.op, I_PHI, I_TMP, nc);
  for (int k = g.klo(); k <= g.khi(); ++k)
    for (int j = 1; j <= nc; ++j)
      for (int i = 1; i <= nc; ++i) {
        int n = g.at(i, j, k);
        box.cc[I_TMP][n] = box.cc[I_RHS][n] - box.cc[I_TMP][n];
      }
}

// gsrb_boxes (:648-687)
template <int ND>
void gsrb_boxes(Tree& t, int lvl, int type_cycle) {
  const int n_cycle = (type_cycle == mg_cycle_down) ? t.n_cycle_down : t.n_cycle_up;
  const auto& ids = t.ids[lvl];
#pragma omp parallel
  for (int n = 1; n <= 2 * n_cycle; ++n) {
#pragma omp for
    for (int i = 0; i < (int)ids.size(); ++i) {
      Box& box = t.boxes[ids[i]];
      stencil_gsrb_357<ND>(box, box.op, n, I_PHI, I_RHS, t.nc);  // mg_auto_gsrb (:813-820)
    }
    bool use_corners = t.use_corners || (type_cycle != mg_cycle_down && n == 2 * n_cycle);
#pragma omp for
    for (int i = 0; i < (int)ids.size(); ++i) af_gc_box<ND>(t, ids[i], I_PHI, use_corners);
  }
}

// update_coarse (:691-738); with_tmp = false gives set_coarse_phi_rhs (:742-776)
template <int ND>
void update_coarse(Tree& t, int lvl, bool with_tmp) {
  const int nc = t.nc;
  G<ND> g(nc);
  const auto& ids = t.ids[lvl];
  if (!with_tmp && lvl == t.highest_lvl) af_gc_lvl<ND>(t, lvl, I_PHI);  // :750-752
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) {
    Box& box = t.boxes[ids[i]];
    Box& box_p = t.boxes[box.parent];
    std::vector<double> tmp;
    if (with_tmp) tmp = box.cc[I_TMP];
    residual_box<ND>(t, box);
    mg_box_rstr_lpl<ND>(t, box, box_p, I_TMP);
    mg_box_rstr_lpl<ND>(t, box, box_p, I_PHI);
    if (with_tmp) {  // restore interior (:715)
      for (int k = g.klo(); k <= g.khi(); ++k)
        for (int j = 1; j <= nc; ++j)
          for (int ii = 1; ii <= nc; ++ii) box.cc[I_TMP][g.at(ii, j, k)] = tmp[g.at(ii, j, k)];
    }
  }
  af_gc_lvl<ND>(t, lvl - 1, I_PHI);
  const auto& par = t.parents[lvl - 1];
#pragma omp parallel for
  for (int i = 0; i < (int)par.size(); ++i) {
    Box& box = t.boxes[par[i]];
    stencil_apply_357<ND>(box, box.op, I_PHI, I_RHS, nc);                                   // rhs = L phi
    for (int n = 0; n < g.size(); ++n) box.cc[I_RHS][n] = box.cc[I_RHS][n] + box.cc[I_TMP][n];  // af_box_add_cc
    if (with_tmp) box.cc[I_TMP] = box.cc[I_PHI];                                           // af_box_copy_cc
  }
}

// init_phi_rhs (:779-799); note lvl runs highest_lvl..2, level 1 phi is not cleared
template <int ND>
void init_phi_rhs(Tree& t) {
  for (int lvl = t.highest_lvl; lvl >= 2; --lvl) {
    const auto& ids = t.ids[lvl];
#pragma omp parallel for
    for (int i = 0; i < (int)ids.size(); ++i) {
      Box& box = t.boxes[ids[i]];
      std::fill(box.cc[I_PHI].begin(), box.cc[I_PHI].end(), 0.0);
      mg_box_rstr_lpl<ND>(t, box, t.boxes[box.parent], I_RHS);
    }
  }
}

// correct_children (:624-646), mg_auto_corr (:943-950)
template <int ND>
void correct_children(Tree& t, int lvl_parents) {
  const auto& ids = t.parents[lvl_parents];
  G<ND> g(t.nc);
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) {
    Box& box = t.boxes[ids[i]];
    for (int n = 0; n < g.size(); ++n) box.cc[I_TMP][n] = box.cc[I_PHI][n] - box.cc[I_TMP][n];
    for (int ic = 0; ic < (1 << ND); ++ic) {
      int c_id = box.children[ic];
      if (c_id == af_no_box) continue;
      Box& bc = t.boxes[c_id];
      stencil_prolong<ND>(box, bc, bc.prolong, I_TMP, I_PHI, t.nc);
    }
  }
}

// af_tree_sum_cc (afivo/src/m_af_utils.f90:966-1027); deterministic box order here
template <int ND>
double af_tree_sum_cc(Tree& t, int iv) {
  const int nc = t.nc;
  G<ND> g(nc);
  double my_sum = 0;
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) {
    double fac = 1.0;
    for (int d = 0; d < ND; ++d) fac *= t.dr_base[d] * std::pow(0.5, lvl - 1);
    for (int id : t.leaves[lvl]) {
      const Box& box = t.boxes[id];
      double tmp = 0.0;
      if (ND == 2 && t.coord_t == af_cyl) {
        for (int j = 1; j <= nc; ++j)
          for (int i = 1; i <= nc; ++i) tmp = tmp + box.cc[iv][g.at(i, j, 0)] * cyl_radius_cc(box, i);
        tmp = tmp * (2 * std::acos(-1.0));
      } else {
        for (int k = g.klo(); k <= g.khi(); ++k)
          for (int j = 1; j <= nc; ++j)
            for (int i = 1; i <= nc; ++i) tmp = tmp + box.cc[iv][g.at(i, j, k)];
      }
      my_sum = my_sum + fac * tmp;
    }
  }
  return my_sum;
}

// af_total_volume (afivo/src/m_af_types.f90:805-825)
template <int ND>
double af_total_volume(Tree& t) {
  double box_len[3];
  for (int d = 0; d < ND; ++d) box_len[d] = t.nc * t.dr_base[d];
  if (ND == 2 && t.coord_t == af_cyl) {
    double vol = 0.0;
    const double pi = std::acos(-1.0);
    for (int id : t.ids[1]) {
      double r0 = t.boxes[id].r_min[0];
      double r1 = r0 + box_len[0];
      vol = vol + pi * (r1 * r1 - r0 * r0) * box_len[1];
    }
    return vol;
  }
  double v = (double)t.ids[1].size();
  for (int d = 0; d < ND; ++d) v *= box_len[d];
  return v;
}

// af_tree_maxabs_cc (afivo/src/m_af_utils.f90:773-785): leaves, interior cells
template <int ND>
double af_tree_maxabs_cc(Tree& t, int iv) {
  const int nc = t.nc;
  G<ND> g(nc);
  double m = 0.0;
  for (int lvl = 1; lvl <= t.highest_lvl; ++lvl) {
    const auto& lv = t.leaves[lvl];
#pragma omp parallel for reduction(max : m)
    for (int q = 0; q < (int)lv.size(); ++q) {
      const Box& box = t.boxes[lv[q]];
      for (int k = g.klo(); k <= g.khi(); ++k)
        for (int j = 1; j <= nc; ++j)
          for (int i = 1; i <= nc; ++i) m = std::max(m, std::fabs(box.cc[iv][g.at(i, j, k)]));
    }
  }
  return m;
}

// mg_fas_vcycle (:185-264)
template <int ND>
void mg_fas_vcycle(Tree& t, bool set_residual, int highest_lvl, bool standalone) {
  if (standalone) mg_set_operators_tree<ND>(t);  // mg_use
  int max_lvl = (highest_lvl > 0) ? highest_lvl : t.highest_lvl;
  for (int lvl = max_lvl; lvl >= 2; --lvl) {
    gsrb_boxes<ND>(t, lvl, mg_cycle_down);
    update_coarse<ND>(t, lvl, true);
  }
  solve_coarse_grid<ND>(t);
  for (int lvl = 2; lvl <= max_lvl; ++lvl) {
    correct_children<ND>(t, lvl - 1);
    af_gc_lvl<ND>(t, lvl, I_PHI);
    gsrb_boxes<ND>(t, lvl, mg_cycle_up);
  }
  if (set_residual) {
    for (int lvl = 1; lvl <= max_lvl; ++lvl) {
      const auto& ids = t.ids[lvl];
#pragma omp parallel for
      for (int i = 0; i < (int)ids.size(); ++i) residual_box<ND>(t, t.boxes[ids[i]]);
    }
  }
  if (t.subtract_mean) {
    double sum_phi = af_tree_sum_cc<ND>(t, I_PHI);
    double mean_phi = sum_phi / af_total_volume<ND>(t);
    for (int lvl = 1; lvl <= max_lvl; ++lvl) {
      const auto& ids = t.ids[lvl];
#pragma omp parallel for
      for (int i = 0; i < (int)ids.size(); ++i)
        for (double& x : t.boxes[ids[i]].cc[I_PHI]) x = x - mean_phi;
    }
  }
}

// mg_fas_fmg (:137-180)
template <int ND>
void mg_fas_fmg(Tree& t, bool set_residual, bool have_guess) {
  mg_set_operators_tree<ND>(t);  // mg_use
  if (have_guess) {
    for (int lvl = t.highest_lvl; lvl >= 2; --lvl) update_coarse<ND>(t, lvl, false);  // set_coarse_phi_rhs
  } else {
    init_phi_rhs<ND>(t);
  }
  for (int id : t.ids[1]) t.boxes[id].cc[I_TMP] = t.boxes[id].cc[I_PHI];
  mg_fas_vcycle<ND>(t, set_residual && 1 == t.highest_lvl, 1, false);
  for (int lvl = 2; lvl <= t.highest_lvl; ++lvl) {
    const auto& ids = t.ids[lvl];
#pragma omp parallel for
    for (int i = 0; i < (int)ids.size(); ++i) t.boxes[ids[i]].cc[I_TMP] = t.boxes[ids[i]].cc[I_PHI];
    correct_children<ND>(t, lvl - 1);
    af_gc_lvl<ND>(t, lvl, I_PHI);
    mg_fas_vcycle<ND>(t, set_residual && lvl == t.highest_lvl, lvl, false);
  }
}

#define DISPATCH(t, call)           \
  do {                              \
    if ((t)->ndim == 2) { constexpr int ND = 2; call; } \
    else { constexpr int ND = 3; call; }                \
  } while (0)

}  // namespace

// =============================================================================================
// C interface for ctypes (tests / bench cpu_baseline only)
// =============================================================================================
extern "C" {

void* orc_create(int ndim, int nc, int coord_t, int highest_lvl, int highest_id, const int* lvl_counts,
                 const int* lvl_ids_concat, const int* lvl, const int* ix, const int* parent, const int* children,
                 const int* neighbors, const int* neighbor_mat, const double* r_min, const double* dr,
                 const int* coarse_grid_size, const double* dr_base, const double* r_base, int with_eps) {
  Tree* t = new Tree();
  t->ndim = ndim;
  t->nc = nc;
  t->coord_t = coord_t;
  t->highest_lvl = highest_lvl;
  t->highest_id = highest_id;
  t->has_eps = with_eps != 0;
  for (int d = 0; d < ndim; ++d) {
    t->coarse_grid_size[d] = coarse_grid_size[d];
    t->dr_base[d] = dr_base[d];
    t->r_base[d] = r_base[d];
  }
  const int nch = 1 << ndim, nnb = 2 * ndim;
  int nm = 1;
  for (int d = 0; d < ndim; ++d) nm *= 3;
  const int n2 = nc + 2;
  const int bl = (ndim == 3) ? n2 * n2 * n2 : n2 * n2;
  t->boxes.resize(highest_id + 1);
  for (int id = 1; id <= highest_id; ++id) {
    Box& b = t->boxes[id];
    b.lvl = lvl[id];
    b.parent = parent[id];
    for (int d = 0; d < ndim; ++d) {
      b.ix[d] = ix[id * ndim + d];
      b.r_min[d] = r_min[id * ndim + d];
      b.dr[d] = dr[id * ndim + d];
    }
    for (int c = 0; c < nch; ++c) b.children[c] = children[id * nch + c];
    for (int c = 0; c < nnb; ++c) b.neighbors[c] = neighbors[id * nnb + c];
    for (int c = 0; c < nm; ++c) b.neighbor_mat[c] = neighbor_mat[id * nm + c];
    for (int v = 0; v < N_VAR; ++v)
      if (v != I_EPS || with_eps) b.cc[v].assign(bl, v == I_EPS ? 1.0 : 0.0);
  }
  t->ids.resize(highest_lvl + 2);
  t->leaves.resize(highest_lvl + 2);
  t->parents.resize(highest_lvl + 2);
  int p = 0;
  for (int l = 1; l <= highest_lvl; ++l) {
    for (int q = 0; q < lvl_counts[l - 1]; ++q) {
      int id = lvl_ids_concat[p++];
      t->ids[l].push_back(id);
      if (af_has_children(t->boxes[id])) t->parents[l].push_back(id);  // set_leaves_parents (m_af_core.f90:504-535)
      else t->leaves[l].push_back(id);
    }
  }
  return t;
}

void orc_destroy(void* h) { delete (Tree*)h; }

void orc_set_opts(void* h, int n_cycle_down, int n_cycle_up, int use_corners, int subtract_mean,
                  double helmholtz_lambda, double lsf_boundary_value, int operator_mask, int prolongation_type) {
  Tree* t = (Tree*)h;
  t->n_cycle_down = n_cycle_down;
  t->n_cycle_up = n_cycle_up;
  t->use_corners = use_corners != 0;
  t->subtract_mean = subtract_mean != 0;
  t->helmholtz_lambda = helmholtz_lambda;
  t->lsf_boundary_value = lsf_boundary_value;
  t->operator_mask = operator_mask;
  t->prolongation_type = prolongation_type;
}

// data: n x (nc+2)^D doubles
void orc_set_cc(void* h, int var, int n, const int* ids, const double* data) {
  Tree* t = (Tree*)h;
  for (int q = 0; q < n; ++q) {
    auto& v = t->boxes[ids[q]].cc[var];
    std::memcpy(v.data(), data + (size_t)q * v.size(), v.size() * sizeof(double));
  }
}
void orc_get_cc(void* h, int var, int n, const int* ids, double* data) {
  Tree* t = (Tree*)h;
  for (int q = 0; q < n; ++q) {
    auto& v = t->boxes[ids[q]].cc[var];
    std::memcpy(data + (size_t)q * v.size(), v.data(), v.size() * sizeof(double));
  }
}

// boundary conditions as data: for n faces (box id, nb 1..2D) the type and nc^(D-1) values
void orc_set_bc(void* h, int n, const int* ids, const int* nbs, const int* types, const double* vals) {
  Tree* t = (Tree*)h;
  const int nface = (t->ndim == 3) ? t->nc * t->nc : t->nc;
  for (int q = 0; q < n; ++q) {
    Box& b = t->boxes[ids[q]];
    b.bc_type[nbs[q] - 1] = types[q];
    b.bc_val[nbs[q] - 1].assign(vals + (size_t)q * nface, vals + (size_t)(q + 1) * nface);
  }
}

// level-set distances all_distances(2*ND, IJK) for n boxes (others: no boundary in box)
void orc_set_lsf_distances(void* h, int n, const int* ids, const double* dd) {
  Tree* t = (Tree*)h;
  const int nc = t->nc;
  const size_t per = (size_t)2 * t->ndim * (t->ndim == 3 ? nc * nc * nc : nc * nc);
  for (int q = 0; q < n; ++q) t->boxes[ids[q]].lsf_dd.assign(dd + q * per, dd + (q + 1) * per);
}

// mg_init (afivo/src/m_af_multigrid.f90:43-109) / mg_update_operator_stencil (:1188-1214)
int orc_mg_init(void* h) {
  Tree* t = (Tree*)h;
  int rc = 0;
  DISPATCH(t, { mg_set_operators_tree<ND>(*t, true); rc = coarse_solver_initialize<ND>(*t); });
  return rc;
}

void orc_fas_vcycle(void* h, int set_residual, int highest_lvl, int standalone) {
  Tree* t = (Tree*)h;
  DISPATCH(t, mg_fas_vcycle<ND>(*t, set_residual != 0, highest_lvl, standalone != 0));
}
void orc_fas_fmg(void* h, int set_residual, int have_guess) {
  Tree* t = (Tree*)h;
  DISPATCH(t, mg_fas_fmg<ND>(*t, set_residual != 0, have_guess != 0));
}

// single operations, for per-kernel parity tests
void orc_box_gsrb_lvl(void* h, int lvl, int redblack) {  // one half-sweep on all boxes, no ghost fill
  Tree* t = (Tree*)h;
  const auto& ids = t->ids[lvl];
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) {
    Box& box = t->boxes[ids[i]];
    DISPATCH(t, stencil_gsrb_357<ND>(box, box.op, redblack, I_PHI, I_RHS, t->nc));
  }
}
void orc_gsrb_boxes(void* h, int lvl, int type_cycle) {
  Tree* t = (Tree*)h;
  DISPATCH(t, gsrb_boxes<ND>(*t, lvl, type_cycle));
}
void orc_gc_lvl(void* h, int lvl, int var, int corners) {
  Tree* t = (Tree*)h;
  DISPATCH(t, af_gc_lvl<ND>(*t, lvl, var, corners != 0));
}
// ---- field from potential ----
void orc_gc_tree(void* h, int var, int corners) {
  Tree* t = (Tree*)h;
  DISPATCH(t, af_gc_tree<ND>(*t, var, corners != 0));
}
void orc_compute_phi_gradient(void* h, double fac, int with_norm) {
  Tree* t = (Tree*)h;
  DISPATCH(t, mg_compute_phi_gradient<ND>(*t, fac, with_norm != 0));
}
void orc_compute_field_norm(void* h) {
  Tree* t = (Tree*)h;
  DISPATCH(t, mg_compute_field_norm<ND>(*t));
}
// fc of n boxes, ND * (nc+1)^ND doubles each (zeros for a box without fc yet)
void orc_get_fc(void* h, int n, const int* ids, double* data) {
  Tree* t = (Tree*)h;
  size_t per = (size_t)t->ndim;
  for (int d = 0; d < t->ndim; ++d) per *= (size_t)(t->nc + 1);
  for (int q = 0; q < n; ++q) {
    const auto& v = t->boxes[ids[q]].fc;
    if (v.size() == per) std::memcpy(data + q * per, v.data(), per * sizeof(double));
    else std::memset(data + q * per, 0, per * sizeof(double));
  }
}
void orc_set_fc(void* h, int n, const int* ids, const double* data) {
  Tree* t = (Tree*)h;
  size_t per = (size_t)t->ndim;
  for (int d = 0; d < t->ndim; ++d) per *= (size_t)(t->nc + 1);
  for (int q = 0; q < n; ++q) t->boxes[ids[q]].fc.assign(data + q * per, data + (q + 1) * per);
}
// mg%lsf_boundary_function evaluated at the cell centres of n boxes (nc^ND each); n = 0 clears
void orc_set_lsf_boundary_values(void* h, int n, const int* ids, const double* v) {
  Tree* t = (Tree*)h;
  const size_t per = (t->ndim == 3) ? (size_t)t->nc * t->nc * t->nc : (size_t)t->nc * t->nc;
  for (auto& b : t->boxes) b.lsf_bv.clear();
  for (int q = 0; q < n; ++q) t->boxes[ids[q]].lsf_bv.assign(v + q * per, v + (q + 1) * per);
}
// mg%lsf_use_custom_prolongation with the prolongation distances (ND+1 per cell) of n boxes
void orc_set_lsf_prolong_distances(void* h, int n, const int* ids, const double* dd) {
  Tree* t = (Tree*)h;
  const size_t per = (size_t)(t->ndim + 1) * (t->ndim == 3 ? t->nc * t->nc * t->nc : t->nc * t->nc);
  t->lsf_use_custom_prolongation = n > 0;
  for (int q = 0; q < n; ++q) t->boxes[ids[q]].lsf_pdd.assign(dd + q * per, dd + (q + 1) * per);
}
// cc(IJK, mg%i_lsf) on the interior (nc^ND) of n boxes
void orc_set_lsf_cc(void* h, int n, const int* ids, const double* v) {
  Tree* t = (Tree*)h;
  const size_t per = (t->ndim == 3) ? (size_t)t->nc * t->nc * t->nc : (size_t)t->nc * t->nc;
  for (int q = 0; q < n; ++q) t->boxes[ids[q]].lsf_cc.assign(v + q * per, v + (q + 1) * per);
}
// boundary conditions of the field-norm variable (default: af_bc_neumann_zero)
void orc_set_fld_bc(void* h, int n, const int* ids, const int* nbs, const int* types, const double* vals) {
  Tree* t = (Tree*)h;
  const int nface = (t->ndim == 3) ? t->nc * t->nc : t->nc;
  for (int q = 0; q < n; ++q) {
    Box& b = t->boxes[ids[q]];
    b.fld_bc_type[nbs[q] - 1] = types[q];
    b.fld_bc_val[nbs[q] - 1].assign(vals + (size_t)q * nface, vals + (size_t)(q + 1) * nface);
  }
}
void orc_update_coarse(void* h, int lvl, int with_tmp) {
  Tree* t = (Tree*)h;
  DISPATCH(t, update_coarse<ND>(*t, lvl, with_tmp != 0));
}
void orc_correct_children(void* h, int lvl_parents) {
  Tree* t = (Tree*)h;
  DISPATCH(t, correct_children<ND>(*t, lvl_parents));
}
void orc_residual_lvl(void* h, int lvl) {
  Tree* t = (Tree*)h;
  const auto& ids = t->ids[lvl];
#pragma omp parallel for
  for (int i = 0; i < (int)ids.size(); ++i) DISPATCH(t, residual_box<ND>(*t, t->boxes[ids[i]]));
}
void orc_solve_coarse_grid(void* h) {
  Tree* t = (Tree*)h;
  DISPATCH(t, solve_coarse_grid<ND>(*t));
}
void orc_init_phi_rhs(void* h) {
  Tree* t = (Tree*)h;
  DISPATCH(t, init_phi_rhs<ND>(*t));
}
double orc_tree_maxabs(void* h, int var) {
  Tree* t = (Tree*)h;
  double r = 0;
  DISPATCH(t, r = af_tree_maxabs_cc<ND>(*t, var));
  return r;
}
double orc_tree_sum(void* h, int var) {
  Tree* t = (Tree*)h;
  double r = 0;
  DISPATCH(t, r = af_tree_sum_cc<ND>(*t, var));
  return r;
}

// stencil read-back (to check the product-side stencil construction against the restatement)
// returns stype; copies up to cap doubles of c or v into out; *has_f = f present
int orc_get_op_stencil(void* h, int id, double* out, int cap, double* f_out, int* has_f, int* cyl) {
  Tree* t = (Tree*)h;
  const Stencil& st = t->boxes[id].op;
  const auto& src = (st.stype == stencil_constant) ? st.c : st.v;
  std::memcpy(out, src.data(), std::min((size_t)cap, src.size()) * sizeof(double));
  *has_f = !st.f.empty();
  if (*has_f && f_out) std::memcpy(f_out, st.f.data(), st.f.size() * sizeof(double));
  *cyl = st.cylindrical_gradient;
  return st.stype;
}
int orc_get_prolong_stencil(void* h, int id, double* out, int cap, int* shape) {
  Tree* t = (Tree*)h;
  const Stencil& st = t->boxes[id].prolong;
  const auto& src = (st.stype == stencil_constant) ? st.c : st.v;
  std::memcpy(out, src.data(), std::min((size_t)cap, src.size()) * sizeof(double));
  *shape = st.shape;
  return st.stype;
}
int orc_get_tag(void* h, int id) { return ((Tree*)h)->boxes[id].tag; }

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
}
