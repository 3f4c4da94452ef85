// Stencil construction on the device: what mg_set_operators_lvl (afivo/src/m_af_multigrid.f90:
// 1147-1185) stores in box%stencils, computed where the data lives -- from the resident permittivity (AFMG_EPS) and
// from the streamer code's built-in electrode shapes (src/m_field.f90:686-904), which are device-callable here instead
// of a host callback per point.  After a refinement nothing but the topology goes up: the 7 nc^3 operator coefficients
// and 4 nc^3 prolongation weights per dielectric / electrode box are never shipped.
//
//   k_dev_tags   mg_set_box_tag (:1100-1145): eps min / max over the whole record; with an electrode the root mask
//                (get_possible_lsf_root_mask :954-975) and the number of cells with an internal boundary
//   k_dev_build  per tagged box: store_lsf_distance_matrix (:977-1097), mg_box_lpld_stencil (:1493-1532),
//                mg_box_lsf_stencil (:1782-1854), mg_box_lpld_lsf_stencil (:1535-1623), mg_box_prolong_eps_stencil
//                (:1308-1388), af_stencil_try_constant (m_af_stencil.f90:1001-1028)
// Both call the same per-cell functions as the host builders (afmg_builders.inc: operator_cell, prolong_eps_cell,
// lsf_mask_cell, lsf_dd_cell) and write the reference's v(n, i, j, k) order, so the result can be compared with the
// host builders bit for bit (tests/test_gpu_builders_device.py); ingest_stencils converts it to the device planes.
#pragma once

namespace afmg {

struct BuildCtx {
  const double* eps;    // [nslots * BOX] permittivity incl. ghost cells (device record layout), or null
  const double* rmin;   // [nslots * 3] box%r_min
  const int* lvl;       // [nslots]
  const int* parent;    // [nslots]
  const int* coff;      // [nslots] child offset bits
  double dr_base[3];
  int has_el;
  afmg_electrode el;
  afmg_lsf_opts lo;
  int operator_mask;
  int prolong_auto;
};

// layout of the per-box scratch the build kernel fills: v | f | pv | dd, reference order
template <int NC>
struct BuildBlob {
  static constexpr int NCELL = NC * NC * NC;
  static constexpr size_t V = 0, F = (size_t)7 * NCELL, PV = F + NCELL, DD = PV + (size_t)4 * NCELL, STRIDE = DD + (size_t)6 * NCELL;
};

__device__ __forceinline__ void box_dr(const BuildCtx& bc, int lvl, double* dr) {
  double fac = 1.0;
  for (int l = 1; l < lvl; ++l) fac = fac * 0.5;  // box%dr = parent%dr * 0.5: exact
  for (int d = 0; d < 3; ++d) dr[d] = bc.dr_base[d] * fac;
}

template <int NC>
__global__ void __launch_bounds__(256) k_dev_tags(BuildCtx bc, int nslots, int* tag_out, int* nb_out) {
  using L = Lay3<NC>;
  const int slot = blockIdx.x;
  if (slot >= nslots) return;
  __shared__ double s_min[256], s_max[256];
  __shared__ int s_cnt[256];
  const int t = threadIdx.x;
  int tag = 0;
  if (bc.eps) {
    const double* e = bc.eps + (size_t)slot * L::BOX;
    double a = e[0], b = e[0];
    for (int q = t; q < L::BOX; q += 256) {
      a = fmin(a, e[q]);
      b = fmax(b, e[q]);
    }
    s_min[t] = a;
    s_max[t] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (t < o) {
        s_min[t] = fmin(s_min[t], s_min[t + o]);
        s_max[t] = fmax(s_max[t], s_max[t + o]);
      }
      __syncthreads();
    }
    a = s_min[0];
    b = s_max[0];
    if (b > a) tag += AFMG_TAG_VEPS_BOX;
    else if (fmax(fabs(a - 1), fabs(b - 1)) > 1e-8) tag += AFMG_TAG_CEPS_BOX;
    __syncthreads();
  }
  int cnt = 0;
  if (bc.has_el) {
    builders::Lsf Lf{nullptr, nullptr, bc.lo, 3, &bc.el};
    double dr[3];
    box_dr(bc, bc.lvl[slot], dr);
    const double* rmin = bc.rmin + (size_t)slot * 3;
    const double dmax = builders::norm2(dr, 3);
    const double min_dr = fmin(dr[0], fmin(dr[1], dr[2]));
    for (int c = t; c < NC * NC * NC; c += 256) {
      const int ijk[3] = {c % NC + 1, (c / NC) % NC + 1, c / (NC * NC) + 1};
      if (!builders::lsf_mask_cell(Lf, 3, rmin, dr, ijk, nullptr, dmax)) continue;
      double dd[6];
      cnt += builders::lsf_dd_cell(Lf, 3, rmin, dr, ijk, nullptr, min_dr, dd) ? 1 : 0;
    }
  }
  s_cnt[t] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_cnt[t] += s_cnt[t + o];
    __syncthreads();
  }
  if (t == 0) {
    if (s_cnt[0] > 0) tag += AFMG_TAG_LSF_BOX;
    tag_out[slot] = tag;
    nb_out[slot] = s_cnt[0];
  }
}

// meta[4 * q]: op_stype (0 implicit, 1 constant, 2 variable), has_f, prolongation stype (0 none, 1 constant, 2 variable), -
template <int NC>
__global__ void __launch_bounds__(256) k_dev_build(BuildCtx bc, const int* list, const int* tags, int n, double* blob, int* meta) {
  using L = Lay3<NC>;
  using B = BuildBlob<NC>;
  const int q = blockIdx.x;
  if (q >= n) return;
  const int slot = list[q], t = threadIdx.x;
  const int tag = tags[slot], masked = tag & bc.operator_mask;
  const bool lsf = masked & AFMG_TAG_LSF_BOX, eps = masked & (AFMG_TAG_VEPS_BOX | AFMG_TAG_CEPS_BOX);
  double* v = blob + (size_t)q * B::STRIDE + B::V;
  double* f = blob + (size_t)q * B::STRIDE + B::F;
  double* pv = blob + (size_t)q * B::STRIDE + B::PV;
  double* ddo = blob + (size_t)q * B::STRIDE + B::DD;
  double dr[3], dr2[3], idr2[6];
  box_dr(bc, bc.lvl[slot], dr);
  for (int d = 0; d < 3; ++d) {
    dr2[d] = dr[d] * dr[d];
    idr2[2 * d] = idr2[2 * d + 1] = 1 / (dr[d] * dr[d]);
  }
  const double* rmin = bc.rmin + (size_t)slot * 3;
  const double* e = bc.eps ? bc.eps + (size_t)slot * L::BOX : nullptr;
  builders::Lsf Lf{nullptr, nullptr, bc.lo, 3, &bc.el};
  const double dmax = builders::norm2(dr, 3);
  const double min_dr = fmin(dr[0], fmin(dr[1], dr[2]));
  // ---- distances (kept for the field at the electrode) and operator
  for (int c = t; c < NC * NC * NC; c += 256) {
    const int i = c % NC + 1, j = (c / NC) % NC + 1, k = c / (NC * NC) + 1;
    const int ijk[3] = {i, j, k};
    double dd[6] = {1, 1, 1, 1, 1, 1};
    if (bc.has_el && (tag & AFMG_TAG_LSF_BOX)) {
      if (builders::lsf_mask_cell(Lf, 3, rmin, dr, ijk, nullptr, dmax)) builders::lsf_dd_cell(Lf, 3, rmin, dr, ijk, nullptr, min_dr, dd);
      for (int m = 0; m < 6; ++m) ddo[(size_t)6 * c + m] = dd[m];
    }
    if (!lsf && !eps) continue;
    double a0 = 0.0, a_nb[6] = {0, 0, 0, 0, 0, 0};
    if (eps) {
      a0 = e[L::cell(i, j, k)];
      a_nb[0] = e[L::cell(i - 1, j, k)];
      a_nb[1] = e[L::cell(i + 1, j, k)];
      a_nb[2] = e[L::cell(i, j - 1, k)];
      a_nb[3] = e[L::cell(i, j + 1, k)];
      a_nb[4] = e[L::cell(i, j, k - 1)];
      a_nb[5] = e[L::cell(i, j, k + 1)];
    }
    double w[7], fc = 0.0;
    builders::operator_cell(3, AFMG_XYZ, lsf, eps, dr, dr2, idr2, 0.0, a0, a_nb, dd, w, &fc);
    for (int m = 0; m < 7; ++m) v[(size_t)7 * c + m] = w[m];
    if (lsf) f[c] = fc;
  }
  // ---- prolongation from the parent's permittivity (mg_prolong_auto, variable-eps boxes above level 1)
  const bool pvar = bc.prolong_auto && (masked & AFMG_TAG_VEPS_BOX) && bc.lvl[slot] > 1;
  if (pvar) {
    const double* ep = bc.eps + (size_t)bc.parent[slot] * L::BOX;
    const int cof = bc.coff[slot];
    const int ox = (cof & 1) * (NC / 2), oy = ((cof >> 1) & 1) * (NC / 2), oz = ((cof >> 2) & 1) * (NC / 2);
    for (int c = t; c < NC * NC * NC; c += 256) {
      const int i = c % NC + 1, j = (c / NC) % NC + 1, k = c / (NC * NC) + 1;
      const int i1 = ox + ((i + 1) >> 1), i2 = i1 + 1 - 2 * (i & 1);
      const int j1 = oy + ((j + 1) >> 1), j2 = j1 + 1 - 2 * (j & 1);
      const int k1 = oz + ((k + 1) >> 1), k2 = k1 + 1 - 2 * (k & 1);
      const double a0 = ep[L::cell(i1, j1, k1)];
      const double a[3] = {ep[L::cell(i2, j1, k1)], ep[L::cell(i1, j2, k1)], ep[L::cell(i1, j1, k2)]};
      builders::prolong_eps_cell(3, a0, a, pv + (size_t)4 * c);
    }
  }
  __syncthreads();
  // ---- af_stencil_try_constant (abs_tol = epsilon): the two permittivity builders end with it, and so does the
  // prolongation builder; mg_box_lsf_stencil keeps its variable stencil
  const double tol = 2.220446049250313e-16;
  int op_const = 1, p_const = 1;
  if (lsf || eps)
    for (int c = t; c < NC * NC * NC; c += 256)
      for (int m = 0; m < 7; ++m)
        if (fabs(v[(size_t)7 * c + m] - v[m]) > tol) op_const = 0;
  if (pvar)
    for (int c = t; c < NC * NC * NC; c += 256)
      for (int m = 0; m < 4; ++m)
        if (fabs(pv[(size_t)4 * c + m] - pv[m]) > tol) p_const = 0;
  op_const = __syncthreads_and(op_const);
  p_const = __syncthreads_and(p_const);
  if (t == 0) {
    meta[4 * q + 0] = (lsf || eps) ? ((eps && op_const) ? 1 : 2) : 0;
    meta[4 * q + 1] = lsf ? 1 : 0;
    meta[4 * q + 2] = pvar ? (p_const ? 1 : 2) : 0;
    meta[4 * q + 3] = 0;
  }
}

}  // namespace afmg
