// afmg.hpp -- C++ host-side mirror of afivo's multigrid interface on top of the C ABI (afmg.h).
//
// The reference's callers are compiled Fortran (src/m_field.f90, src/m_photoi_helmh.f90, afivo/examples/*.f90); this
// header gives a compiled-language caller the same vocabulary: `mg_t` with the option members of
// afivo/src/m_af_types.f90:572-665, `mg_init / mg_destroy / mg_fas_fmg / mg_fas_vcycle / mg_update_operator_stencil`
// with the argument lists of afivo/src/m_af_multigrid.f90:43, :111, :137, :185, :1188, the reductions
// `af_tree_maxabs_cc / af_tree_sum_cc` (m_af_utils.f90:773, :966) and the field routines
// `mg_compute_phi_gradient / mg_compute_field_norm` (:1857, :2002).  `error stop` becomes `afmg::error`.
// Header only; link with -lafmg.  tools/poisson_benchmark.cpp is written against it.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "afmg.h"

namespace afmg {

struct error : std::runtime_error {
  int code;
  error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// The parts of af_t / box_t the solver reads (m_af_types.f90:286-393) as flat arrays with reference conventions:
// 1-based ids (row 0 unused), af_no_box = 0, af_phys_boundary = -1.  Cell data lives on the device.
struct af_t {
  int ndim = 3, n_cell = 0, coord_t = AFMG_XYZ;
  int highest_lvl = 0, highest_id = 0;
  int coarse_grid_size[3] = {1, 1, 1};
  bool periodic[3] = {false, false, false};
  double r_base[3] = {0, 0, 0}, dr_base[3] = {0, 0, 0};
  std::vector<std::vector<int32_t>> lvl_ids;  // [highest_lvl + 1]: lvls(l)%ids
  std::vector<int32_t> lvl, ix, parent, children, neighbors, neighbor_mat;
  std::vector<double> r_min;                  // (n+1, ndim): box%r_min
  std::vector<double> dr;                     // (n+1, ndim): box%dr

  int num_children() const { return 1 << ndim; }
  int num_neighbors() const { return 2 * ndim; }
  bool has_children(int id) const { return children[(size_t)id * num_children()] != 0; }
  std::vector<int32_t> ids(bool leaves_only) const {
    std::vector<int32_t> out;
    for (int l = 1; l <= highest_lvl; ++l)
      for (int32_t id : lvl_ids[l])
        if (!leaves_only || !has_children(id)) out.push_back(id);
    return out;
  }
  // af_r_cc (m_af_types.f90:1035-1040): centre of cell (i, j, k), ghost indices allowed
  void r_cc(int id, const int* ijk, double* r) const {
    for (int d = 0; d < ndim; ++d) r[d] = r_min[(size_t)id * ndim + d] + (ijk[d] - 0.5) * dr[(size_t)id * ndim + d];
  }
  // af_get_face_coords (m_af_types.f90:1215-1253): coords(ndim, nc^(ndim-1)) of the cell-face centres on side nb,
  // the transverse dimensions in increasing order, first one fastest
  void face_coords(int id, int nb, std::vector<double>& coords) const {
    const int d = (nb - 1) / 2;
    const bool low = (nb % 2) == 1;
    size_t nface = 1;
    for (int q = 1; q < ndim; ++q) nface *= (size_t)n_cell;
    coords.assign(nface * ndim, 0.0);
    int td[2] = {0, 0}, ntd = 0;
    for (int q = 0; q < ndim; ++q)
      if (q != d) td[ntd++] = q;
    const double* rm = &r_min[(size_t)id * ndim];
    const double* h = &dr[(size_t)id * ndim];
    for (size_t n = 0; n < nface; ++n) {
      double* c = &coords[n * ndim];
      c[d] = low ? rm[d] : rm[d] + n_cell * h[d];
      c[td[0]] = (rm[td[0]] + 0.5 * h[td[0]]) + (double)(n % n_cell) * h[td[0]];
      if (ndim == 3) c[td[1]] = (rm[td[1]] + 0.5 * h[td[1]]) + (double)(n / n_cell) * h[td[1]];
    }
  }
  size_t box_len() const {
    size_t n = 1;
    for (int d = 0; d < ndim; ++d) n *= (size_t)(n_cell + 2);
    return n;
  }
};

// subroutine sides_bc(box, nb, iv, coords, bc_val, bc_type) (m_af_types.f90:401-420) reduced to what the built-in
// conditions need: (box id, nb) -> type and one value for the whole face; use mg_t::set_bc for per-cell values
using sides_bc_t = std::function<void(int id, int nb, int& bc_type, double& bc_val)>;
inline void af_bc_dirichlet_zero(int, int, int& bc_type, double& bc_val) {  // m_af_ghostcell.f90:628-638
  bc_type = AFMG_BC_DIRICHLET;
  bc_val = 0.0;
}
inline void af_bc_neumann_zero(int, int, int& bc_type, double& bc_val) {  // m_af_ghostcell.f90:615-625
  bc_type = AFMG_BC_NEUMANN;
  bc_val = 0.0;
}

// the full callback: coords(ndim, nface) in, bc_val(nface) and bc_type out (e.g. sides_bc of
// afivo/examples/poisson_basic.f90:219-235: Dirichlet values from an analytic solution)
using sides_bc_coords_t =
    std::function<void(int id, int nb, const std::vector<double>& coords, std::vector<double>& bc_val, int& bc_type)>;

struct mg_t {
  // options a caller sets before mg_init (m_af_types.f90:572-665)
  int n_cycle_down = 2, n_cycle_up = 2;
  bool use_corners = false, subtract_mean = false;
  double helmholtz_lambda = 0.0, lsf_boundary_value = 0.0;
  int operator_mask = -1, prolongation_type = AFMG_PROLONG_AUTO;
  sides_bc_t sides_bc;
  sides_bc_coords_t sides_bc_coords;  // takes precedence over sides_bc when set
  int device = -1;
  int n_gpus = 0;  // > 1: this many GPUs of the node from ONE process (afmg_opts.n_gpus; 3D)
  bool initialized = false;
  afmg_handle* h = nullptr;

  mg_t() = default;
  mg_t(const mg_t&) = delete;
  mg_t& operator=(const mg_t&) = delete;
  ~mg_t() {
    if (h) afmg_destroy(h);
  }
  void check(int rc, const char* what) const {
    if (rc != AFMG_OK) throw error(rc, std::string(what) + ": " + afmg_last_error(h));
  }
  void need_init() const {
    if (!initialized) throw error(AFMG_ERR_STATE, "mg_t not initialized");  // the reference: error stop in mg_use
  }
  // box%cc(:, ..., iv) of the listed boxes, packed in list order (first index fastest)
  void set_cc(int var, const std::vector<int32_t>& ids, const double* packed) {
    need_init();
    check(afmg_upload(h, var, (int32_t)ids.size(), ids.data(), packed), "afmg_upload");
  }
  void set_cc_interior(int var, const std::vector<int32_t>& ids, const double* packed) {
    need_init();
    check(afmg_upload_interior(h, var, (int32_t)ids.size(), ids.data(), packed), "afmg_upload_interior");
  }
  void get_cc(int var, const std::vector<int32_t>& ids, double* packed) {
    need_init();
    check(afmg_download(h, var, (int32_t)ids.size(), ids.data(), packed), "afmg_download");
  }
  // interior cells only, nc^ndim doubles per box (what a caller that fills ghost cells itself needs)
  void get_cc_interior(int var, const std::vector<int32_t>& ids, double* packed) {
    need_init();
    check(afmg_download_interior(h, var, (int32_t)ids.size(), ids.data(), packed), "afmg_download_interior");
  }
  // wrapping sum / xor of the bit patterns of a variable over all boxes: equal for bit-identical solves
  std::pair<uint64_t, uint64_t> checksum(int var) {
    need_init();
    uint64_t a = 0, b = 0;
    check(afmg_checksum(h, var, &a, &b), "afmg_checksum");
    return {a, b};
  }
  // per-face boundary rows: ids, nb (1..2*ndim), types, values (nc^(ndim-1) each)
  void set_bc(const std::vector<int32_t>& ids, const std::vector<int32_t>& nb, const std::vector<int32_t>& types,
              const std::vector<double>& vals) {
    need_init();
    check(afmg_set_bc(h, (int32_t)ids.size(), ids.data(), nb.data(), types.data(), vals.data()), "afmg_set_bc");
  }
};

// Forward a (new) topology, e.g. after af_adjust_refinement, and evaluate mg%sides_bc on every physical face
inline void mg_set_tree(const af_t& tree, mg_t& mg) {
  mg.need_init();
  std::vector<int32_t> counts, concat;
  for (int l = 1; l <= tree.highest_lvl; ++l) {
    counts.push_back((int32_t)tree.lvl_ids[l].size());
    concat.insert(concat.end(), tree.lvl_ids[l].begin(), tree.lvl_ids[l].end());
  }
  afmg_tree td{};
  td.highest_lvl = tree.highest_lvl;
  td.highest_id = tree.highest_id;
  td.lvl_counts = counts.data();
  td.lvl_ids = concat.data();
  td.lvl = tree.lvl.data();
  td.ix = tree.ix.data();
  td.parent = tree.parent.data();
  td.children = tree.children.data();
  td.neighbors = tree.neighbors.data();
  td.neighbor_mat = tree.neighbor_mat.data();
  td.r_min = tree.r_min.empty() ? nullptr : tree.r_min.data();
  mg.check(afmg_set_tree(mg.h, &td), "afmg_set_tree");
  size_t nface = 1;
  for (int d = 1; d < tree.ndim; ++d) nface *= (size_t)tree.n_cell;
  std::vector<int32_t> bid, bnb, bty;
  std::vector<double> bval;
  for (int32_t id : concat)
    for (int nb = 1; nb <= tree.num_neighbors(); ++nb)
      if (tree.neighbors[(size_t)id * tree.num_neighbors() + nb - 1] == -1) {
        int ty = 0;
        if (mg.sides_bc_coords) {
          std::vector<double> coords, vals(nface, 0.0);
          tree.face_coords(id, nb, coords);
          mg.sides_bc_coords(id, nb, coords, vals, ty);
          bval.insert(bval.end(), vals.begin(), vals.end());
        } else {
          double v = 0.0;
          mg.sides_bc(id, nb, ty, v);
          bval.insert(bval.end(), nface, v);
        }
        bid.push_back(id);
        bnb.push_back(nb);
        bty.push_back(ty);
      }
  mg.set_bc(bid, bnb, bty, bval);
}

// mg_init (afivo/src/m_af_multigrid.f90:43-109)
inline void mg_init(const af_t& tree, mg_t& mg) {
  if (!mg.sides_bc && !mg.sides_bc_coords) throw error(AFMG_ERR_ARG, "mg_init: sides_bc not set");  // :50-51 stop
  afmg_opts o{};
  o.ndim = tree.ndim;
  o.n_cell = tree.n_cell;
  o.coord_t = tree.coord_t;
  o.n_cycle_down = mg.n_cycle_down;
  o.n_cycle_up = mg.n_cycle_up;
  o.use_corners = mg.use_corners;
  o.subtract_mean = mg.subtract_mean;
  o.prolongation_type = mg.prolongation_type;
  o.operator_mask = mg.operator_mask;
  o.device = mg.device;
  o.n_gpus = mg.n_gpus;
  o.helmholtz_lambda = mg.helmholtz_lambda;
  o.lsf_boundary_value = mg.lsf_boundary_value;
  for (int d = 0; d < 3; ++d) {
    o.coarse_grid_size[d] = d < tree.ndim ? tree.coarse_grid_size[d] : 1;
    o.periodic[d] = d < tree.ndim && tree.periodic[d];
    o.dr_base[d] = d < tree.ndim ? tree.dr_base[d] : 0.0;
    o.r_base[d] = d < tree.ndim ? tree.r_base[d] : 0.0;
  }
  const int rc = afmg_create(&mg.h, &o);
  if (rc != AFMG_OK) throw error(rc, std::string("afmg_create: ") + afmg_last_error(nullptr));
  mg.initialized = true;
  mg_set_tree(tree, mg);
}

// mg_destroy (:111-115)
inline void mg_destroy(mg_t& mg) {
  if (mg.h) afmg_destroy(mg.h);
  mg.h = nullptr;
  mg.initialized = false;
}

// mg_use (:118-126): the library keeps the operators of a handle current itself and every mg_t owns its handle, so only
// the reference's check remains
inline void mg_use(const af_t&, const mg_t& mg) { mg.need_init(); }

// mg_fas_fmg(tree, mg, set_residual, have_guess) (:137-180)
inline void mg_fas_fmg(const af_t&, mg_t& mg, bool set_residual, bool have_guess) {
  mg.need_init();
  mg.check(afmg_fas_fmg(mg.h, set_residual, have_guess), "afmg_fas_fmg");
}

// mg_fas_vcycle(tree, mg, set_residual, highest_lvl, standalone) (:185-264); highest_lvl = 0: tree%highest_lvl
inline void mg_fas_vcycle(const af_t&, mg_t& mg, bool set_residual, int highest_lvl = 0, bool standalone = true) {
  mg.need_init();
  mg.check(afmg_fas_vcycle(mg.h, set_residual, highest_lvl, standalone), "afmg_fas_vcycle");
}

// mg_update_operator_stencil (:1188-1214) after mg%helmholtz_lambda / mg%lsf_boundary_value changed
inline void mg_update_operator_stencil(const af_t&, mg_t& mg) {
  mg.need_init();
  mg.check(afmg_set_helmholtz_lambda(mg.h, mg.helmholtz_lambda), "afmg_set_helmholtz_lambda");
  mg.check(afmg_set_lsf_boundary_value(mg.h, mg.lsf_boundary_value), "afmg_set_lsf_boundary_value");
  mg.check(afmg_update_operator_stencil(mg.h), "afmg_update_operator_stencil");
}

// mg%lsf (m_af_types.f90:667-722, mg_func_lsf): level-set function of a point r(ndim)
using lsf_t = std::function<double(const double* r)>;

// one of the built-in electrode shapes (afmg_electrode, src/m_field.f90:254-362) as mg%lsf / mg%lsf_boundary_function;
// the struct must outlive the returned function
inline lsf_t electrode_lsf(afmg_electrode& e) {
  if (afmg_electrode_prepare(&e) != AFMG_OK) throw error(AFMG_ERR_ARG, "afmg_electrode_prepare: invalid electrode parameters");
  return [&e](const double* r) { return afmg_electrode_lsf(r, &e); };
}
inline lsf_t electrode_potential(afmg_electrode& e) {
  return [&e](const double* r) { return afmg_electrode_potential(r, &e); };
}

// The stencils of one tree as afmg_set_stencils takes them, plus the dense level-set distances
struct stencil_set_t {
  std::vector<afmg_stencil_desc> desc;
  std::vector<double> blob;
  std::vector<int32_t> lsf_ids;                   // boxes with mg_lsf_box
  std::vector<std::vector<double>> lsf_dd;        // per such box: 2*ndim * nc^ndim relative distances
  std::vector<std::vector<double>> lsf_cells;     // per such box: mg%lsf at the cell centres
};

// mg_set_operators_tree (m_af_multigrid.f90:1216-1225) for a tree with permittivity and / or a level-set function:
// mg_set_box_tag (:1100-1145) with store_lsf_distance_matrix (:977-1097), mg_store_operator_stencil (:823-859),
// mg_store_prolongation_stencil (:862-903), evaluated by the library's host-side builders (afmg_build_box_*).  No
// device is needed.  eps_cc: nullptr or (highest_id + 1) * (nc+2)^ndim doubles, box id major, ghost cells filled.
inline stencil_set_t mg_build_stencils(const af_t& tree, const mg_t& mg, const double* eps_cc, const lsf_t& lsf = nullptr,
                                       const afmg_lsf_opts* lsf_opts = nullptr, bool lsf_use_custom_prolongation = false) {
  stencil_set_t out;
  const int nd = tree.ndim, nc = tree.n_cell;
  size_t ncell = 1;
  for (int d = 0; d < nd; ++d) ncell *= (size_t)nc;
  const size_t blen = tree.box_len();
  auto trampoline = [](const double* r, void* user) -> double { return (*static_cast<const lsf_t*>(user))(r); };
  auto fail = [](int rc, const char* what) {
    if (rc != AFMG_OK) throw error(rc, what);
  };
  auto put = [&out](const double* a, size_t n) {
    const int64_t off = (int64_t)out.blob.size();
    out.blob.insert(out.blob.end(), a, a + n);
    return off;
  };
  std::vector<double> v((2 * nd + 1) * ncell), f(ncell), pv((nd + 1) * ncell), dd(2 * nd * ncell), pdd((nd + 1) * ncell);
  std::vector<uint8_t> mask(ncell);
  for (int l = 1; l <= tree.highest_lvl; ++l)
    for (int32_t id : tree.lvl_ids[l]) {
      const double* rmin = &tree.r_min[(size_t)id * nd];
      const double* dr = &tree.dr[(size_t)id * nd];
      const double* eps = eps_cc ? eps_cc + (size_t)id * blen : nullptr;
      int32_t n_boundary = 0;
      if (lsf)
        fail(afmg_build_box_lsf_distances(nd, nc, rmin, dr, trampoline, (void*)&lsf, lsf_opts, nullptr, mask.data(),
                                          dd.data(), &n_boundary), "afmg_build_box_lsf_distances");
      const int32_t tag = afmg_build_box_tag(nd, nc, eps, n_boundary > 0);
      if (tag < 0) throw error(tag, "afmg_build_box_tag");
      if (n_boundary > 0) {
        out.lsf_ids.push_back(id);
        out.lsf_dd.push_back(dd);
        std::vector<double> cells(ncell);
        for (size_t c = 0; c < ncell; ++c) {
          int ijk[3] = {(int)(c % nc) + 1, (int)((c / nc) % nc) + 1, (int)(c / ((size_t)nc * nc)) + 1};
          double r[3];
          tree.r_cc(id, ijk, r);
          cells[c] = lsf(r);
        }
        out.lsf_cells.push_back(std::move(cells));
      }
      if (tag == 0) continue;
      const int32_t masked = tag & mg.operator_mask;
      afmg_stencil_desc d{};
      d.box_id = id;
      d.tag = tag;
      d.f_offset = -1;
      if (masked != 0) {
        int32_t stype = 0, has_f = 0, cyl = 0;
        fail(afmg_build_box_operator(nd, nc, tree.coord_t, masked, dr, rmin, eps, dd.data(), v.data(), f.data(), &stype,
                                     &has_f, &cyl), "afmg_build_box_operator");
        d.op_stype = stype;
        d.cylindrical_gradient = cyl;
        d.op_offset = put(v.data(), stype == 1 ? (size_t)(2 * nd + 1) : v.size());
        if (has_f) d.f_offset = put(f.data(), ncell);
      }
      if (l > 1 && mg.prolongation_type == AFMG_PROLONG_AUTO) {
        const bool veps = masked & AFMG_TAG_VEPS_BOX;
        const bool vlsf = !veps && (masked & AFMG_TAG_LSF_BOX) && lsf_use_custom_prolongation;
        if (veps || vlsf) {
          const int32_t p = tree.parent[id];
          if (vlsf)
            fail(afmg_build_box_lsf_prolong_distances(nd, nc, rmin, dr, &tree.ix[(size_t)id * nd],
                                                      &tree.r_min[(size_t)p * nd], &tree.dr[(size_t)p * nd], trampoline,
                                                      (void*)&lsf, lsf_opts, mask.data(), pdd.data()),
                 "afmg_build_box_lsf_prolong_distances");
          int32_t pst = 0, psh = 0;
          fail(afmg_build_box_prolongation(nd, nc, masked, &tree.ix[(size_t)id * nd],
                                           veps ? eps_cc + (size_t)p * blen : nullptr, vlsf ? pdd.data() : nullptr,
                                           pv.data(), &pst, &psh), "afmg_build_box_prolongation");
          d.prolong_stype = pst;
          d.prolong_shape = psh;
          d.prolong_offset = put(pv.data(), pst == 1 ? (size_t)(nd + 1) : pv.size());
        }
      }
      out.desc.push_back(d);
    }
  if (lsf) {  // check_coarse_representation_lsf (m_af_multigrid.f90:2142-2161): error stop in the reference
    bool on_coarse = false;
    for (int32_t id : out.lsf_ids) on_coarse = on_coarse || tree.lvl[id] == 1;
    if (!on_coarse)
      throw error(AFMG_ERR_ARG, "level set function not resolved on coarse grid: no roots found on level 1, use a finer coarse grid");
  }
  return out;
}

// Build (mg_build_stencils) and ship: stencils, permittivity (AFMG_EPS) and level-set distances
inline stencil_set_t mg_set_operators_tree(const af_t& tree, mg_t& mg, const double* eps_cc, const lsf_t& lsf = nullptr,
                                           const afmg_lsf_opts* lsf_opts = nullptr,
                                           bool lsf_use_custom_prolongation = false) {
  mg.need_init();
  stencil_set_t st = mg_build_stencils(tree, mg, eps_cc, lsf, lsf_opts, lsf_use_custom_prolongation);
  const double zero = 0.0;
  mg.check(afmg_set_stencils(mg.h, (int32_t)st.desc.size(), st.desc.data(), st.blob.empty() ? &zero : st.blob.data(),
                             (int64_t)st.blob.size()), "afmg_set_stencils");
  const int nd = tree.ndim, nc = tree.n_cell;
  if (eps_cc) {
    const std::vector<int32_t> ids = tree.ids(false);
    std::vector<double> packed(ids.size() * tree.box_len());
    for (size_t n = 0; n < ids.size(); ++n)
      std::copy(eps_cc + (size_t)ids[n] * tree.box_len(), eps_cc + (size_t)(ids[n] + 1) * tree.box_len(),
                packed.begin() + n * tree.box_len());
    mg.set_cc(AFMG_EPS, ids, packed.data());
  }
  if (!st.lsf_ids.empty()) {  // the sparse form of the distance stencils (store_lsf_distance_matrix :1075-1094)
    std::vector<int32_t> n_entries, cell_ix;
    std::vector<double> dd, lv;
    for (size_t b = 0; b < st.lsf_ids.size(); ++b) {
      int32_t n = 0;
      const size_t ncell = st.lsf_cells[b].size();
      for (size_t c = 0; c < ncell; ++c) {
        const double* q = &st.lsf_dd[b][c * 2 * nd];
        bool any = false;
        for (int m = 0; m < 2 * nd; ++m) any = any || q[m] < 1.0;
        if (!any) continue;
        ++n;
        size_t r = c;
        for (int d = 0; d < nd; ++d) {
          cell_ix.push_back((int32_t)(r % nc) + 1);
          r /= nc;
        }
        dd.insert(dd.end(), q, q + 2 * nd);
        lv.push_back(st.lsf_cells[b][c]);
      }
      n_entries.push_back(n);
    }
    mg.check(afmg_set_lsf_distances(mg.h, (int32_t)st.lsf_ids.size(), st.lsf_ids.data(), n_entries.data(), cell_ix.data(),
                                    dd.data(), lv.data()), "afmg_set_lsf_distances");
  }
  return st;
}

// af_tree_maxabs_cc (m_af_utils.f90:773-785): max |cc| over the interior of the leaves
// field_set_rhs (src/m_field.f90:406-444) on the device: cc(:, i_rhs) = sum_n charges[n] * densities[n] on the listed
// leaves, in the reference's order.  densities[n] = packed records ((nc+2)^ndim doubles per box, list order) in host
// memory, or in device memory with on_device.
inline void field_set_rhs(const af_t&, mg_t& mg, const std::vector<int32_t>& leaf_ids, const std::vector<double>& charges,
                          const std::vector<const double*>& densities, bool on_device = false) {
  mg.need_init();
  if (charges.size() != densities.size()) throw error(AFMG_ERR_ARG, "field_set_rhs: one density array per charge");
  mg.check(afmg_field_set_rhs(mg.h, (int32_t)leaf_ids.size(), leaf_ids.data(), (int32_t)charges.size(), charges.data(),
                              densities.data(), on_device ? 1 : 0),
           "afmg_field_set_rhs");
}

// mg_set_operators_tree with the stencils built ON THE DEVICE from the resident permittivity (AFMG_EPS, if the caller
// uploaded one) and a built-in electrode shape: nothing but the topology travels after a refinement.  Single-GPU
// handles, 3D.
inline void mg_set_operators_tree_device(const af_t&, mg_t& mg, const afmg_electrode* electrode = nullptr,
                                         const afmg_lsf_opts* lsf_opts = nullptr) {
  mg.need_init();
  mg.check(afmg_build_stencils_device(mg.h, electrode, lsf_opts), "afmg_build_stencils_device");
}

inline double af_tree_maxabs_cc(const af_t&, mg_t& mg, int var) {
  mg.need_init();
  double v = 0.0;
  mg.check(afmg_max_abs(mg.h, var, &v), "afmg_max_abs");
  return v;
}

// af_tree_sum_cc (m_af_utils.f90:966-1027): volume-weighted sum over the leaves
inline double af_tree_sum_cc(const af_t&, mg_t& mg, int var) {
  mg.need_init();
  double v = 0.0;
  mg.check(afmg_tree_sum(mg.h, var, &v), "afmg_tree_sum");
  return v;
}

// mg_compute_phi_gradient(tree, mg, i_fc, fac, i_norm) (:1857-1898); results stay on the device
// (afmg_download_fc, get_cc(AFMG_FLD, ...))
inline void mg_compute_phi_gradient(const af_t&, mg_t& mg, double fac, bool with_norm = true) {
  mg.need_init();
  mg.check(afmg_compute_phi_gradient(mg.h, fac, with_norm), "afmg_compute_phi_gradient");
}

// mg_compute_field_norm (:2002-2020)
inline void mg_compute_field_norm(const af_t&, mg_t& mg) {
  mg.need_init();
  mg.check(afmg_compute_field_norm(mg.h), "afmg_compute_field_norm");
}

// af_gc_tree(tree, [iv], corners) (m_af_ghostcell.f90:25-46) for AFMG_PHI or AFMG_FLD
inline void af_gc_tree(const af_t&, mg_t& mg, int var, bool corners = true) {
  mg.need_init();
  mg.check(afmg_gc_tree(mg.h, var, corners), "afmg_gc_tree");
}

// field_from_potential without dielectric (src/m_field.f90:543-547): gradient with norm, then af_gc_tree of the norm
inline void field_from_potential(const af_t&, mg_t& mg, double fac = -1.0) {
  mg.need_init();
  mg.check(afmg_field_from_potential(mg.h, fac), "afmg_field_from_potential");
}

// box%fc(:, ..., i_fc) of the listed boxes: ndim * (nc+1)^ndim doubles per box, in the reference's element order
inline std::vector<double> get_fc(const af_t& tree, mg_t& mg, const std::vector<int32_t>& ids) {
  mg.need_init();
  size_t per = (size_t)tree.ndim;
  for (int d = 0; d < tree.ndim; ++d) per *= (size_t)(tree.n_cell + 1);
  std::vector<double> out(ids.size() * per);
  mg.check(afmg_download_fc(mg.h, (int32_t)ids.size(), ids.data(), out.data()), "afmg_download_fc");
  return out;
}

// mg%lsf_boundary_function as data (m_af_types.f90:628): its values at the nc^ndim cell centres of the listed boxes,
// e.g. from electrode_potential(); an empty list returns to the scalar mg%lsf_boundary_value
inline void mg_set_lsf_boundary_values(const af_t&, mg_t& mg, const std::vector<int32_t>& ids, const std::vector<double>& values) {
  mg.need_init();
  mg.check(afmg_set_lsf_boundary_values(mg.h, (int32_t)ids.size(), ids.data(), values.data()), "afmg_set_lsf_boundary_values");
}

// The FMG / V-cycle convergence loop of field_compute (src/m_field.f90:491-524) run next to the device
struct field_solve_result_t {
  std::vector<double> residuals;
  int n_fmg = 0, n_vcycles = 0;
};
inline field_solve_result_t field_solve(const af_t&, mg_t& mg, bool have_guess, double residual_threshold,
                                        double max_residual = 1e8, int max_initial_iterations = 100, int num_vcycles = 2) {
  mg.need_init();
  field_solve_result_t r;
  r.residuals.assign((size_t)max_initial_iterations + num_vcycles, 0.0);
  int32_t n_fmg = 0, n_vc = 0;
  mg.check(afmg_field_solve(mg.h, have_guess, residual_threshold, max_residual, max_initial_iterations, num_vcycles,
                            r.residuals.data(), &n_fmg, &n_vc), "afmg_field_solve");
  r.n_fmg = n_fmg;
  r.n_vcycles = n_vc;
  r.residuals.resize((size_t)n_fmg + n_vc);
  return r;
}

// photoi_helmh_compute (src/m_photoi_helmh.f90:162-204) on the device: mg_helm = the mg_t of every mode
// (helmholtz_lambda = lambdas(n)**2, boundary conditions photoi_helmh_bc), right-hand side uploaded to mg_helm[0]; the
// source is read with mg_helm[0]->get_cc(AFMG_PHOTO, ...).  Returns the FMG count of every mode.
inline std::vector<int32_t> photoi_helmh_compute(const af_t&, const std::vector<mg_t*>& mg_helm, const std::vector<double>& coeffs,
                                                 int max_fmg_cycles = 10, double max_rel_residual = 1.0e-2,
                                                 std::vector<double>* residuals = nullptr) {
  if (mg_helm.empty() || coeffs.size() != mg_helm.size()) throw error(AFMG_ERR_ARG, "photoi_helmh_compute: one coefficient per mode");
  std::vector<afmg_handle*> hs;
  for (mg_t* m : mg_helm) {
    m->need_init();
    hs.push_back(m->h);
  }
  std::vector<int32_t> n_cycles(hs.size(), 0);
  std::vector<double> res(hs.size(), 0.0);
  mg_helm[0]->check(afmg_helmholtz_compute(hs.data(), (int32_t)hs.size(), coeffs.data(), max_fmg_cycles, max_rel_residual,
                                           n_cycles.data(), res.data()), "afmg_helmholtz_compute");
  if (residuals) *residuals = res;
  return n_cycles;
}

// photoi_helmh%author parameter sets (src/m_photoi_helmh.f90:80-136): lambdas [1/m] and coeffs [1/m^2], scaled by the O2
// fraction and the pressure in bar as there.  mg_helm(n)%helmholtz_lambda = lambdas(n)^2, mg_prolong_linear (:146-155).
struct helmh_params_t {
  std::vector<double> lambdas, coeffs;
};
inline helmh_params_t photoi_helmh_parameters(const std::string& author = "Bourdon-3", double frac_O2 = 0.2, double gas_pressure = 1.0,
                                              double eta = 1.0) {
  if (frac_O2 <= 0.0) throw error(AFMG_ERR_ARG, "Photoionization: no oxygen present");
  helmh_params_t p;
  double scale = frac_O2 * gas_pressure;
  if (author == "Luque") {
    if (std::fabs(eta - 1.0) > 0) throw error(AFMG_ERR_ARG, "With Luque photoionization, photoi%eta should be 1.0");
    p.lambdas = {4425.38, 750.06};
    p.coeffs = {337557.38, 19972.14};
    scale = (frac_O2 / 0.2) * gas_pressure;
  } else if (author == "Bourdon-2") {
    p.lambdas = {7305.62, 44081.25};
    p.coeffs = {11814508.38, 998607256.0};
  } else if (author == "Bourdon-3") {
    p.lambdas = {4147.85, 10950.93, 66755.67};
    p.coeffs = {1117314.935, 28692377.5, 2748842283.0};
  } else {
    throw error(AFMG_ERR_ARG, "Unknown photoi_helmh_author: " + author);
  }
  for (double& l : p.lambdas) l *= scale;
  for (double& c : p.coeffs) c *= scale * scale;
  return p;
}

// The convergence threshold of field_compute (src/m_field.f90:467-480)
inline double field_residual_threshold(const af_t& tree, double max_rhs, double current_voltage, bool use_electrode = false,
                                       double multigrid_max_rel_residual = 1.0e-4) {
  const int nd = tree.ndim;
  const double conv_fac = use_electrode ? 1.0e-8 : 1.0e-10;
  const double domain_len = tree.coarse_grid_size[nd - 1] * tree.dr_base[nd - 1];
  double min_dr = tree.dr_base[0];
  for (int d = 1; d < nd; ++d) min_dr = std::fmin(min_dr, tree.dr_base[d]);
  min_dr *= std::pow(0.5, tree.highest_lvl - 1);  // af_min_dr
  return std::fmax(1.0e-6, std::fmax(max_rhs * multigrid_max_rel_residual, conv_fac * std::fabs(current_voltage) / (domain_len * min_dr)));
}

// af_init followed by af_adjust_refinement until nothing is added: a 2:1 balanced 2D or 3D tree in the reference's
// conventions -- level-1 ids i + (j-1) nx + (k-1) nx ny (m_af_core.f90:436-501), children appended parent by parent
// in af_child_dix order (:1187-1254), neighbours / neighbor_mat with af_phys_boundary = -1 outside a non-periodic
// domain (:595-661), 2:1 balance over faces from the finest level down (ensure_two_one_balance, :1016-1057).
// refine(lvl, ix, centre) is asked for every existing box of level lvl < max_lvl (the union over the cells of a box
// of the reference's per-cell flags); nullptr refines everything.
using refine_t = std::function<bool(int lvl, const int* ix, const double* centre)>;

inline af_t af_build_tree_nd(int ndim, int n_cell, const int* coarse_grid_size, int max_lvl, const refine_t& refine = nullptr,
                             const double* r_lo = nullptr, const double* r_hi = nullptr, const bool* periodic = nullptr,
                             int coord_t = AFMG_XYZ) {
  if (ndim != 2 && ndim != 3) throw error(AFMG_ERR_UNSUPPORTED, "af_build_tree_nd: ndim must be 2 or 3");
  af_t t;
  const int nd = ndim, nch = 1 << ndim, nnb = 2 * ndim, nmat = ndim == 3 ? 27 : 9;
  t.ndim = ndim;
  t.coord_t = coord_t;
  t.n_cell = n_cell;
  int nb1[3] = {1, 1, 1};
  for (int d = 0; d < nd; ++d) {
    t.coarse_grid_size[d] = coarse_grid_size[d];
    t.periodic[d] = periodic ? periodic[d] : false;
    t.r_base[d] = r_lo ? r_lo[d] : 0.0;
    t.dr_base[d] = ((r_hi ? r_hi[d] : 1.0) - t.r_base[d]) / coarse_grid_size[d];
    nb1[d] = coarse_grid_size[d] / n_cell;
  }
  auto nbl = [&](int l, int d) { return d < nd ? nb1[d] << (l - 1) : 1; };  // the unused third dimension has one box
  auto lin = [&](int l, int x, int y, int z) { return (size_t)x + (size_t)nbl(l, 0) * ((size_t)y + (size_t)nbl(l, 1) * z); };
  auto cells = [&](int l) { return (size_t)nbl(l, 0) * nbl(l, 1) * nbl(l, 2); };
  std::vector<std::vector<char>> exists(max_lvl + 2), refined(max_lvl + 2);
  exists[1].assign(cells(1), 1);
  auto upsample = [&](int l) {  // children of the refined boxes of level l
    exists[l + 1].assign(cells(l + 1), 0);
    for (int z = 0; z < nbl(l, 2); ++z)
      for (int y = 0; y < nbl(l, 1); ++y)
        for (int x = 0; x < nbl(l, 0); ++x)
          if (refined[l][lin(l, x, y, z)])
            for (int c = 0; c < nch; ++c) exists[l + 1][lin(l + 1, 2 * x + (c & 1), 2 * y + ((c >> 1) & 1), nd == 3 ? 2 * z + ((c >> 2) & 1) : 0)] = 1;
  };
  for (int l = 1; l < max_lvl; ++l) {
    refined[l].assign(cells(l), 0);
    for (int z = 0; z < nbl(l, 2); ++z)
      for (int y = 0; y < nbl(l, 1); ++y)
        for (int x = 0; x < nbl(l, 0); ++x) {
          if (!exists[l][lin(l, x, y, z)]) continue;
          bool r = true;
          if (refine) {
            const int ix[3] = {x + 1, y + 1, z + 1};
            double ctr[3] = {0, 0, 0};
            for (int d = 0; d < nd; ++d) ctr[d] = t.r_base[d] + (ix[d] - 0.5) * (t.dr_base[d] * std::pow(0.5, l - 1)) * n_cell;
            r = refine(l, ix, ctr);
          }
          refined[l][lin(l, x, y, z)] = r;
        }
    upsample(l);
  }
  refined[max_lvl].assign(cells(max_lvl), 0);
  for (int l = max_lvl - 1; l > 1; --l)  // 2:1 balance over faces
    for (int z = 0; z < nbl(l, 2); ++z)
      for (int y = 0; y < nbl(l, 1); ++y)
        for (int x = 0; x < nbl(l, 0); ++x) {
          if (!refined[l][lin(l, x, y, z)]) continue;
          for (int nb = -1; nb < nnb; ++nb) {  // the box itself and its face neighbours must exist
            int q[3] = {x, y, z};
            if (nb >= 0) {
              q[nb >> 1] += (nb & 1) ? 1 : -1;
              const int d = nb >> 1;
              if (q[d] < 0 || q[d] >= nbl(l, d)) {
                if (!t.periodic[d]) continue;
                q[d] = (q[d] + nbl(l, d)) % nbl(l, d);
              }
            }
            refined[l - 1][lin(l - 1, q[0] / 2, q[1] / 2, nd == 3 ? q[2] / 2 : 0)] = 1;
          }
        }
  for (int l = 1; l < max_lvl; ++l) upsample(l);
  int highest = 1;
  for (int l = 1; l <= max_lvl; ++l)
    for (char e : exists[l])
      if (e) {
        highest = l;
        break;
      }
  t.highest_lvl = highest;
  // enumerate level by level in the reference's list order
  struct P { int x, y, z; };
  std::vector<std::vector<P>> order(highest + 1);
  for (int z = 0; z < nbl(1, 2); ++z)
    for (int y = 0; y < nbl(1, 1); ++y)
      for (int x = 0; x < nbl(1, 0); ++x) order[1].push_back({x, y, z});
  for (int l = 1; l < highest; ++l)
    for (const P& p : order[l])
      if (refined[l][lin(l, p.x, p.y, p.z)])
        for (int c = 0; c < nch; ++c) order[l + 1].push_back({2 * p.x + (c & 1), 2 * p.y + ((c >> 1) & 1), nd == 3 ? 2 * p.z + ((c >> 2) & 1) : 0});
  size_t total = 0;
  for (int l = 1; l <= highest; ++l) total += order[l].size();
  t.highest_id = (int)total;
  const size_t N = total + 1;
  t.lvl.assign(N, 0);
  t.ix.assign(N * nd, 0);
  t.parent.assign(N, 0);
  t.children.assign(N * nch, 0);
  t.neighbors.assign(N * nnb, 0);
  t.neighbor_mat.assign(N * nmat, 0);
  t.r_min.assign(N * nd, 0.0);
  t.dr.assign(N * nd, 0.0);
  t.lvl_ids.assign(highest + 1, {});
  std::vector<std::vector<int32_t>> idgrid(highest + 1);
  int next = 1;
  for (int l = 1; l <= highest; ++l) {
    idgrid[l].assign(cells(l), 0);
    for (const P& p : order[l]) {
      const int id = next++;
      t.lvl_ids[l].push_back(id);
      t.lvl[id] = l;
      const int q[3] = {p.x, p.y, p.z};
      for (int d = 0; d < nd; ++d) {
        t.ix[(size_t)id * nd + d] = q[d] + 1;
        t.dr[(size_t)id * nd + d] = t.dr_base[d] * std::pow(0.5, l - 1);
      }
      idgrid[l][lin(l, p.x, p.y, p.z)] = id;
      if (l == 1) {
        for (int d = 0; d < nd; ++d) t.r_min[(size_t)id * nd + d] = t.r_base[d] + q[d] * t.dr_base[d] * n_cell;
      } else {
        const int pid = idgrid[l - 1][lin(l - 1, p.x / 2, p.y / 2, nd == 3 ? p.z / 2 : 0)];
        t.parent[id] = pid;
        const int c = (p.x & 1) | ((p.y & 1) << 1) | (nd == 3 ? (p.z & 1) << 2 : 0);
        t.children[(size_t)pid * nch + c] = id;
        for (int d = 0; d < nd; ++d)  // add_children: r_min = r_min_p + 0.5 * dr_p * dix * n_cell
          t.r_min[(size_t)id * nd + d] = t.r_min[(size_t)pid * nd + d] + 0.5 * t.dr[(size_t)pid * nd + d] * (q[d] & 1) * n_cell;
      }
    }
  }
  for (int l = 1; l <= highest; ++l)
    for (int32_t id : t.lvl_ids[l]) {
      const int32_t* q = &t.ix[(size_t)id * nd];
      for (int dz = (nd == 3 ? -1 : 0); dz <= (nd == 3 ? 1 : 0); ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            int p[3] = {q[0] - 1 + dx, q[1] - 1 + dy, nd == 3 ? q[2] - 1 + dz : 0};
            bool out = false;
            for (int d = 0; d < nd; ++d) {
              if (t.periodic[d]) p[d] = (p[d] + nbl(l, d)) % nbl(l, d);
              else out = out || p[d] < 0 || p[d] >= nbl(l, d);
            }
            t.neighbor_mat[(size_t)id * nmat + (dx + 1) + 3 * (dy + 1) + (nd == 3 ? 9 * (dz + 1) : 0)] = out ? -1 : idgrid[l][lin(l, p[0], p[1], p[2])];
          }
      for (int nb = 0; nb < nnb; ++nb) {
        int d[3] = {0, 0, 0};
        d[nb >> 1] = (nb & 1) ? 1 : -1;
        t.neighbors[(size_t)id * nnb + nb] = t.neighbor_mat[(size_t)id * nmat + (d[0] + 1) + 3 * (d[1] + 1) + (nd == 3 ? 9 * (d[2] + 1) : 0)];
      }
    }
  return t;
}

// the 3D form used by the examples in tools/
inline af_t af_build_tree(int n_cell, const int* coarse_grid_size, int max_lvl, const refine_t& refine = nullptr,
                          const double* r_lo = nullptr, const double* r_hi = nullptr, const bool* periodic = nullptr) {
  return af_build_tree_nd(3, n_cell, coarse_grid_size, max_lvl, refine, r_lo, r_hi, periodic, AFMG_XYZ);
}

// the tree of afivo/examples/poisson_benchmark.f90:72-90: unit cube, everything refined up to max_lvl
inline af_t af_init_fully_refined(int n_cell, int coarse_grid_size, int max_lvl) {
  const int cgs[3] = {coarse_grid_size, coarse_grid_size, coarse_grid_size};
  return af_build_tree(n_cell, cgs, max_lvl);
}

}  // namespace afmg
