// afivo `.dat` tree files, version 3, for compiled hosts: reader of the binary stream format of
// af_write_tree / af_read_tree (afivo/src/m_af_output.f90:41-192 writer, :197-373 reader, version constant :10), the
// C++ twin of afivo_streamer_b200/datfile.py, and mg_from_dat: a solver set up from a stored simulation state alone
// (topology, the boundary conditions stored in the boxes, phi / rhs / eps, the stored operator / prolongation stencils
// and the level-set distance stencils) -- restart or post-processing without Fortran and without Python.
//
// Format notes (gfortran, access='stream', no record markers): default integers and logicals are 4 bytes, reals 8
// bytes, names character(len=af_nlen=20) in arrays of af_max_num_vars = 1024 (m_af_types.f90:20, 72, 341-356).  NDIM is a
// compile-time constant of the writer and is NOT in the file: pass ndim or let the reader try 3, then 2 (only one of
// them consumes the file consistently).  The reference ships no .dat fixture: checked against the Python reader on
// files written by its writer (tests/test_cpp_host.py).
#pragma once

#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "afmg.hpp"

namespace afmg {

struct dat_stencil_t {  // stencil_t as stored (m_af_output.f90:134-182)
  int key = 0, shape = 0, stype = 0;
  bool cylindrical_gradient = false;
  std::vector<double> c, v, f, bc_correction;  // v: (cells, n_coeff) with the coefficient index fastest
  int n_coeff = 0;                             // coefficients per cell of v
  std::vector<int32_t> sparse_ix;              // (k, NDIM)
  std::vector<double> sparse_v;                // (k, m)
  int n_sparse = 0, m_sparse = 0;
};

struct dat_bc_t {  // boundary-condition storage of one box (af_init_box, m_af_core.f90:558-577)
  std::vector<int32_t> bc_index_to_nb, nb_to_bc_index, bc_type;  // bc_type: (n_bc, n_var_cell)
  std::vector<double> bc_val;                                    // (n_bc, n_var_cell, nc^(D-1))
};

struct dat_t {
  af_t tree;
  bool ready = false;
  int box_limit = 0;
  std::vector<std::string> cc_names, fc_names;
  std::vector<char> cc_write_binary;
  std::vector<int32_t> tag;                      // (n+1)
  std::map<int, std::vector<double>> cc;         // 1-based variable index -> (n+1) * (nc+2)^D
  std::map<int, dat_bc_t> bc;                    // box id ->
  std::map<int, std::vector<dat_stencil_t>> stencils;

  int var_index(const std::string& name) const {  // af_find_cc_variable: 1-based
    for (size_t i = 0; i < cc_names.size(); ++i)
      if (cc_names[i] == name) return (int)i + 1;
    throw error(AFMG_ERR_ARG, "no cell-centred variable named " + name);
  }
  const double* cc_of(int iv, int id) const {
    auto it = cc.find(iv);
    if (it == cc.end()) throw error(AFMG_ERR_ARG, "variable was not written (cc_write_binary = F)");
    return &it->second[(size_t)id * tree.box_len()];
  }
};

namespace detail {

struct cursor_t {
  const std::vector<char>& buf;
  size_t pos = 0;
  explicit cursor_t(const std::vector<char>& b) : buf(b) {}
  void need(size_t n) const {
    if (pos + n > buf.size()) throw error(AFMG_ERR_ARG, "af_read_tree: unexpected end of file (wrong NDIM?)");
  }
  int32_t i4() {
    need(4);
    int32_t v;
    std::memcpy(&v, &buf[pos], 4);
    pos += 4;
    return v;
  }
  void i4(int32_t* out, size_t n) {
    need(4 * n);
    std::memcpy(out, &buf[pos], 4 * n);
    pos += 4 * n;
  }
  std::vector<int32_t> i4v(long n) {
    if (n < 0) throw error(AFMG_ERR_ARG, "af_read_tree: negative count (wrong NDIM?)");
    std::vector<int32_t> v((size_t)n);
    if (n) i4(v.data(), (size_t)n);
    return v;
  }
  void f8(double* out, size_t n) {
    need(8 * n);
    std::memcpy(out, &buf[pos], 8 * n);
    pos += 8 * n;
  }
  std::vector<double> f8v(long n) {
    if (n < 0) throw error(AFMG_ERR_ARG, "af_read_tree: negative count (wrong NDIM?)");
    std::vector<double> v((size_t)n);
    if (n) f8(v.data(), (size_t)n);
    return v;
  }
  std::vector<std::string> names(int n_total, int n_used) {
    const int nlen = 20;
    need((size_t)n_total * nlen);
    std::vector<std::string> out;
    for (int i = 0; i < n_used; ++i) {
      std::string s(&buf[pos + (size_t)i * nlen], nlen);
      while (!s.empty() && s.back() == ' ') s.pop_back();
      out.push_back(s);
    }
    pos += (size_t)n_total * nlen;
    return out;
  }
};

inline dat_t parse_dat(const std::vector<char>& buf, int nd) {
  const int max_vars = 1024;
  cursor_t c(buf);
  dat_t d;
  af_t& t = d.tree;
  const int version = c.i4();
  if (version != 3) throw error(AFMG_ERR_ARG, "af_read_tree: incompatible file versions (read " + std::to_string(version) + ", required 3)");
  d.ready = c.i4() != 0;
  d.box_limit = c.i4();
  t.highest_lvl = c.i4();
  t.highest_id = c.i4();
  t.n_cell = c.i4();
  const int n_var_cell = c.i4(), n_var_face = c.i4();
  t.coord_t = c.i4();
  t.ndim = nd;
  const int nc = t.n_cell, n = t.highest_id;
  if (!(t.highest_lvl > 0 && t.highest_lvl <= 30 && n > 0 && n <= d.box_limit && nc >= 2 && nc <= 1024 && nc % 2 == 0 &&
        n_var_cell >= 0 && n_var_cell <= max_vars && n_var_face >= 0 && n_var_face <= max_vars))
    throw error(AFMG_ERR_ARG, "af_read_tree: implausible header");
  int32_t cgs[3] = {1, 1, 1}, per[3] = {0, 0, 0};
  c.i4(cgs, nd);
  c.i4(per, nd);
  c.f8(t.r_base, nd);
  c.f8(t.dr_base, nd);
  for (int q = 0; q < nd; ++q) {
    t.coarse_grid_size[q] = cgs[q];
    t.periodic[q] = per[q] != 0;
    if (cgs[q] < nc || cgs[q] % nc || !(t.dr_base[q] > 0)) throw error(AFMG_ERR_ARG, "af_read_tree: implausible coarse grid (wrong NDIM?)");
  }
  d.cc_names = c.names(max_vars, n_var_cell);
  d.fc_names = c.names(max_vars, n_var_face);
  c.i4v(max_vars);  // cc_num_copies
  c.i4v(max_vars);  // cc_write_output
  const std::vector<int32_t> cc_wb = c.i4v(max_vars), fc_wb = c.i4v(max_vars);
  d.cc_write_binary.assign(cc_wb.begin(), cc_wb.begin() + n_var_cell);
  c.i4v(c.i4());  // removed ids
  t.lvl_ids.assign(t.highest_lvl + 1, {});
  for (int l = 1; l <= t.highest_lvl; ++l) {
    t.lvl_ids[l] = c.i4v(c.i4());
    c.i4v(c.i4());  // leaves
    c.i4v(c.i4());  // parents
  }
  const int nch = 1 << nd, nnb = 2 * nd, nm = nd == 3 ? 27 : 9;
  size_t box_len = 1, fc_len = nd, nface = 1, ncell = 1;
  for (int q = 0; q < nd; ++q) box_len *= nc + 2, fc_len *= nc + 1, ncell *= nc;
  for (int q = 1; q < nd; ++q) nface *= nc;
  const size_t N = (size_t)n + 1;
  t.lvl.assign(N, 0);
  d.tag.assign(N, 0);
  t.ix.assign(N * nd, 0);
  t.parent.assign(N, 0);
  t.children.assign(N * nch, 0);
  t.neighbors.assign(N * nnb, 0);
  t.neighbor_mat.assign(N * nm, 0);
  t.dr.assign(N * nd, 0.0);
  t.r_min.assign(N * nd, 0.0);
  std::vector<int> cc_written, fc_written;
  for (int iv = 0; iv < n_var_cell; ++iv)
    if (cc_wb[iv]) {
      cc_written.push_back(iv + 1);
      d.cc[iv + 1].assign(N * box_len, 0.0);
    }
  for (int iv = 0; iv < n_var_face; ++iv)
    if (fc_wb[iv]) fc_written.push_back(iv + 1);
  std::vector<double> skip_fc(fc_len);
  for (int id = 1; id <= n; ++id) {
    if (c.i4() == 0) continue;  // box%in_use
    const int b_nc = c.i4(), n_bc = c.i4(), n_st = c.i4();
    if (b_nc != nc || n_bc < 0 || n_bc > nnb || n_st < 0)
      throw error(AFMG_ERR_ARG, "af_read_tree: implausible box record " + std::to_string(id) + " (wrong NDIM?)");
    t.lvl[id] = c.i4();
    d.tag[id] = c.i4();
    c.i4(&t.ix[(size_t)id * nd], nd);
    t.parent[id] = c.i4();
    c.i4(&t.children[(size_t)id * nch], nch);
    c.i4(&t.neighbors[(size_t)id * nnb], nnb);
    c.i4(&t.neighbor_mat[(size_t)id * nm], nm);
    c.f8(&t.dr[(size_t)id * nd], nd);
    c.f8(&t.r_min[(size_t)id * nd], nd);
    c.i4();  // box%coord_t
    for (int iv : cc_written) c.f8(&d.cc[iv][(size_t)id * box_len], box_len);
    for (size_t q = 0; q < fc_written.size(); ++q) c.f8(skip_fc.data(), fc_len);
    if (n_bc > 0) {
      dat_bc_t b;
      b.bc_index_to_nb = c.i4v(n_bc);
      b.nb_to_bc_index = c.i4v(nnb);
      b.bc_type = c.i4v((long)n_var_cell * n_bc);
      b.bc_val = c.f8v((long)(nface * n_var_cell) * n_bc);
      c.f8v((long)(nd * nface) * n_bc);  // bc_coords
      d.bc[id] = std::move(b);
    }
    for (int s = 0; s < n_st; ++s) {
      dat_stencil_t st;
      st.key = c.i4();
      st.shape = c.i4();
      st.stype = c.i4();
      st.cylindrical_gradient = c.i4() != 0;
      int k = c.i4();
      if (k > 0) st.c = c.f8v(k);
      k = c.i4();
      if (k > 0) {
        st.n_coeff = k;
        st.v = c.f8v((long)k * (long)ncell);
      }
      if (c.i4() > 0) st.f = c.f8v((long)ncell);
      if (c.i4() > 0) st.bc_correction = c.f8v((long)ncell);
      k = c.i4();
      if (k > 0) {
        st.n_sparse = k;
        st.sparse_ix = c.i4v((long)nd * k);
      }
      const int m = c.i4();
      if (k > 0 && m > 0) {
        st.m_sparse = m;
        st.sparse_v = c.f8v((long)m * k);
      }
      d.stencils[id].push_back(std::move(st));
    }
  }
  const bool other_present = c.i4() != 0;
  if (!other_present && c.pos != buf.size()) throw error(AFMG_ERR_ARG, "af_read_tree: trailing bytes (wrong NDIM?)");
  return d;
}

}  // namespace detail

// af_read_tree (afivo/src/m_af_output.f90:197-373); ndim = 0: try 3, then 2
inline dat_t read_tree(const std::string& path, int ndim = 0) {
  std::ifstream fh(path, std::ios::binary);
  if (!fh) throw error(AFMG_ERR_ARG, "af_read_tree: cannot open " + path);
  std::vector<char> buf((std::istreambuf_iterator<char>(fh)), std::istreambuf_iterator<char>());
  if (ndim) return detail::parse_dat(buf, ndim);
  std::string errors;
  for (int nd : {3, 2}) {
    try {
      return detail::parse_dat(buf, nd);
    } catch (const error& e) {
      errors += " NDIM=" + std::to_string(nd) + ": " + e.what() + ";";
    }
  }
  throw error(AFMG_ERR_ARG, "af_read_tree: the file parses with neither NDIM;" + errors);
}

// The stencils of a file as afmg_set_stencils takes them (the C++ twin of DatFile.stencil_entries): boxes tagged
// mg_normal_box whose stored stencils are constant get no entry and run through the fast kernels
inline stencil_set_t dat_stencil_set(const dat_t& d, int operator_key = 1, int prolongation_key = 2, int operator_mask = -1) {
  stencil_set_t out;
  const af_t& t = d.tree;
  for (int l = 1; l <= t.highest_lvl; ++l)
    for (int32_t id : t.lvl_ids[l]) {
      afmg_stencil_desc e{};
      e.box_id = id;
      e.tag = d.tag[id] >= 0 ? d.tag[id] : 0;
      e.f_offset = -1;
      const dat_stencil_t *op = nullptr, *pr = nullptr;
      auto it = d.stencils.find(id);
      if (it != d.stencils.end())
        for (const dat_stencil_t& st : it->second) {
          if (st.key == operator_key && st.shape == 1) op = &st;
          else if (st.key == prolongation_key && (st.shape == AFMG_STENCIL_P234 || st.shape == AFMG_STENCIL_P248)) pr = &st;
        }
      const bool plain = (e.tag & operator_mask) == 0 && (!op || op->stype == 1) && (!pr || pr->stype == 1) && (!op || op->f.empty());
      if (plain || (!op && !pr && !e.tag)) continue;
      auto put = [&out](const std::vector<double>& a) {
        const int64_t off = (int64_t)out.blob.size();
        out.blob.insert(out.blob.end(), a.begin(), a.end());
        return off;
      };
      if (op) {
        e.op_stype = op->stype;
        e.cylindrical_gradient = op->cylindrical_gradient;
        e.op_offset = put(op->stype == 1 ? op->c : op->v);
        if (!op->f.empty()) e.f_offset = put(op->f);
      }
      if (pr) {
        e.prolong_stype = pr->stype;
        e.prolong_shape = pr->shape;
        e.prolong_offset = put(pr->stype == 1 ? pr->c : pr->v);
      }
      out.desc.push_back(e);
    }
  return out;
}

// Set a solver up from a .dat file alone (the C++ twin of mg.mg_from_dat): mg's options are the caller's; boundary
// conditions, phi, rhs (eps if named) and the stencils come from the file.  The file's tree must outlive the solver.
inline void mg_from_dat(const dat_t& d, mg_t& mg, const std::string& phi = "phi", const std::string& rhs = "rhs",
                        const std::string& eps = "", const std::string& lsf = "lsf") {
  const af_t& t = d.tree;
  const int iv_phi = d.var_index(phi);
  // the boundary conditions the last ghost-cell fill stored in the boxes, as the per-cell callback
  mg.sides_bc_coords = [&d, iv_phi](int id, int nb, const std::vector<double>&, std::vector<double>& vals, int& type) {
    const dat_bc_t& b = d.bc.at(id);
    const int q = b.nb_to_bc_index[nb - 1] - 1;
    const size_t nvar = d.cc_names.size(), nface = vals.size();
    type = b.bc_type[(size_t)q * nvar + (iv_phi - 1)];
    std::copy(&b.bc_val[((size_t)q * nvar + (iv_phi - 1)) * nface], &b.bc_val[((size_t)q * nvar + iv_phi) * nface], vals.begin());
  };
  mg_init(t, mg);
  const stencil_set_t st = dat_stencil_set(d, 1, 2, mg.operator_mask);
  if (!st.desc.empty())
    mg.check(afmg_set_stencils(mg.h, (int32_t)st.desc.size(), st.desc.data(), st.blob.data(), (int64_t)st.blob.size()),
             "afmg_set_stencils");
  const std::vector<int32_t> ids = t.ids(false);
  std::vector<double> packed(ids.size() * t.box_len());
  auto send = [&](int var, int iv) {
    for (size_t n = 0; n < ids.size(); ++n) std::copy(d.cc_of(iv, ids[n]), d.cc_of(iv, ids[n]) + t.box_len(), packed.begin() + n * t.box_len());
    mg.set_cc(var, ids, packed.data());
  };
  send(AFMG_PHI, iv_phi);
  send(AFMG_RHS, d.var_index(rhs));
  if (!eps.empty()) send(AFMG_EPS, d.var_index(eps));
  // level-set distance stencils (mg_lsf_distance_key = 31) with the stored level-set values at their cells
  std::vector<int32_t> lids, n_entries, cell_ix;
  std::vector<double> dd, lv;
  int iv_lsf = 0;
  for (size_t i = 0; i < d.cc_names.size(); ++i)
    if (d.cc_names[i] == lsf && d.cc.count((int)i + 1)) iv_lsf = (int)i + 1;
  const int nd = t.ndim, n2 = t.n_cell + 2;
  for (int32_t id : ids) {
    auto it = d.stencils.find(id);
    if (it == d.stencils.end()) continue;
    for (const dat_stencil_t& s : it->second)
      if (s.key == 31 && s.n_sparse > 0) {
        lids.push_back(id);
        n_entries.push_back(s.n_sparse);
        cell_ix.insert(cell_ix.end(), s.sparse_ix.begin(), s.sparse_ix.end());
        dd.insert(dd.end(), s.sparse_v.begin(), s.sparse_v.end());
        if (iv_lsf)
          for (int k = 0; k < s.n_sparse; ++k) {
            size_t lin = 0;
            for (int q = nd - 1; q >= 0; --q) lin = lin * n2 + s.sparse_ix[(size_t)k * nd + q];
            lv.push_back(d.cc_of(iv_lsf, id)[lin]);
          }
      }
  }
  if (!lids.empty())
    mg.check(afmg_set_lsf_distances(mg.h, (int32_t)lids.size(), lids.data(), n_entries.data(), cell_ix.data(), dd.data(),
                                    iv_lsf ? lv.data() : nullptr), "afmg_set_lsf_distances");
}

}  // namespace afmg
