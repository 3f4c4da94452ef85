"""ctypes front-end of the CPU oracle (oracle/afmg_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libafmg_oracle.so")

I_PHI, I_RHS, I_TMP, I_EPS, I_FLD = 0, 1, 2, 3, 4
MG_CYCLE_DOWN, MG_CYCLE_UP = 1, 3
MG_PROLONG_LINEAR, MG_PROLONG_SPARSE, MG_PROLONG_AUTO = 17, 18, 19


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "afmg_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_tree_maxabs.restype = C.c_double
        L.orc_tree_sum.restype = C.c_double
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One ``af_t`` + ``mg_t`` pair held by the CPU restatement."""

    def __init__(self, tree, *, with_eps=False, n_cycle_down=2, n_cycle_up=2, use_corners=False,
                 subtract_mean=False, helmholtz_lambda=0.0, lsf_boundary_value=0.0,
                 operator_mask=-1, prolongation_type=MG_PROLONG_AUTO):
        self.tree = tree
        self.L = lib()
        t = tree
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        counts = i32([len(a) for a in t.lvl_ids])
        concat = i32(np.concatenate(t.lvl_ids))
        self.h = C.c_void_p(self.L.orc_create(
            t.ndim, t.nc, t.coord_t, t.highest_lvl, t.highest_id, _ip(counts), _ip(concat),
            _ip(i32(t.lvl)), _ip(i32(t.ix)), _ip(i32(t.parent)), _ip(i32(t.children)),
            _ip(i32(t.neighbors)), _ip(i32(t.neighbor_mat)), _dp(f64(t.r_min)), _dp(f64(t.dr)),
            _ip(i32(t.coarse_grid_size)), _dp(f64(t.dr_base)), _dp(f64(t.r_base)), int(with_eps)))
        self.opts = dict(n_cycle_down=n_cycle_down, n_cycle_up=n_cycle_up, use_corners=use_corners,
                         subtract_mean=subtract_mean, helmholtz_lambda=helmholtz_lambda,
                         lsf_boundary_value=lsf_boundary_value, operator_mask=operator_mask,
                         prolongation_type=prolongation_type)
        self._push_opts()

    def _push_opts(self):
        o = self.opts
        self.L.orc_set_opts(self.h, o["n_cycle_down"], o["n_cycle_up"], int(o["use_corners"]),
                            int(o["subtract_mean"]), C.c_double(o["helmholtz_lambda"]),
                            C.c_double(o["lsf_boundary_value"]), o["operator_mask"], o["prolongation_type"])

    def set_opts(self, **kw):
        self.opts.update(kw)
        self._push_opts()

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- data ----
    def set_cc(self, var, ids, data):
        ids = np.ascontiguousarray(ids, np.int32)
        data = np.ascontiguousarray(data, np.float64).reshape(len(ids), -1)
        assert data.shape[1] == self.tree.box_len
        self.L.orc_set_cc(self.h, var, len(ids), _ip(ids), _dp(data))

    def get_cc(self, var, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.empty((len(ids), self.tree.box_len))
        self.L.orc_get_cc(self.h, var, len(ids), _ip(ids), _dp(out))
        return out

    def set_bc(self, bc):
        ids = np.ascontiguousarray(bc.ids, np.int32)
        nbs = np.ascontiguousarray(bc.nbs, np.int32)
        types = np.ascontiguousarray(bc.types, np.int32)
        vals = np.ascontiguousarray(bc.vals, np.float64)
        self.L.orc_set_bc(self.h, len(ids), _ip(ids), _ip(nbs), _ip(types), _dp(vals))

    def set_lsf_distances(self, ids, dd):
        ids = np.ascontiguousarray(ids, np.int32)
        dd = np.ascontiguousarray(dd, np.float64)
        self.L.orc_set_lsf_distances(self.h, len(ids), _ip(ids), _dp(dd))

    # ---- solver ----
    def mg_init(self):
        rc = self.L.orc_mg_init(self.h)
        if rc != 0:
            raise RuntimeError(f"oracle mg_init failed with code {rc}")

    def fas_vcycle(self, set_residual=True, highest_lvl=0, standalone=True):
        self.L.orc_fas_vcycle(self.h, int(set_residual), int(highest_lvl), int(standalone))

    def fas_fmg(self, set_residual=True, have_guess=False):
        self.L.orc_fas_fmg(self.h, int(set_residual), int(have_guess))

    # ---- single operations ----
    def box_gsrb_lvl(self, lvl, redblack):
        self.L.orc_box_gsrb_lvl(self.h, lvl, redblack)

    def gsrb_boxes(self, lvl, type_cycle):
        self.L.orc_gsrb_boxes(self.h, lvl, type_cycle)

    def gc_lvl(self, lvl, var=I_PHI, corners=True):
        self.L.orc_gc_lvl(self.h, lvl, var, int(corners))

    def update_coarse(self, lvl, with_tmp=True):
        self.L.orc_update_coarse(self.h, lvl, int(with_tmp))

    def correct_children(self, lvl_parents):
        self.L.orc_correct_children(self.h, lvl_parents)

    def residual_lvl(self, lvl):
        self.L.orc_residual_lvl(self.h, lvl)

    def solve_coarse_grid(self):
        self.L.orc_solve_coarse_grid(self.h)

    def init_phi_rhs(self):
        self.L.orc_init_phi_rhs(self.h)

    def maxabs(self, var=I_TMP):
        return float(self.L.orc_tree_maxabs(self.h, var))

    def tree_sum(self, var=I_PHI):
        return float(self.L.orc_tree_sum(self.h, var))

    # ---- field from potential (mg_compute_phi_gradient & co) ----
    def fc_len(self):
        t = self.tree
        return t.ndim * (t.nc + 1) ** t.ndim

    def compute_phi_gradient(self, fac, with_norm=True):
        self.L.orc_compute_phi_gradient(self.h, C.c_double(fac), int(with_norm))

    def compute_field_norm(self):
        self.L.orc_compute_field_norm(self.h)

    def gc_tree(self, var, corners=True):
        self.L.orc_gc_tree(self.h, var, int(corners))

    def get_fc(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.empty((len(ids), self.fc_len()))
        self.L.orc_get_fc(self.h, len(ids), _ip(ids), _dp(out))
        return out

    def set_fc(self, ids, data):
        ids = np.ascontiguousarray(ids, np.int32)
        data = np.ascontiguousarray(data, np.float64).reshape(len(ids), self.fc_len())
        self.L.orc_set_fc(self.h, len(ids), _ip(ids), _dp(data))

    def set_lsf_boundary_values(self, ids, vals):
        """mg%lsf_boundary_function evaluated at the cell centres of the listed boxes; call mg_init() again
        afterwards (bc_correction = f * value is rebuilt there, m_af_multigrid.f90:1171-1174)."""
        ids = np.ascontiguousarray(ids, np.int32)
        vals = np.ascontiguousarray(vals, np.float64).reshape(len(ids), self.tree.nc ** self.tree.ndim)
        self.L.orc_set_lsf_boundary_values(self.h, len(ids), _ip(ids), _dp(vals))

    def set_lsf_prolong_distances(self, ids, dd):
        """mg%lsf_use_custom_prolongation: distances from every fine cell to its ndim+1 coarse prolongation points
        (dd[..., 0] < 0: cell outside the root mask); before mg_init()."""
        ids = np.ascontiguousarray(ids, np.int32)
        dd = np.ascontiguousarray(dd, np.float64).reshape(len(ids), -1)
        self.L.orc_set_lsf_prolong_distances(self.h, len(ids), _ip(ids), _dp(dd))

    def set_lsf_cc(self, ids, vals):
        ids = np.ascontiguousarray(ids, np.int32)
        vals = np.ascontiguousarray(vals, np.float64).reshape(len(ids), self.tree.nc ** self.tree.ndim)
        self.L.orc_set_lsf_cc(self.h, len(ids), _ip(ids), _dp(vals))

    def set_fld_bc(self, bc):
        ids = np.ascontiguousarray(bc.ids, np.int32)
        nbs = np.ascontiguousarray(bc.nbs, np.int32)
        types = np.ascontiguousarray(bc.types, np.int32)
        vals = np.ascontiguousarray(bc.vals, np.float64)
        self.L.orc_set_fld_bc(self.h, len(ids), _ip(ids), _ip(nbs), _ip(types), _dp(vals))

    # ---- stencil read-back ----
    def op_stencil(self, box_id):
        t = self.tree
        ncell = t.nc ** t.ndim
        ncf = 2 * t.ndim + 1
        out = np.zeros(ncf * ncell)
        f = np.zeros(ncell)
        has_f = C.c_int(0)
        cyl = C.c_int(0)
        stype = self.L.orc_get_op_stencil(self.h, int(box_id), _dp(out), out.size, _dp(f), C.byref(has_f), C.byref(cyl))
        coeff = out[:ncf] if stype == 1 else out.reshape(ncell, ncf)
        return stype, coeff, (f if has_f.value else None), bool(cyl.value)

    def prolong_stencil(self, box_id):
        t = self.tree
        ncell = t.nc ** t.ndim
        out = np.zeros((1 << t.ndim) * ncell)
        shape = C.c_int(0)
        stype = self.L.orc_get_prolong_stencil(self.h, int(box_id), _dp(out), out.size, C.byref(shape))
        ncf = (1 << t.ndim) if shape.value == 3 else t.ndim + 1
        coeff = out[:ncf] if stype == 1 else out[: ncf * ncell].reshape(ncell, ncf)
        return stype, shape.value, coeff

    def tag(self, box_id):
        return int(self.L.orc_get_tag(self.h, int(box_id)))

    def num_threads(self):
        return int(self.L.orc_num_threads())

    def set_num_threads(self, n):
        self.L.orc_set_num_threads(int(n))
