"""Host-side construction of the per-box operator / prolongation stencils: the Python mirror of
``mg_set_operators_lvl`` / ``mg_set_box_tag`` (afivo/src/m_af_multigrid.f90:1100-1185), for trees that do not come
from the Fortran reference (which ships the stencils it built itself, fortran/m_af_multigrid_gpu.f90) or from a
``.dat`` file (datfile.DatFile.stencil_entries).  The arithmetic is the library's (``afmg_build_box_*`` of
include/afmg.h, pure host functions of libafmg.so); this module only walks the tree.

    entries, lsf = build_stencils(tree, eps_cc=eps, lsf=my_lsf)      # eps: (highest_id + 1, nc+2, ...) or None
    mg.set_stencils(entries)                                          # afmg_set_stencils
    if lsf is not None: mg.set_lsf_distances(lsf.ids, lsf.dd, lsf.lsf_cells)   # for the field at the electrode
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _lib
from .tree import Tree

MG_LSF_BOX, MG_VEPS_BOX, MG_CEPS_BOX = 1, 2, 4  # m_af_types.f90:497-508
MG_PROLONG_LINEAR, MG_PROLONG_SPARSE, MG_PROLONG_AUTO = 17, 18, 19
STENCIL_P234, STENCIL_P248 = 2, 3
LSF_DIST_LINEAR, LSF_DIST_GSS = 0, 1


@dataclass
class LsfData:
    """What store_lsf_distance_matrix (:977-1097) keeps per box, in dense form."""
    ids: np.ndarray          # boxes with at least one boundary cell (they carry mg_lsf_box)
    dd: np.ndarray           # (len(ids), nc^ndim, 2*ndim) relative distances, 1 = no boundary
    lsf_cells: np.ndarray    # (len(ids), nc^ndim) the level-set function at the cell centres
    root_mask: dict          # box id -> (nc^ndim,) uint8: the mg_lsf_mask_key stencil, for every box that has one
    pdd: Optional[dict] = None  # box id -> (nc^ndim, ndim+1) coarse-point distances (custom prolongation only)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def lsf_opts(dist_method=LSF_DIST_LINEAR, **kw) -> _lib.LsfOpts:
    """mg_t's lsf_* defaults (m_af_types.f90:607-619) with overrides: gradient_safety_factor, length_scale, tol,
    min_rel_distance."""
    o = _lib.LsfOpts()
    _lib.lib().afmg_lsf_opts_default(C.byref(o))
    o.dist_method = dist_method
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown level-set option {k}")
        setattr(o, k, v)
    return o


ELECTRODE_TYPES = {"sphere": 1, "rod": 2, "rod_cone_top": 3, "rod_rod": 4, "sphere_rod": 5,
                   "two_rod_cone_electrodes": 6, "coaxial": 7}  # field_electrode_type, src/m_field.f90:254-362


def electrode(kind: str, ndim: int, **params) -> _lib.Electrode:
    """One of the streamer code's built-in electrode shapes (src/m_field.f90:686-904) as a C-side level-set function:
    pass the result as `lsf` to lsf_distances / build_stencils / mg_set_operators_tree.  params: rod_r0, rod_r1,
    rod_radius, cone_tip_radius, cone_length_frac, rod2_*, cone2_*, domain_center, current_voltage,
    electrode_grounded, electrode2_grounded.  Raises AFMG_ERR_ARG where the reference would `error stop`."""
    e = _lib.Electrode()
    e.type, e.ndim = ELECTRODE_TYPES[kind], ndim
    for k, v in params.items():
        if not hasattr(e, k) or k in ("type", "ndim", "cone_tip_center", "cone2_tip_center", "cone_tip_r_curvature",
                                      "cone2_tip_r_curvature"):  # the last four are derived by afmg_electrode_prepare
            raise TypeError(f"unknown electrode parameter {k}")
        if isinstance(getattr(e, k), C.Array):
            for d, x in enumerate(v):
                getattr(e, k)[d] = float(x)
        else:
            setattr(e, k, v)
    _check(_lib.lib().afmg_electrode_prepare(C.byref(e)), "afmg_electrode_prepare")
    return e


def _callback(lsf, nd):
    """(afmg_lsf_fn, user pointer, python callable of one point) for a Python function or a built-in electrode."""
    L = _lib.lib()
    if isinstance(lsf, _lib.Electrode):
        user = C.cast(C.pointer(lsf), C.c_void_p)

        def point(r):
            a = (C.c_double * 3)(*[float(x) for x in r], *([0.0] * (3 - len(r))))
            return L.afmg_electrode_lsf(a, user)

        return C.cast(L.afmg_electrode_lsf, _lib.LSF_FN), user, point
    return _lib.LSF_FN(lambda r, _u: float(lsf(np.array([r[d] for d in range(nd)])))), None, lsf


def electrode_potential(el: _lib.Electrode, points: np.ndarray) -> np.ndarray:
    """mg%lsf_boundary_function at the given points (..., ndim) -- the values afmg_set_lsf_boundary_values takes."""
    L = _lib.lib()
    pts = np.asarray(points, np.float64)
    out = np.empty(pts.shape[:-1])
    user = C.cast(C.pointer(el), C.c_void_p)
    flat = pts.reshape(-1, pts.shape[-1])
    res = out.reshape(-1)
    for n, r in enumerate(flat):
        a = (C.c_double * 3)(*r, *([0.0] * (3 - len(r))))
        res[n] = L.afmg_electrode_potential(a, user)
    return out


def _check(rc, what):
    if rc != 0:
        raise _lib.AfmgError(rc, what)


def lsf_distances(tree: Tree, lsf: Callable[[np.ndarray], float], opts: Optional[_lib.LsfOpts] = None,
                  custom_prolongation: bool = False) -> Optional[LsfData]:
    """mg_set_box_tag's level-set half for every box: root mask and distance matrix; `lsf` takes one point (ndim,)
    and returns the level-set value there (mg%lsf)."""
    L = _lib.lib()
    nd, nc = tree.ndim, tree.nc
    ncell = nc ** nd
    opts = opts or lsf_opts()
    cb, user, lsf = _callback(lsf, nd)
    ids, dds, vals, masks, pdds = [], [], [], {}, {}
    for lvl_ids in tree.lvl_ids:
        for b in lvl_ids:
            b = int(b)
            rmin = np.ascontiguousarray(tree.r_min[b], np.float64)
            dr = np.ascontiguousarray(tree.dr[b], np.float64)
            mask = np.zeros(ncell, np.uint8)
            dd = np.empty((ncell, 2 * nd))
            nb = C.c_int32(0)
            _check(L.afmg_build_box_lsf_distances(nd, nc, _dp(rmin), _dp(dr), cb, user, C.byref(opts), None,
                                                  mask.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(dd), C.byref(nb)),
                   "afmg_build_box_lsf_distances")
            if mask.any():
                masks[b] = mask
            if nb.value == 0:
                continue
            ids.append(b)
            dds.append(dd)
            idx = np.arange(ncell)
            centre = np.stack([rmin[d] + ((idx // nc ** d) % nc + 0.5) * dr[d] for d in range(nd)], axis=-1)
            vals.append(np.array([lsf(c) for c in centre]))
            if custom_prolongation and tree.parent[b] > 0:
                p = int(tree.parent[b])
                pdd = np.empty((ncell, nd + 1))
                ix = np.ascontiguousarray(tree.ix[b], np.int32)
                _check(L.afmg_build_box_lsf_prolong_distances(
                    nd, nc, _dp(rmin), _dp(dr), _ip(ix), _dp(np.ascontiguousarray(tree.r_min[p], np.float64)),
                    _dp(np.ascontiguousarray(tree.dr[p], np.float64)), cb, user, C.byref(opts),
                    mask.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(pdd)), "afmg_build_box_lsf_prolong_distances")
                pdds[b] = pdd
    if not ids:
        return None
    return LsfData(np.array(ids, np.int32), np.stack(dds), np.stack(vals), masks, pdds if custom_prolongation else None)


def build_stencils(tree: Tree, *, eps_cc: Optional[np.ndarray] = None, lsf: Optional[Callable] = None,
                   lsf_data: Optional[LsfData] = None, lsf_options: Optional[_lib.LsfOpts] = None,
                   operator_mask: int = -1, prolongation_type: int = MG_PROLONG_AUTO,
                   lsf_use_custom_prolongation: bool = False):
    """Entries for mg_t.set_stencils: one per box whose iand(tag, operator_mask) is not mg_normal_box, as
    mg_store_operator_stencil (:823-859) and mg_store_prolongation_stencil (:862-903) would store them.

    eps_cc: permittivity cc(:, ..., mg_i_eps) of every box, indexed by box id, ghost cells included (the reference
    fills them with af_gc_tree before mg_init).  lsf: mg%lsf, or pass precomputed `lsf_data` (e.g. distances read
    from a .dat file).  Returns (entries, lsf_data)."""
    L = _lib.lib()
    nd, nc = tree.ndim, tree.nc
    ncell = nc ** nd
    if lsf is not None and lsf_data is None:
        lsf_data = lsf_distances(tree, lsf, lsf_options, lsf_use_custom_prolongation)
    dd_of = {int(b): lsf_data.dd[n] for n, b in enumerate(lsf_data.ids)} if lsf_data is not None else {}
    if (lsf is not None and lsf_data is None) or \
            (lsf_data is not None and not any(tree.lvl[int(b)] == 1 for b in lsf_data.ids)):
        # check_coarse_representation_lsf (m_af_multigrid.f90:2142-2161): the reference stops here, because a
        # coarse grid that does not see the electrode makes the cycles diverge
        raise _lib.AfmgError(-1, "level set function not resolved on coarse grid: no roots found on level 1, "
                                 "use a finer coarse grid")
    entries = []
    for lvl, lvl_ids in enumerate(tree.lvl_ids, start=1):
        for b in lvl_ids:
            b = int(b)
            eps = None if eps_cc is None else np.ascontiguousarray(eps_cc[b], np.float64).reshape(-1)
            tag = L.afmg_build_box_tag(nd, nc, None if eps is None else _dp(eps), int(b in dd_of))
            _check(min(tag, 0), "afmg_build_box_tag")
            masked = tag & operator_mask
            if tag == 0:
                continue
            e = dict(box_id=b, tag=tag, op=None, f=None, cyl=False)
            if masked != 0:
                v = np.empty((ncell, 2 * nd + 1))
                f = np.empty(ncell)
                stype, has_f, cyl = C.c_int32(0), C.c_int32(0), C.c_int32(0)
                dd = np.ascontiguousarray(dd_of[b]).reshape(-1) if masked & MG_LSF_BOX else None
                _check(L.afmg_build_box_operator(
                    nd, nc, tree.coord_t, masked, _dp(np.ascontiguousarray(tree.dr[b], np.float64)),
                    _dp(np.ascontiguousarray(tree.r_min[b], np.float64)), None if eps is None else _dp(eps),
                    None if dd is None else _dp(dd), _dp(v), _dp(f), C.byref(stype), C.byref(has_f), C.byref(cyl)),
                    "afmg_build_box_operator")
                e["op"] = (stype.value, v[0].copy() if stype.value == 1 else v)
                e["f"] = f if has_f.value else None
                e["cyl"] = bool(cyl.value)
            if lvl > 1 and prolongation_type == MG_PROLONG_AUTO:
                pdd = None
                if lsf_use_custom_prolongation and lsf_data is not None and lsf_data.pdd is not None:
                    pdd = lsf_data.pdd.get(b)
                variable_eps = bool(masked & MG_VEPS_BOX)
                variable_lsf = (not variable_eps) and bool(masked & MG_LSF_BOX) and pdd is not None
                if variable_eps or variable_lsf:
                    pv = np.empty((ncell, nd + 1))
                    pst, psh = C.c_int32(0), C.c_int32(0)
                    p = int(tree.parent[b])
                    eps_p = None if not variable_eps else np.ascontiguousarray(eps_cc[p], np.float64).reshape(-1)
                    _check(L.afmg_build_box_prolongation(
                        nd, nc, masked, _ip(np.ascontiguousarray(tree.ix[b], np.int32)),
                        None if eps_p is None else _dp(eps_p),
                        None if pdd is None else _dp(np.ascontiguousarray(pdd).reshape(-1)), _dp(pv), C.byref(pst),
                        C.byref(psh)), "afmg_build_box_prolongation")
                    e["prolong"] = (pst.value, psh.value, pv[0].copy() if pst.value == 1 else pv)
            entries.append(e)
    return entries, lsf_data
