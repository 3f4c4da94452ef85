"""Host-side mirror of the reference's multigrid interface (afivo/src/m_af_multigrid.f90:14-38)
on top of the C ABI: same names, argument meaning and error behaviour (``error stop`` becomes
``AfmgError``).

    mg = mg_t(sides_bc=af_bc_dirichlet_zero)          ! type(mg_t) :: mg ; mg%sides_bc => ...
    mg_init(tree, mg)                                 ! call mg_init(tree, mg)
    mg.set_cc(I_RHS, ids, rhs)                        ! box%cc(:, :, :, i_rhs) = ...
    mg_fas_fmg(tree, mg, set_residual=True, have_guess=False)
    mg_fas_vcycle(tree, mg, set_residual=True)
    res = af_tree_maxabs_cc(tree, mg, I_TMP)

The tree (``tree.Tree``) is the flat copy of ``af_t``; cell data lives on the GPU and is
moved with ``set_cc`` / ``get_cc`` in the reference's own box layout cc(0:nc+1, ...).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Callable, Optional

import numpy as np

from . import _lib
from ._lib import AfmgError, Opts, TreeDesc
from .tree import Tree
from .workloads import (AF_BC_DIRICHLET, AF_BC_NEUMANN, BCTable, bc_table)

I_PHI, I_RHS, I_TMP, I_EPS, I_FLD, I_PHOTO = 0, 1, 2, 3, 4, 5
MG_CYCLE_DOWN, MG_CYCLE_UP = 1, 3
MG_PROLONG_LINEAR, MG_PROLONG_SPARSE, MG_PROLONG_AUTO = 17, 18, 19


# built-in boundary conditions (afivo/src/m_af_ghostcell.f90:615-652), as callbacks (nb, coords)
def af_bc_dirichlet_zero(nb, coords):
    return AF_BC_DIRICHLET, 0.0


def af_bc_neumann_zero(nb, coords):
    return AF_BC_NEUMANN, 0.0


@dataclasses.dataclass
class mg_t:
    """mg_t (afivo/src/m_af_types.f90:572-665): the options a caller may set before mg_init."""
    n_cycle_down: int = 2
    n_cycle_up: int = 2
    use_corners: bool = False
    subtract_mean: bool = False
    helmholtz_lambda: float = 0.0
    lsf_boundary_value: float = 0.0
    operator_mask: int = -1
    prolongation_type: int = MG_PROLONG_AUTO
    sides_bc: Optional[Callable] = None  # (nb, coords[n, nface, D]) -> (bc_type, values)
    device: int = -1
    # multi-GPU: (rank, world_size, allgather) with allgather(bytes) -> list of every rank's bytes, e.g.
    # built on torch.distributed (see comm_from_torch); None = single GPU
    comm: Optional[tuple] = None
    # single-process multi-GPU (afmg_opts.n_gpus): the handle drives devices device .. device + n_gpus - 1 itself
    n_gpus: int = 0
    initialized: bool = False
    _h: Optional[C.c_void_p] = None
    _tree: Optional[Tree] = None

    # ---- helpers -------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = _lib.lib().afmg_last_error(self._h)
            raise AfmgError(rc, msg.decode() if msg else "")

    def _need_init(self):
        if not self.initialized:
            raise AfmgError(-4, "mg_t not initialized (reference: error stop in mg_use)")

    def set_bc(self, bc: BCTable):
        """Ship the evaluated mg%sides_bc callback (one row per physical face)."""
        self._need_init()
        ids = np.ascontiguousarray(bc.ids, np.int32)
        nbs = np.ascontiguousarray(bc.nbs, np.int32)
        ty = np.ascontiguousarray(bc.types, np.int32)
        vals = np.ascontiguousarray(bc.vals, np.float64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(_lib.lib().afmg_set_bc(self._h, len(ids), ip(ids), ip(nbs), ip(ty),
                                           vals.ctypes.data_as(C.POINTER(C.c_double))))

    def set_cc(self, var, ids, data):
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        data = np.ascontiguousarray(data, np.float64)
        assert data.size == len(ids) * self._tree.box_len
        self._check(_lib.lib().afmg_upload(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                           data.ctypes.data))

    def get_cc(self, var, ids):
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.empty((len(ids),) + (self._tree.nc + 2,) * self._tree.ndim)
        self._check(_lib.lib().afmg_download(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                             out.ctypes.data))
        return out

    def upload_ptr(self, var, ids, host_ptr):
        """Upload from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        ids = np.ascontiguousarray(ids, np.int32)
        self._check(_lib.lib().afmg_upload(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), host_ptr))

    def upload_interior_ptr(self, var, ids, host_ptr):
        """Interior cells only (nc^ndim doubles per box): what the rhs needs."""
        ids = np.ascontiguousarray(ids, np.int32)
        self._check(_lib.lib().afmg_upload_interior(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    host_ptr))

    def set_cc_interior(self, var, ids, data):
        self._need_init()
        data = np.ascontiguousarray(data, np.float64)
        assert data.size == len(ids) * self._tree.nc ** self._tree.ndim
        self.upload_interior_ptr(var, ids, data.ctypes.data)

    def field_set_rhs(self, ids, charges, densities, on_device=False):
        """field_set_rhs (src/m_field.f90:406-444): rhs = sum_s charges[s] * densities[s] on the listed boxes, summed on
        the device in the reference's order.  densities: list of packed arrays (n, (nc+2)^ndim), or raw pointers
        (host or, with on_device, device memory)."""
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        q = np.ascontiguousarray(charges, np.float64)
        keep, ptrs = [], (C.c_void_p * len(densities))()
        for s, d in enumerate(densities):
            if isinstance(d, (int, np.integer)):
                ptrs[s] = int(d)
            else:
                a = np.ascontiguousarray(d, np.float64)
                assert a.size == len(ids) * self._tree.box_len
                keep.append(a)
                ptrs[s] = a.ctypes.data
        self._check(_lib.lib().afmg_field_set_rhs(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), len(q),
                                                  q.ctypes.data_as(C.POINTER(C.c_double)), ptrs, int(on_device)))

    def build_stencils_device(self, electrode=None, lsf_options=None):
        """mg_set_operators_tree on the device (afmg_build_stencils_device): from the uploaded AFMG_EPS and / or a
        built-in electrode (stencils.electrode(...)); replaces build_stencils + set_stencils + set_lsf_distances."""
        self._need_init()
        self._check(_lib.lib().afmg_build_stencils_device(
            self._h, C.byref(electrode) if electrode is not None else None,
            C.byref(lsf_options) if lsf_options is not None else None))

    def built_stencils(self):
        """(ids, tags, meta (n, 4), blob (n, 18 * nc^3)) of the last build_stencils_device, reference order:
        v(7, cells) | f(cells) | pv(4, cells) | dd(6, cells) per box"""
        L = _lib.lib()
        n = C.c_int32(0)
        self._check(L.afmg_built_stencils(self._h, C.byref(n), None, None, None, None))
        nb, ncell = n.value, self._tree.nc ** 3
        ids, tags, meta = np.zeros(nb, np.int32), np.zeros(nb, np.int32), np.zeros((nb, 4), np.int32)
        blob = np.zeros((nb, 18 * ncell))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        if nb:
            self._check(L.afmg_built_stencils(self._h, C.byref(n), ip(ids), ip(tags), ip(meta),
                                              blob.ctypes.data_as(C.POINTER(C.c_double))))
        return ids, tags, meta, blob

    def download_interior_ptr(self, var, ids, host_ptr):
        """Interior cells only (nc^ndim doubles per box) into a raw host pointer."""
        ids = np.ascontiguousarray(ids, np.int32)
        self._check(_lib.lib().afmg_download_interior(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      host_ptr))

    def get_cc_interior(self, var, ids):
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.empty((len(ids),) + (self._tree.nc,) * self._tree.ndim)
        self.download_interior_ptr(var, ids, out.ctypes.data)
        return out

    def download_ptr(self, var, ids, host_ptr):
        ids = np.ascontiguousarray(ids, np.int32)
        self._check(_lib.lib().afmg_download(self._h, var, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), host_ptr))

    def set_stencils(self, entries):
        """Ship per-box operator / prolongation stencils (what mg_set_operators_lvl stored in
        box%stencils, afivo/src/m_af_multigrid.f90:1147-1185).  entries: iterable of dicts with keys
        box_id, tag, op=(stype, coeff) or None, f=array or None, prolong=(stype, shape, coeff) or None;
        coeff in the reference layout (c(n) or v(n, cells) with n fastest)."""
        self._need_init()
        entries = list(entries)
        descs = (_lib.StencilDesc * max(1, len(entries)))()
        blob = []
        off = 0

        def put(a):
            nonlocal off
            a = np.ascontiguousarray(a, np.float64).reshape(-1)
            blob.append(a)
            off += a.size
            return off - a.size

        for d, e in zip(descs, entries):
            d.box_id, d.tag = int(e["box_id"]), int(e.get("tag", 0))
            d.cylindrical_gradient = int(bool(e.get("cyl", False)))
            d.op_stype, d.f_offset, d.prolong_shape = 0, -1, 0
            if e.get("op") is not None:
                d.op_stype = int(e["op"][0])
                d.op_offset = put(e["op"][1])
            if e.get("f") is not None:
                d.f_offset = put(e["f"])
            if e.get("prolong") is not None:
                d.prolong_stype, d.prolong_shape = int(e["prolong"][0]), int(e["prolong"][1])
                d.prolong_offset = put(e["prolong"][2])
        flat = np.concatenate(blob) if blob else np.zeros(1)
        self._check(_lib.lib().afmg_set_stencils(self._h, len(entries), C.byref(descs),
                                                 flat.ctypes.data_as(C.POINTER(C.c_double)), flat.size))

    def set_lsf_boundary_value(self, value):
        self.lsf_boundary_value = float(value)
        if self.initialized:
            self._check(_lib.lib().afmg_set_lsf_boundary_value(self._h, float(value)))

    def set_lsf_boundary_values(self, ids, values):
        """mg%lsf_boundary_function evaluated at the cell centres of the listed boxes (nc^ndim values per box);
        an empty list returns to the scalar mg%lsf_boundary_value."""
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        ncell = self._tree.nc ** self._tree.ndim
        vals = np.ascontiguousarray(values, np.float64).reshape(len(ids), ncell)
        self._check(_lib.lib().afmg_set_lsf_boundary_values(
            self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)), vals.ctypes.data_as(C.POINTER(C.c_double))))

    def clear(self, var):
        self._need_init()
        self._check(_lib.lib().afmg_clear(self._h, var))

    # ---- field from potential --------------------------------------
    def fc_len(self):
        t = self._tree
        return t.ndim * (t.nc + 1) ** t.ndim

    def get_fc(self, ids):
        """box%fc(nc+1, nc+1[, nc+1], NDIM) of the face-centred field, first index fastest."""
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.empty((len(ids), self.fc_len()))
        self._check(_lib.lib().afmg_download_fc(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                out.ctypes.data))
        return out

    def set_fc(self, ids, data):
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        data = np.ascontiguousarray(data, np.float64)
        assert data.size == len(ids) * self.fc_len()
        self._check(_lib.lib().afmg_upload_fc(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                              data.ctypes.data))

    def set_fld_bc(self, bc: BCTable):
        """Boundary condition of the field norm (cc_methods(i_electric_fld)%bc); default af_bc_neumann_zero."""
        self._need_init()
        ids = np.ascontiguousarray(bc.ids, np.int32)
        nbs = np.ascontiguousarray(bc.nbs, np.int32)
        ty = np.ascontiguousarray(bc.types, np.int32)
        vals = np.ascontiguousarray(bc.vals, np.float64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(_lib.lib().afmg_set_fld_bc(self._h, len(ids), ip(ids), ip(nbs), ip(ty),
                                               vals.ctypes.data_as(C.POINTER(C.c_double))))

    def set_lsf_distances(self, ids, dd, lsf=None):
        """Ship the sparse level-set distance stencils.  dd: dense all_distances(2*ndim, cells) per box as the
        oracle takes them; only cells with any(dd < 1) become entries, in IJK order
        (store_lsf_distance_matrix, afivo/src/m_af_multigrid.f90:1075-1080)."""
        self._need_init()
        t = self._tree
        nd, nc = t.ndim, t.nc
        ncell = nc ** nd
        ids = np.ascontiguousarray(ids, np.int32)
        dd = np.ascontiguousarray(dd, np.float64).reshape(len(ids), ncell, 2 * nd)
        lsf_a = None if lsf is None else np.ascontiguousarray(lsf, np.float64).reshape(len(ids), ncell)
        n_ent, cells, vals, lv = [], [], [], []
        for b in range(len(ids)):
            sel = np.nonzero((dd[b] < 1.0).any(axis=1))[0]
            n_ent.append(len(sel))
            ix = np.empty((len(sel), nd), np.int32)
            r = sel.copy()
            for d in range(nd):
                ix[:, d] = r % nc + 1
                r //= nc
            cells.append(ix)
            vals.append(dd[b][sel])
            lv.append(np.zeros(len(sel)) if lsf_a is None else lsf_a[b][sel])
        n_ent = np.ascontiguousarray(n_ent, np.int32)
        cells = np.ascontiguousarray(np.concatenate(cells) if cells else np.zeros((0, nd)), np.int32)
        vals = np.ascontiguousarray(np.concatenate(vals) if vals else np.zeros((0, 2 * nd)), np.float64)
        lv = np.ascontiguousarray(np.concatenate(lv) if lv else np.zeros(0), np.float64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(_lib.lib().afmg_set_lsf_distances(self._h, len(ids), ip(ids), ip(n_ent), ip(cells), dp(vals), dp(lv)))

    def set_lsf_distances_sparse(self, ids, n_entries, cell_ix, dd, lsf=None):
        """afmg_set_lsf_distances with the sparse stencils as stored (e.g. DatFile.lsf_distances())."""
        self._need_init()
        ids = np.ascontiguousarray(ids, np.int32)
        n_ent = np.ascontiguousarray(n_entries, np.int32)
        cells = np.ascontiguousarray(cell_ix, np.int32)
        vals = np.ascontiguousarray(dd, np.float64)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        lv = None if lsf is None else np.ascontiguousarray(lsf, np.float64)
        self._check(_lib.lib().afmg_set_lsf_distances(self._h, len(ids), ip(ids), ip(n_ent), ip(cells), dp(vals),
                                                      None if lv is None else dp(lv)))

    # ---- per-level building blocks (exported for parity tests) -----------------------
    def gsrb_boxes(self, lvl, type_cycle):
        self._check(_lib.lib().afmg_gsrb_boxes(self._h, lvl, type_cycle))

    def gsrb_halfsweep(self, lvl, redblack):
        self._check(_lib.lib().afmg_gsrb_halfsweep(self._h, lvl, redblack))

    def gc_lvl(self, lvl, var=I_PHI, corners=True):
        self._check(_lib.lib().afmg_gc_lvl(self._h, lvl, var, int(corners)))

    def update_coarse(self, lvl, with_tmp=True):
        self._check(_lib.lib().afmg_update_coarse(self._h, lvl, int(with_tmp)))

    def correct_children(self, lvl_parents):
        self._check(_lib.lib().afmg_correct_children(self._h, lvl_parents))

    def correct_children_gc(self, lvl_parents):
        self._check(_lib.lib().afmg_correct_children_gc(self._h, lvl_parents))

    def residual_lvl(self, lvl):
        self._check(_lib.lib().afmg_residual_lvl(self._h, lvl))

    def solve_coarse_grid(self):
        self._check(_lib.lib().afmg_solve_coarse_grid(self._h))

    def init_phi_rhs(self):
        self._check(_lib.lib().afmg_init_phi_rhs(self._h))

    # ---- asynchronous cycles + instrumentation (bench) ------------------------------
    def fas_vcycle_async(self, set_residual=True, highest_lvl=0, n_cycles=1):
        self._check(_lib.lib().afmg_fas_vcycle_async(self._h, int(set_residual), int(highest_lvl), n_cycles))

    def fas_fmg_async(self, set_residual=True, have_guess=True, n_cycles=1):
        self._check(_lib.lib().afmg_fas_fmg_async(self._h, int(set_residual), int(have_guess), n_cycles))

    def sync(self):
        self._check(_lib.lib().afmg_sync(self._h))

    def last_cycle_ms(self):
        v = C.c_double()
        self._check(_lib.lib().afmg_last_cycle_ms(self._h, C.byref(v)))
        return v.value

    def checksum(self, var):
        """(wrapping sum, xor) of the 64-bit patterns of `var` over the full records of this rank's boxes"""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(_lib.lib().afmg_checksum(self._h, var, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def slab_bytes(self):
        """(mapped, full) bytes of the cell-data slab per GPU of this handle (afmg_slab_bytes)"""
        a, b = (C.c_int64 * 8)(), (C.c_int64 * 8)()
        n = C.c_int32(0)
        self._check(_lib.lib().afmg_slab_bytes(self._h, 8, a, b, C.byref(n)))
        return np.array(a[:n.value]), np.array(b[:n.value])

    def set_mega(self, enabled=True, max_boxes=0):
        """persistent-kernel segments on / off (afmg_set_mega); max_boxes = 0: the default level-size limit"""
        self._check(_lib.lib().afmg_set_mega(self._h, int(enabled), int(max_boxes)))

    def mega_active(self):
        return int(_lib.lib().afmg_mega_active(self._h))

    def kernel_launches(self):
        return int(_lib.lib().afmg_kernel_launches(self._h))

    def cell_updates(self, highest_lvl=0, fmg=False):
        v = C.c_double()
        self._check(_lib.lib().afmg_cell_updates(self._h, highest_lvl, int(fmg), C.byref(v)))
        return v.value

    def set_profiling(self, on):
        self._check(_lib.lib().afmg_set_profiling(self._h, int(on)))

    def profile(self):
        cap = 64
        names = ((C.c_char * 32) * cap)()
        ms = (C.c_double * cap)()
        calls = (C.c_int64 * cap)()
        n = C.c_int32()
        self._check(_lib.lib().afmg_profile(self._h, cap, names, ms, calls, C.byref(n)))
        return {names[i].value.decode(): (ms[i], calls[i]) for i in range(n.value)}

    def slot_of_box(self, box_id):
        return int(_lib.lib().afmg_slot_of_box(self._h, int(box_id)))

    def owner_of_box(self, box_id):
        return int(_lib.lib().afmg_owner_of_box(self._h, int(box_id)))

    def owners(self, ids):
        f = _lib.lib().afmg_owner_of_box
        return np.fromiter((f(self._h, int(i)) for i in ids), dtype=np.int32, count=len(ids))

    def own_boxes(self, lvl):
        """number of boxes of level lvl this rank computes"""
        ids = self._tree.lvl_ids[lvl - 1]
        if self.comm is None or self.comm[1] == 1:
            return len(ids)
        return int(np.count_nonzero(self.owners(ids) == self.comm[0]))


def mg_compute_phi_gradient(tree: Tree, mg: mg_t, fac: float, with_norm: bool = True):
    """mg_compute_phi_gradient(tree, mg, i_fc, fac, i_norm) (afivo/src/m_af_multigrid.f90:1857-1898); the
    face-centred result is read with mg.get_fc, the norm with mg.get_cc(I_FLD, ids)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_compute_phi_gradient(mg._h, float(fac), int(with_norm)))


def mg_compute_field_norm(tree: Tree, mg: mg_t):
    """mg_compute_field_norm (afivo/src/m_af_multigrid.f90:2002-2020)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_compute_field_norm(mg._h))


def af_gc_tree(tree: Tree, mg: mg_t, var: int, corners: bool = True):
    """af_gc_tree(tree, [var], corners) (afivo/src/m_af_ghostcell.f90:25-46) for I_PHI or I_FLD."""
    mg._need_init()
    mg._check(_lib.lib().afmg_gc_tree(mg._h, int(var), int(corners)))


def field_from_potential(tree: Tree, mg: mg_t, fac: float = -1.0):
    """field_from_potential without dielectric (src/m_field.f90:531-548)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_field_from_potential(mg._h, float(fac)))


def photoi_helmh_bc(nb, coords):
    """photoi_helmh_bc (src/m_photoi_helmh.f90:210-228): Dirichlet 0 in the last dimension, Neumann 0 elsewhere."""
    if (nb - 1) // 2 == coords.shape[-1] - 1:
        return AF_BC_DIRICHLET, 0.0
    return AF_BC_NEUMANN, 0.0


def photoi_helmh_parameters(author: str = "Bourdon-3", frac_O2: float = 0.2, gas_pressure: float = 1.0, eta: float = 1.0,
                            lambdas=None, coeffs=None):
    """The Helmholtz expansion of the photoionization source the streamer code offers (photoi_helmh%author,
    src/m_photoi_helmh.f90:80-136): (lambdas [1/m], coeffs [1/m^2]) scaled by the O2 fraction and the pressure in bar
    exactly as there; "custom" takes lambdas / coeffs and scales by the pressure only."""
    if author != "custom" and frac_O2 <= 0.0:
        raise ValueError("Photoionization: no oxygen present")  # error stop in the reference
    if author == "Luque":
        if abs(eta - 1.0) > 0:
            raise ValueError("With Luque photoionization, photoi%eta should be 1.0")
        lam, cf, scale = [4425.38, 750.06], [337557.38, 19972.14], (frac_O2 / 0.2) * gas_pressure
    elif author == "Bourdon-2":
        lam, cf, scale = [7305.62, 44081.25], [11814508.38, 998607256.0], frac_O2 * gas_pressure
    elif author == "Bourdon-3":
        lam, cf, scale = [4147.85, 10950.93, 66755.67], [1117314.935, 28692377.5, 2748842283.0], frac_O2 * gas_pressure
    elif author == "custom":
        if lambdas is None or coeffs is None or len(lambdas) < 1 or len(lambdas) != len(coeffs):
            raise ValueError("Custom photoionization lambdas and coeffs missing.")
        lam, cf, scale = list(lambdas), list(coeffs), gas_pressure
    else:
        raise ValueError(f"Unknown photoi_helmh_author: {author}")
    return np.asarray(lam, float) * scale, np.asarray(cf, float) * scale ** 2


def photoi_helmh_initialize(tree: Tree, author: str = "Bourdon-3", *, length_unit: float = 1.0, **kw):
    """photoi_helmh_initialize (src/m_photoi_helmh.f90:138-156): one mg_t per mode with helmholtz_lambda =
    lambdas(n)**2, mg_prolong_linear and photoi_helmh_bc, initialised on `tree`.  length_unit = metres per unit of the
    tree's coordinates (lambdas are 1/m).  Returns (mg_helm, coeffs) for photoi_helmh_compute; device and the other
    mg_t options pass through **kw, the parameter-set options (frac_O2, gas_pressure, eta, lambdas, coeffs) too."""
    from . import workloads as W
    par = {k: kw.pop(k) for k in ("frac_O2", "gas_pressure", "eta", "lambdas", "coeffs") if k in kw}
    lam, cf = photoi_helmh_parameters(author, **par)
    lam, cf = lam * length_unit, cf * length_unit ** 2
    bc = W.bc_table(tree, photoi_helmh_bc)
    mg_helm = []
    for l in lam:
        m = mg_t(sides_bc=bc, helmholtz_lambda=float(l ** 2), prolongation_type=MG_PROLONG_LINEAR, **kw)
        mg_init(tree, m)
        mg_helm.append(m)
    return mg_helm, cf


def photoi_helmh_compute(tree: Tree, mg_helm, coeffs, max_fmg_cycles: int = 10, max_rel_residual: float = 1.0e-2):
    """photoi_helmh_compute (src/m_photoi_helmh.f90:162-204) on the device: mg_helm = the mg_t of every mode
    (helmholtz_lambda = lambdas(n)**2), rhs uploaded to mg_helm[0]; the source is read with
    mg_helm[0].get_cc(I_PHOTO, ids).  Returns (FMG cycles per mode, last residual max-norm per mode)."""
    n = len(mg_helm)
    for m in mg_helm:
        m._need_init()
    hs = (C.c_void_p * n)(*[m._h for m in mg_helm])
    cf = np.ascontiguousarray(coeffs, np.float64)
    assert cf.size == n
    ncyc = np.zeros(n, np.int32)
    res = np.zeros(n)
    mg_helm[0]._check(_lib.lib().afmg_helmholtz_compute(
        hs, n, cf.ctypes.data_as(C.POINTER(C.c_double)), int(max_fmg_cycles), float(max_rel_residual),
        ncyc.ctypes.data_as(C.POINTER(C.c_int32)), res.ctypes.data_as(C.POINTER(C.c_double))))
    return ncyc, res


def mg_from_dat(dat, phi="phi", rhs="rhs", eps=None, lsf="lsf", operator_key=1, prolongation_key=2, **opts):
    """Set a solver up from an afivo .dat file (datfile.DatFile) alone: topology, the boundary conditions
    stored in the boxes for `phi`, phi / rhs (/ eps) data, the stored operator / prolongation stencils (rebuilt with
    the library's builders from eps and the distance stencils when the file holds none) and level-set distance
    stencils.  Returns (tree, mg) ready for mg_fas_fmg / mg_fas_vcycle."""
    tree = dat.tree
    mg = mg_t(sides_bc=dat.bc_table(phi), **opts)
    mg_init(tree, mg)
    entries = dat.stencil_entries(operator_key, prolongation_key, mg.operator_mask)
    if not entries:  # a file without stored operator stencils: rebuild them from eps / the distance stencils
        from .datfile import rebuild_stencil_entries
        entries = rebuild_stencil_entries(dat, eps, mg.operator_mask, mg.prolongation_type)
    if entries:
        mg.set_stencils(entries)
    ids = dat.ids_in_use()
    mg.set_cc(I_PHI, ids, dat.cc_of(phi, ids))
    mg.set_cc(I_RHS, ids, dat.cc_of(rhs, ids))
    if eps is not None:
        mg.set_cc(I_EPS, ids, dat.cc_of(eps, ids))
    ld = dat.lsf_distances(lsf)
    if ld is not None:
        mg.set_lsf_distances_sparse(*ld)
    return tree, mg


def comm_from_torch(group=None):
    """(rank, world, allgather) on top of an initialised torch.distributed process group (nccl or gloo)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def allgather(raw: bytes):
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        mine = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(out, mine, group=group)
        return [bytes(t.cpu().numpy().tobytes()) for t in out]

    return rank, world, allgather


def partition(n_ranks: int, lvl_counts, min_split_boxes: int = 0):
    """afmg_partition_min: per level, the first position (Morton order) of every rank; shape (L, n_ranks + 1).
    Levels with fewer than min_split_boxes boxes stay on rank 0 (0: plain afmg_partition)."""
    counts = np.ascontiguousarray(lvl_counts, np.int32)
    cuts = np.zeros((len(counts), n_ranks + 1), np.int32)
    rc = _lib.lib().afmg_partition_min(n_ranks, len(counts), counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                       int(min_split_boxes), cuts.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise AfmgError(rc, "afmg_partition: invalid arguments")
    return cuts


def _opts_from(tree: Tree, mg: mg_t) -> Opts:
    o = Opts()
    o.ndim, o.n_cell, o.coord_t = tree.ndim, tree.nc, tree.coord_t
    o.n_cycle_down, o.n_cycle_up = mg.n_cycle_down, mg.n_cycle_up
    o.use_corners, o.subtract_mean = int(mg.use_corners), int(mg.subtract_mean)
    o.prolongation_type, o.operator_mask = mg.prolongation_type, mg.operator_mask
    o.has_eps, o.device = 0, mg.device
    o.n_gpus = mg.n_gpus
    o.helmholtz_lambda, o.lsf_boundary_value = mg.helmholtz_lambda, mg.lsf_boundary_value
    for d in range(3):
        o.coarse_grid_size[d] = int(tree.coarse_grid_size[d]) if d < tree.ndim else 1
        o.periodic[d] = int(tree.periodic[d]) if d < tree.ndim else 0
        o.dr_base[d] = float(tree.dr_base[d]) if d < tree.ndim else 0.0
        o.r_base[d] = float(tree.r_base[d]) if d < tree.ndim else 0.0
    return o


def _tree_desc(tree: Tree):
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    keep = dict(
        counts=i32([len(a) for a in tree.lvl_ids]), ids=i32(np.concatenate(tree.lvl_ids)), lvl=i32(tree.lvl),
        ix=i32(tree.ix), parent=i32(tree.parent), children=i32(tree.children), neighbors=i32(tree.neighbors),
        nmat=i32(tree.neighbor_mat), r_min=np.ascontiguousarray(tree.r_min, np.float64))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    td = TreeDesc(tree.highest_lvl, tree.highest_id, ip(keep["counts"]), ip(keep["ids"]), ip(keep["lvl"]),
                  ip(keep["ix"]), ip(keep["parent"]), ip(keep["children"]), ip(keep["neighbors"]), ip(keep["nmat"]),
                  keep["r_min"].ctypes.data_as(C.POINTER(C.c_double)))
    return td, keep


def mg_init(tree: Tree, mg: mg_t):
    """mg_init (afivo/src/m_af_multigrid.f90:43-109)."""
    if mg.sides_bc is None:
        raise AfmgError(-1, "mg_init: sides_bc not set")  # :50-51 error stop
    L = _lib.lib()
    h = C.c_void_p()
    o = _opts_from(tree, mg)
    rc = L.afmg_create(C.byref(h), C.byref(o))
    if rc != 0:
        raise AfmgError(rc, (L.afmg_last_error(None) or b"").decode())
    mg._h, mg._tree = h, tree
    mg.initialized = True
    if mg.comm is not None:
        mg._check(L.afmg_comm_init(h, int(mg.comm[1]), int(mg.comm[0])))
    mg_set_tree(tree, mg)


def mg_set_tree(tree: Tree, mg: mg_t):
    """Forward a (new) topology, e.g. after af_adjust_refinement, and re-evaluate sides_bc."""
    mg._need_init()
    td, keep = _tree_desc(tree)
    mg._check(_lib.lib().afmg_set_tree(mg._h, C.byref(td)))
    mg._tree = tree
    if mg.comm is not None and mg.comm[1] > 1:
        # exchange the CUDA IPC handles of the ranks' device arrays (plumbing; the data path itself is
        # peer-memory loads / stores inside the kernels)
        blob = C.create_string_buffer(_lib.AFMG_COMM_BLOB_BYTES)
        mg._check(_lib.lib().afmg_comm_export(mg._h, blob))
        blobs = mg.comm[2](blob.raw)
        allb = C.create_string_buffer(b"".join(blobs), _lib.AFMG_COMM_BLOB_BYTES * len(blobs))
        mg._check(_lib.lib().afmg_comm_connect(mg._h, allb))
    bc = mg.sides_bc if isinstance(mg.sides_bc, BCTable) else bc_table(tree, mg.sides_bc)
    mg.set_bc(bc)


def mg_destroy(mg: mg_t):
    """mg_destroy (afivo/src/m_af_multigrid.f90:111-115)."""
    if mg._h is not None:
        _lib.lib().afmg_destroy(mg._h)
    mg._h, mg.initialized = None, False


def mg_use(tree: Tree, mg: mg_t):
    """mg_use (afivo/src/m_af_multigrid.f90:118-126): "make sure box tags and operators are set".  The library keeps
    the operators of a handle current itself (afmg_set_tree / afmg_set_stencils / afmg_update_operator_stencil), and
    every mg_t has its own handle, so nothing is switched here; what remains is the reference's check."""
    if not mg.initialized:
        raise _lib.AfmgError(-4, "mg%initialized is false")  # error stop in the reference


def mg_fas_fmg(tree: Tree, mg: mg_t, set_residual: bool, have_guess: bool):
    """mg_fas_fmg (afivo/src/m_af_multigrid.f90:137-180)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_fas_fmg(mg._h, int(set_residual), int(have_guess)))


def mg_fas_vcycle(tree: Tree, mg: mg_t, set_residual: bool, highest_lvl: Optional[int] = None,
                  standalone: bool = True):
    """mg_fas_vcycle (afivo/src/m_af_multigrid.f90:185-264)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_fas_vcycle(mg._h, int(set_residual), int(highest_lvl or 0), int(standalone)))


def field_solve(tree: Tree, mg: mg_t, have_guess: bool, residual_threshold: float, *, max_residual: float = 1e8,
                max_initial_iterations: int = 100, num_vcycles: int = 2):
    """The FMG / V-cycle convergence loop of field_compute (src/m_field.f90:491-524) run next to the
    device; returns (residuals, n_fmg, n_vcycles)."""
    mg._need_init()
    res = np.zeros(max_initial_iterations + num_vcycles)
    n_fmg, n_vc = C.c_int32(0), C.c_int32(0)
    mg._check(_lib.lib().afmg_field_solve(mg._h, int(have_guess), float(residual_threshold), float(max_residual),
                                          int(max_initial_iterations), int(num_vcycles),
                                          res.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n_fmg), C.byref(n_vc)))
    return res[: n_fmg.value + n_vc.value], n_fmg.value, n_vc.value


def field_residual_threshold(tree: Tree, max_rhs: float, current_voltage: float, *, use_electrode: bool = False,
                             multigrid_max_rel_residual: float = 1.0e-4) -> float:
    """The convergence threshold of field_compute (src/m_field.f90:467-480): max(min_residual = 1e-6, max|rhs| *
    ST_multigrid_max_rel_residual, conv_fac * |voltage| / (domain_len(NDIM) * af_min_dr(tree))) with conv_fac = 1e-8
    with an electrode and 1e-10 without -- the last term estimates the round-off of (phi / L * dx) / dx^2."""
    nd = tree.ndim
    conv_fac = 1.0e-8 if use_electrode else 1.0e-10
    domain_len = float(tree.coarse_grid_size[nd - 1] * tree.dr_base[nd - 1])
    min_dr = float(np.min(np.asarray(tree.dr_base)[:nd]) * 0.5 ** (tree.highest_lvl - 1))  # af_min_dr
    return max(1.0e-6, max_rhs * multigrid_max_rel_residual, conv_fac * abs(current_voltage) / (domain_len * min_dr))


def field_compute(tree: Tree, mg: mg_t, current_voltage: float, have_guess: bool, *, use_electrode: bool = False,
                  electrode_grounded: bool = False, multigrid_max_rel_residual: float = 1.0e-4,
                  multigrid_num_vcycles: int = 2):
    """field_compute (src/m_field.f90:448-528) after field_set_rhs / field_set_voltage: the right-hand side is on the
    device (mg.set_cc_interior(I_RHS, leaves, ...)) and the boundary conditions carry the voltage.  Threshold from
    max|rhs|, the electrode potential (mg%lsf_boundary_value = 0 if grounded else the voltage, :482-487), the FMG loop
    when there is no guess and the V-cycles (afmg_field_solve), then field_from_potential.  Returns
    (residuals, n_fmg, n_vcycles); a non-converging start raises AFMG_ERR_NOT_CONVERGED ("No convergence in initial
    field computation")."""
    mg._need_init()
    max_rhs = af_tree_maxabs_cc(tree, mg, I_RHS)
    thr = field_residual_threshold(tree, max_rhs, current_voltage, use_electrode=use_electrode,
                                   multigrid_max_rel_residual=multigrid_max_rel_residual)
    if use_electrode:
        mg.set_lsf_boundary_value(0.0 if electrode_grounded else current_voltage)
    out = field_solve(tree, mg, have_guess, thr, num_vcycles=multigrid_num_vcycles)
    field_from_potential(tree, mg, -1.0)
    return out


def mg_update_operator_stencil(tree: Tree, mg: mg_t):
    """mg_update_operator_stencil (afivo/src/m_af_multigrid.f90:1188-1214)."""
    mg._need_init()
    mg._check(_lib.lib().afmg_set_helmholtz_lambda(mg._h, float(mg.helmholtz_lambda)))
    mg._check(_lib.lib().afmg_update_operator_stencil(mg._h))


def mg_set_operators_tree(tree: Tree, mg: mg_t, *, eps_cc=None, lsf=None, lsf_options=None,
                          lsf_use_custom_prolongation: bool = False):
    """What mg_init does for a tree with mg_i_eps / mg_i_lsf set (mg_set_operators_tree,
    afivo/src/m_af_multigrid.f90:1216-1225 -> mg_set_box_tag :1100-1145, mg_store_operator_stencil :823-859,
    mg_store_prolongation_stencil :862-903): tag every box, build the explicit operator / prolongation stencils with
    the library's host-side builders (stencils.build_stencils) and ship them, the permittivity (AFMG_EPS, for the
    eps-weighted field of mg_box_lpl_gradient) and the level-set distances (for mg_box_lpllsf_gradient).

    eps_cc: (highest_id + 1, nc+2, ...) permittivity indexed by box id, ghost cells filled; lsf: mg%lsf as a function
    of one point.  Returns (entries, lsf_data) for inspection."""
    from . import stencils as S
    mg._need_init()
    entries, data = S.build_stencils(tree, eps_cc=eps_cc, lsf=lsf, lsf_options=lsf_options,
                                     operator_mask=mg.operator_mask, prolongation_type=mg.prolongation_type,
                                     lsf_use_custom_prolongation=lsf_use_custom_prolongation)
    mg.set_stencils(entries)
    if eps_cc is not None:
        ids = np.concatenate(tree.lvl_ids).astype(np.int32)
        mg.set_cc(I_EPS, ids, np.ascontiguousarray(eps_cc)[ids])
    if data is not None:
        mg.set_lsf_distances(data.ids, data.dd, data.lsf_cells)
    return entries, data


def af_tree_maxabs_cc(tree: Tree, mg: mg_t, iv: int) -> float:
    """af_tree_maxabs_cc (afivo/src/m_af_utils.f90:773-785): max |cc| over leaves, interior cells."""
    v = C.c_double()
    mg._check(_lib.lib().afmg_max_abs(mg._h, iv, C.byref(v)))
    return v.value


def af_tree_sum_cc(tree: Tree, mg: mg_t, iv: int) -> float:
    """af_tree_sum_cc (afivo/src/m_af_utils.f90:966-1027)."""
    v = C.c_double()
    mg._check(_lib.lib().afmg_tree_sum(mg._h, iv, C.byref(v)))
    return v.value
