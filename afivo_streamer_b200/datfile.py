"""afivo ``.dat`` tree files, version 3: reader and writer of the binary stream format of
``af_write_tree`` / ``af_read_tree`` (afivo/src/m_af_output.f90:41-192, :197-373, version constant :10).

A ``.dat`` file holds everything the multigrid path needs to re-run a solve of a real simulation without any
Fortran at run time: the tree topology (``lvls(:)%ids``, per box ``lvl, tag, ix, parent, children, neighbors,
neighbor_mat, dr, r_min``), the cell-centred variables (phi, rhs, eps, lsf, ...), the boundary conditions the
last ghost-cell fill stored in the boxes (``bc_type``, ``bc_val``; m_af_ghostcell.f90:108-113) and the stored
stencils (operator, prolongation, level-set distances; ``stencil_t`` m_af_types.f90:260-282).

Format notes (gfortran, ``access='stream'``, no record markers): default integers and logicals are 4 bytes,
reals 8 bytes, names ``character(len=af_nlen=20)`` in arrays of ``af_max_num_vars = 1024``
(m_af_types.f90:20, 72, 341-356).  NDIM is a compile-time constant of the writer and is NOT in the file: pass
``ndim`` or let the reader try 2 and 3 (only one of them consumes the file consistently).

The reference ships no ``.dat`` fixture, so this module is checked by round trips through its own writer (which
follows the same write statements) and by solving from a written file; "parity unpinned" against genuine files.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

import numpy as np

from .tree import Tree
from .workloads import BCTable

AF_DAT_FILE_VERSION = 3
AF_MAX_NUM_VARS = 1024
AF_NLEN = 20
# stencil shapes (afivo/src/m_af_stencil.f90:19-36) and types (m_af_types.f90:247-255)
STENCIL_357, STENCIL_P234, STENCIL_P248, STENCIL_246, STENCIL_MASK = 1, 2, 3, 4, 5
STENCIL_CONSTANT, STENCIL_VARIABLE, STENCIL_SPARSE = 1, 2, 3
MG_LSF_DISTANCE_KEY, MG_LSF_MASK_KEY = 31, 32  # m_af_types.f90:538-540


@dataclasses.dataclass
class DatStencil:
    """stencil_t as stored in the file (m_af_output.f90:134-182)."""
    key: int
    shape: int
    stype: int
    cylindrical_gradient: bool = False
    c: Optional[np.ndarray] = None              # (n_coeff,)
    v: Optional[np.ndarray] = None              # (cells, n_coeff): v(n, i, j, k) with n fastest
    f: Optional[np.ndarray] = None              # (cells,)
    bc_correction: Optional[np.ndarray] = None  # (cells,)
    sparse_ix: Optional[np.ndarray] = None      # (k, NDIM)
    sparse_v: Optional[np.ndarray] = None       # (k, m)


@dataclasses.dataclass
class DatBC:
    """Boundary-condition storage of one box (af_init_box, m_af_core.f90:558-577)."""
    bc_index_to_nb: np.ndarray  # (n_bc,)
    nb_to_bc_index: np.ndarray  # (2*NDIM,)
    bc_type: np.ndarray         # (n_bc, n_var_cell)
    bc_val: np.ndarray          # (n_bc, n_var_cell, nc^(D-1))
    bc_coords: np.ndarray       # (n_bc, nc^(D-1), NDIM)


@dataclasses.dataclass
class DatFile:
    ndim: int
    tree: Tree
    ready: bool
    box_limit: int
    cc_names: List[str]
    fc_names: List[str]
    cc_num_copies: np.ndarray
    cc_write_output: np.ndarray
    cc_write_binary: np.ndarray
    fc_write_binary: np.ndarray
    removed_ids: np.ndarray
    lvl_leaves: List[np.ndarray]
    lvl_parents: List[np.ndarray]
    in_use: np.ndarray                 # (n+1,) bool
    tag: np.ndarray                    # (n+1,)
    cc: Dict[int, np.ndarray]          # variable index (1-based, as in the reference) -> (n+1, (nc+2)^D)
    fc: Dict[int, np.ndarray]          # face variable index -> (n+1, D*(nc+1)^D)
    bc: Dict[int, DatBC]               # box id -> boundary-condition storage
    stencils: Dict[int, List[DatStencil]]
    other_data: Optional[bytes] = None  # whatever write_other_data appended (opaque)

    # ---- lookups ----------------------------------------------------------------------------
    def var_index(self, name: str) -> int:
        """1-based index of a cell-centred variable (af_find_cc_variable)."""
        try:
            return self.cc_names.index(name) + 1
        except ValueError:
            raise KeyError(f"no cell-centred variable named {name!r}; have {self.cc_names}") from None

    def ids_in_use(self) -> np.ndarray:
        return np.concatenate(self.tree.lvl_ids).astype(np.int32)

    def cc_of(self, name_or_index, ids) -> np.ndarray:
        iv = name_or_index if isinstance(name_or_index, int) else self.var_index(name_or_index)
        if iv not in self.cc:
            raise KeyError(f"variable {name_or_index!r} was not written (cc_write_binary = F)")
        return self.cc[iv][np.asarray(ids)]

    def bc_table(self, name_or_index) -> BCTable:
        """The boundary conditions the boxes stored for a variable, one row per physical face, in the row
        format of afmg_set_bc."""
        iv = name_or_index if isinstance(name_or_index, int) else self.var_index(name_or_index)
        ids, nbs, types, vals = [], [], [], []
        for bid in self.ids_in_use():
            b = self.bc.get(int(bid))
            if b is None:
                continue
            for q, nb in enumerate(b.bc_index_to_nb):
                ids.append(bid)
                nbs.append(nb)
                types.append(b.bc_type[q, iv - 1])
                vals.append(b.bc_val[q, iv - 1])
        nface = self.tree.nc ** (self.ndim - 1)
        return BCTable(np.asarray(ids, np.int32), np.asarray(nbs, np.int32), np.asarray(types, np.int32),
                       np.asarray(vals, np.float64).reshape(len(ids), nface))

    def stencil_entries(self, operator_key: int = 1, prolongation_key: int = 2, operator_mask: int = -1,
                        skip_plain: bool = True):
        """Entries for mg_t.set_stencils from the stored stencils.  The first mg_t initialised on a tree gets
        operator_key 1 and prolongation_key 2 (mg_init, m_af_multigrid.f90:63-72).  With skip_plain, boxes
        tagged mg_normal_box whose stored stencils are constant (mg_box_lpl_stencil and the constant
        prolongation: what the library derives itself from dr, helmholtz_lambda and prolongation_type) get no
        entry and run through the fast kernels."""
        out = []
        for bid in self.ids_in_use():
            e = dict(box_id=int(bid), tag=int(self.tag[bid]) if self.tag[bid] >= 0 else 0)
            for st in self.stencils.get(int(bid), []):
                if st.key == operator_key and st.shape == STENCIL_357:
                    e["op"] = (st.stype, st.c if st.stype == STENCIL_CONSTANT else st.v)
                    e["cyl"] = st.cylindrical_gradient
                    if st.f is not None:
                        e["f"] = st.f
                elif st.key == prolongation_key and st.shape in (STENCIL_P234, STENCIL_P248):
                    e["prolong"] = (st.stype, st.shape, st.c if st.stype == STENCIL_CONSTANT else st.v)
            plain = (e["tag"] & operator_mask) == 0 and e.get("op", (STENCIL_CONSTANT,))[0] == STENCIL_CONSTANT \
                and e.get("prolong", (STENCIL_CONSTANT,))[0] == STENCIL_CONSTANT and "f" not in e
            if skip_plain and plain:
                continue
            if "op" in e or "prolong" in e or e["tag"]:
                out.append(e)
        return out

    def lsf_distances(self, lsf_name: Optional[str] = "lsf"):
        """(ids, n_entries, cell_ix, dd, lsf) of the mg_lsf_distance_key stencils, the arguments of
        afmg_set_lsf_distances.  lsf = the stored level-set variable at those cells, None if not written."""
        ids, n_ent, cells, dd, lv = [], [], [], [], []
        iv = None
        if lsf_name is not None and lsf_name in self.cc_names and self.var_index(lsf_name) in self.cc:
            iv = self.var_index(lsf_name)
        n2 = self.tree.nc + 2
        for bid in self.ids_in_use():
            for st in self.stencils.get(int(bid), []):
                if st.key == MG_LSF_DISTANCE_KEY and st.sparse_ix is not None:
                    ids.append(bid)
                    n_ent.append(len(st.sparse_ix))
                    cells.append(st.sparse_ix)
                    dd.append(st.sparse_v)
                    if iv is not None:
                        lin = np.zeros(len(st.sparse_ix), np.int64)
                        for d in reversed(range(self.ndim)):
                            lin = lin * n2 + st.sparse_ix[:, d]
                        lv.append(self.cc[iv][bid][lin])
        if not ids:
            return None
        return (np.asarray(ids, np.int32), np.asarray(n_ent, np.int32), np.concatenate(cells).astype(np.int32),
                np.concatenate(dd).astype(np.float64), np.concatenate(lv) if iv is not None else None)


def rebuild_stencil_entries(dat: "DatFile", eps: Optional[str] = None, operator_mask: int = -1,
                            prolongation_type: int = 19):
    """Entries for mg_t.set_stencils rebuilt from the DATA of a .dat file instead of its stored operator /
    prolongation stencils: the permittivity variable `eps` and the stored level-set distance stencils
    (mg_lsf_distance_key) go through the library's host-side builders (stencils.build_stencils = the reference's
    mg_set_operators_lvl).  For files written without operator stencils (before mg_init, or by a tool that strips
    them); for a file that has them the result equals DatFile.stencil_entries()."""
    from . import stencils as S
    t = dat.tree
    nd, nc = t.ndim, t.nc
    ncell = nc ** nd
    eps_cc = None
    if eps is not None:
        ids = dat.ids_in_use()
        eps_cc = np.ones((t.highest_id + 1,) + (nc + 2,) * nd)
        eps_cc[ids] = dat.cc_of(eps, ids).reshape((len(ids),) + (nc + 2,) * nd)
    lsf_data = None
    ld = dat.lsf_distances(None)
    if ld is not None:
        ids, n_ent, cells, dd, _ = ld
        dense = np.ones((len(ids), ncell, 2 * nd))
        pos = 0
        for b, n in enumerate(n_ent):
            cx = cells[pos:pos + n] - 1
            lin = np.zeros(n, np.int64)
            for d in reversed(range(nd)):
                lin = lin * nc + cx[:, d]
            dense[b, lin] = dd[pos:pos + n]
            pos += n
        lsf_data = S.LsfData(np.asarray(ids, np.int32), dense, np.zeros((len(ids), ncell)), {}, None)
    if eps_cc is None and lsf_data is None:
        return []
    entries, _ = S.build_stencils(t, eps_cc=eps_cc, lsf_data=lsf_data, operator_mask=operator_mask,
                                  prolongation_type=prolongation_type)
    return entries


class _Cursor:
    def __init__(self, buf: bytes):
        self.buf = memoryview(buf)
        self.pos = 0

    def take(self, dtype, n):
        nbytes = np.dtype(dtype).itemsize * n
        if n < 0 or self.pos + nbytes > len(self.buf):
            raise ValueError("af_read_tree: unexpected end of file (wrong NDIM?)")
        a = np.frombuffer(self.buf, dtype=dtype, count=n, offset=self.pos)
        self.pos += nbytes
        return a

    def i4(self, n=None):
        return int(self.take("<i4", 1)[0]) if n is None else self.take("<i4", n).copy()

    def f8(self, n):
        return self.take("<f8", n).copy()

    def logical(self, n=None):
        return bool(self.take("<i4", 1)[0]) if n is None else self.take("<i4", n) != 0

    def names(self, n):
        raw = bytes(self.take("S1", n * AF_NLEN))
        return [raw[i * AF_NLEN:(i + 1) * AF_NLEN].decode("ascii", "replace").rstrip() for i in range(n)]


def read_tree(path: str, ndim: Optional[int] = None) -> DatFile:
    """af_read_tree (afivo/src/m_af_output.f90:197-373)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if ndim is not None:
        return _parse(buf, ndim)
    errors = []
    for nd in (3, 2):
        try:
            return _parse(buf, nd)
        except ValueError as e:
            errors.append(f"NDIM={nd}: {e}")
    raise ValueError("af_read_tree: the file parses with neither NDIM; " + "; ".join(errors))


def _parse(buf: bytes, nd: int) -> DatFile:
    c = _Cursor(buf)
    version = c.i4()
    if version != AF_DAT_FILE_VERSION:
        raise ValueError(f"af_read_tree: incompatible file versions (read {version}, required {AF_DAT_FILE_VERSION})")
    ready = c.logical()
    box_limit, highest_lvl, highest_id, nc, n_var_cell, n_var_face, coord_t = (c.i4() for _ in range(7))
    if not (0 < highest_lvl <= 30 and 0 < highest_id <= box_limit and 2 <= nc <= 1024 and nc % 2 == 0
            and 0 <= n_var_cell <= AF_MAX_NUM_VARS and 0 <= n_var_face <= AF_MAX_NUM_VARS):
        raise ValueError("af_read_tree: implausible header")
    coarse_grid_size = c.i4(nd)
    periodic = c.logical(nd).copy()
    r_base = c.f8(nd)
    dr_base = c.f8(nd)
    if np.any(coarse_grid_size < nc) or np.any(coarse_grid_size % nc) or np.any(dr_base <= 0):
        raise ValueError("af_read_tree: implausible coarse grid (wrong NDIM?)")
    cc_names = c.names(AF_MAX_NUM_VARS)[:n_var_cell]
    fc_names = c.names(AF_MAX_NUM_VARS)[:n_var_face]
    cc_num_copies = c.i4(AF_MAX_NUM_VARS)[:n_var_cell]
    cc_write_output = c.logical(AF_MAX_NUM_VARS)[:n_var_cell].copy()
    cc_write_binary = c.logical(AF_MAX_NUM_VARS)[:n_var_cell].copy()
    fc_write_binary = c.logical(AF_MAX_NUM_VARS)[:n_var_face].copy()
    n_removed = c.i4()
    removed_ids = c.i4(n_removed)
    lvl_ids, lvl_leaves, lvl_parents = [], [], []
    for _ in range(highest_lvl):
        lvl_ids.append(c.i4(c.i4()))
        lvl_leaves.append(c.i4(c.i4()))
        lvl_parents.append(c.i4(c.i4()))
    n = highest_id
    nch, nnb, nm = 1 << nd, 2 * nd, 3 ** nd
    box_len, fc_len, nface = (nc + 2) ** nd, nd * (nc + 1) ** nd, nc ** (nd - 1)
    ncell = nc ** nd
    in_use = np.zeros(n + 1, bool)
    lvl = np.zeros(n + 1, np.int32)
    tag = np.zeros(n + 1, np.int32)
    ix = np.zeros((n + 1, nd), np.int32)
    parent = np.zeros(n + 1, np.int32)
    children = np.zeros((n + 1, nch), np.int32)
    neighbors = np.zeros((n + 1, nnb), np.int32)
    neighbor_mat = np.zeros((n + 1, nm), np.int32)
    dr = np.zeros((n + 1, nd))
    r_min = np.zeros((n + 1, nd))
    cc = {iv + 1: np.zeros((n + 1, box_len)) for iv in range(n_var_cell) if cc_write_binary[iv]}
    fc = {iv + 1: np.zeros((n + 1, fc_len)) for iv in range(n_var_face) if fc_write_binary[iv]}
    bcs: Dict[int, DatBC] = {}
    stencils: Dict[int, List[DatStencil]] = {}
    for bid in range(1, n + 1):
        in_use[bid] = c.logical()
        if not in_use[bid]:
            continue
        b_nc, n_bc, n_st = c.i4(), c.i4(), c.i4()
        if b_nc != nc or not (0 <= n_bc <= nnb) or n_st < 0:
            raise ValueError(f"af_read_tree: implausible box record {bid} (wrong NDIM?)")
        lvl[bid] = c.i4()
        tag[bid] = c.i4()
        ix[bid] = c.i4(nd)
        parent[bid] = c.i4()
        children[bid] = c.i4(nch)
        neighbors[bid] = c.i4(nnb)
        neighbor_mat[bid] = c.i4(nm)
        dr[bid] = c.f8(nd)
        r_min[bid] = c.f8(nd)
        c.i4()  # box%coord_t
        for iv in cc:
            cc[iv][bid] = c.f8(box_len)
        for iv in fc:
            fc[iv][bid] = c.f8(fc_len)
        if n_bc > 0:
            i2n = c.i4(n_bc)
            n2i = c.i4(nnb)
            bt = c.i4(n_var_cell * n_bc).reshape(n_bc, n_var_cell)
            bv = c.f8(nface * n_var_cell * n_bc).reshape(n_bc, n_var_cell, nface)
            bco = c.f8(nd * nface * n_bc).reshape(n_bc, nface, nd)
            bcs[bid] = DatBC(i2n, n2i, bt, bv, bco)
        lst = []
        for _ in range(n_st):
            st = DatStencil(key=c.i4(), shape=c.i4(), stype=c.i4(), cylindrical_gradient=c.logical())
            k = c.i4()
            if k > 0:
                st.c = c.f8(k)
            k = c.i4()
            if k > 0:
                st.v = c.f8(k * ncell).reshape(ncell, k)
            if c.i4() > 0:
                st.f = c.f8(ncell)
            if c.i4() > 0:
                st.bc_correction = c.f8(ncell)
            k = c.i4()
            if k > 0:
                st.sparse_ix = c.i4(nd * k).reshape(k, nd)
            m = c.i4()
            if k > 0 and m > 0:
                st.sparse_v = c.f8(m * k).reshape(k, m)
            lst.append(st)
        if lst:
            stencils[bid] = lst
    other_present = c.logical()
    other = bytes(c.buf[c.pos:]) if other_present else None
    if not other_present and c.pos != len(buf):
        raise ValueError("af_read_tree: trailing bytes (wrong NDIM?)")
    tree = Tree(ndim=nd, nc=nc, coord_t=coord_t, coarse_grid_size=coarse_grid_size.astype(np.int64), periodic=periodic,
                r_base=r_base, dr_base=dr_base, highest_lvl=highest_lvl, highest_id=highest_id,
                lvl_ids=[a.astype(np.int64) for a in lvl_ids], lvl=lvl, ix=ix, parent=parent, children=children,
                neighbors=neighbors, neighbor_mat=neighbor_mat, r_min=r_min, dr=dr)
    return DatFile(ndim=nd, tree=tree, ready=ready, box_limit=box_limit, cc_names=cc_names, fc_names=fc_names,
                   cc_num_copies=cc_num_copies, cc_write_output=cc_write_output, cc_write_binary=cc_write_binary,
                   fc_write_binary=fc_write_binary, removed_ids=removed_ids, lvl_leaves=lvl_leaves,
                   lvl_parents=lvl_parents, in_use=in_use, tag=tag, cc=cc, fc=fc, bc=bcs, stencils=stencils,
                   other_data=other)


def write_tree(path: str, d: DatFile) -> None:
    """af_write_tree (afivo/src/m_af_output.f90:41-192): same statements, same order."""
    t = d.tree
    nd, nc = d.ndim, t.nc
    out = []
    i4 = lambda *v: out.append(np.asarray(v, "<i4").tobytes())
    ia = lambda a: out.append(np.ascontiguousarray(a, "<i4").tobytes())
    fa = lambda a: out.append(np.ascontiguousarray(a, "<f8").tobytes())
    la = lambda a: out.append(np.asarray(a, bool).astype("<i4").tobytes())

    def names(lst):
        raw = b"".join(s.encode("ascii")[:AF_NLEN].ljust(AF_NLEN) for s in lst)
        out.append(raw.ljust(AF_MAX_NUM_VARS * AF_NLEN))

    def padded(a, fill, dtype):
        full = np.full(AF_MAX_NUM_VARS, fill, dtype)
        full[:len(a)] = a
        return full

    n_var_cell, n_var_face = len(d.cc_names), len(d.fc_names)
    i4(AF_DAT_FILE_VERSION)
    la([d.ready])
    i4(d.box_limit, t.highest_lvl, t.highest_id, nc, n_var_cell, n_var_face, t.coord_t)
    ia(t.coarse_grid_size[:nd])
    la(t.periodic[:nd])
    fa(t.r_base[:nd])
    fa(t.dr_base[:nd])
    names(d.cc_names)
    names(d.fc_names)
    ia(padded(d.cc_num_copies, 1, "<i4"))
    la(padded(d.cc_write_output, True, bool))
    la(padded(d.cc_write_binary, True, bool))
    la(padded(d.fc_write_binary, True, bool))
    i4(len(d.removed_ids))
    ia(d.removed_ids)
    for l in range(t.highest_lvl):
        for lst in (t.lvl_ids[l], d.lvl_leaves[l], d.lvl_parents[l]):
            i4(len(lst))
            ia(lst)
    for bid in range(1, t.highest_id + 1):
        la([d.in_use[bid]])
        if not d.in_use[bid]:
            continue
        b = d.bc.get(bid)
        sts = d.stencils.get(bid, [])
        i4(nc, 0 if b is None else len(b.bc_index_to_nb), len(sts), t.lvl[bid], d.tag[bid])
        ia(t.ix[bid])
        i4(t.parent[bid])
        ia(t.children[bid])
        ia(t.neighbors[bid])
        ia(t.neighbor_mat[bid])
        fa(t.dr[bid])
        fa(t.r_min[bid])
        i4(t.coord_t)
        for iv in sorted(d.cc):
            fa(d.cc[iv][bid])
        for iv in sorted(d.fc):
            fa(d.fc[iv][bid])
        if b is not None and len(b.bc_index_to_nb) > 0:
            ia(b.bc_index_to_nb)
            ia(b.nb_to_bc_index)
            ia(b.bc_type)
            fa(b.bc_val)
            fa(b.bc_coords)
        for st in sts:
            i4(st.key, st.shape, st.stype)
            la([st.cylindrical_gradient])
            if st.c is not None:
                i4(len(st.c))
                fa(st.c)
            else:
                i4(0)
            if st.v is not None:
                i4(st.v.shape[1])
                fa(st.v)
            else:
                i4(0)
            for a in (st.f, st.bc_correction):
                if a is not None:
                    i4(1)
                    fa(a)
                else:
                    i4(0)
            if st.sparse_ix is not None:
                i4(len(st.sparse_ix))
                ia(st.sparse_ix)
            else:
                i4(0)
            if st.sparse_v is not None:
                i4(st.sparse_v.shape[1])
                fa(st.sparse_v)
            else:
                i4(0)
    la([d.other_data is not None])
    if d.other_data is not None:
        out.append(d.other_data)
    with open(path, "wb") as fh:
        fh.write(b"".join(out))
