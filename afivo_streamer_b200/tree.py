"""Flat description of an afivo tree (``af_t`` / ``box_t``) and a synthetic tree builder.

The multigrid path only *reads* the tree topology; in a drop-in setting the Fortran
side owns the tree and the shim forwards these same flat arrays (see
``include/afmg.h``: ``afmg_tree``).  For tests and benchmarks there is no Fortran
side, so this module builds 2:1-balanced trees that follow the reference's
conventions exactly:

* ids are 1-based, ``af_no_box = 0``, ``af_phys_boundary = -1``
  (afivo/src/m_af_types.f90:38-41);
* level-1 boxes are numbered ``i + (j-1)*nx + (k-1)*nx*ny``
  (afivo/src/m_af_core.f90:436-501);
* a child ``c`` (1..2^D) of a box with spatial index ``ix`` has index
  ``2*ix - 1 + af_child_dix(:, c)`` (afivo/src/m_af_core.f90:1187-1232,
  table afivo/src/m_af_types.f90:172-174);
* ``lvls(l+1)%ids`` is the concatenation, over ``lvls(l)%parents`` in list order, of
  ``children(1:2^D)`` (afivo/src/m_af_core.f90:1237-1254), parents/leaves keep the
  relative order of ids (:504-535);
* ``neighbors(nb)`` with nb = lowx, highx, lowy, highy, lowz, highz and
  ``neighbor_mat(-1:1, ...)`` hold a same-level id, 0 where the neighbouring region is
  only covered by a coarser box, and -1 outside the domain (:595-661);
* 2:1 balance is enforced over faces only (``ensure_two_one_balance``, :1016-1057).

This is host-side workload construction, not part of the timed path.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, List, Optional, Sequence

import numpy as np

AF_NO_BOX = 0
AF_PHYS_BOUNDARY = -1
AF_XYZ = 1
AF_CYL = 2


def child_dix(ndim: int) -> np.ndarray:
    """af_child_dix (afivo/src/m_af_types.f90:172-174): shape (2^D, D), child c-1 -> offset."""
    n = 1 << ndim
    return np.array([[(c >> d) & 1 for d in range(ndim)] for c in range(n)], dtype=np.int64)


def neighb_dix(ndim: int) -> np.ndarray:
    """af_neighb_dix (afivo/src/m_af_types.f90:199-200): shape (2D, D)."""
    out = np.zeros((2 * ndim, ndim), dtype=np.int64)
    for d in range(ndim):
        out[2 * d, d] = -1
        out[2 * d + 1, d] = 1
    return out


@dataclasses.dataclass
class Tree:
    """Struct-of-arrays copy of the parts of ``af_t`` the multigrid path reads.

    Per-box arrays have ``highest_id + 1`` rows; row 0 is unused so that the
    reference's 1-based ids index them directly.
    """

    ndim: int
    nc: int
    coord_t: int
    coarse_grid_size: np.ndarray  # (D,) cells
    periodic: np.ndarray  # (D,) bool
    r_base: np.ndarray  # (D,)
    dr_base: np.ndarray  # (D,)
    highest_lvl: int
    highest_id: int
    lvl_ids: List[np.ndarray]  # per level (index 0 = level 1): ids in list order
    lvl: np.ndarray  # (n+1,)
    ix: np.ndarray  # (n+1, D) 1-based spatial index on the level grid
    parent: np.ndarray  # (n+1,)
    children: np.ndarray  # (n+1, 2^D)
    neighbors: np.ndarray  # (n+1, 2D)
    neighbor_mat: np.ndarray  # (n+1, 3^D), first offset fastest
    r_min: np.ndarray  # (n+1, D)
    dr: np.ndarray  # (n+1, D)

    # ---- derived helpers -------------------------------------------------
    def has_children(self, ids: np.ndarray) -> np.ndarray:
        return self.children[ids, 0] != AF_NO_BOX

    def leaves(self, lvl: int) -> np.ndarray:
        ids = self.lvl_ids[lvl - 1]
        return ids[~self.has_children(ids)]

    def parents(self, lvl: int) -> np.ndarray:
        ids = self.lvl_ids[lvl - 1]
        return ids[self.has_children(ids)]

    @property
    def n_boxes(self) -> int:
        return int(sum(len(a) for a in self.lvl_ids))

    def n_cells_level(self, lvl: int) -> int:
        return len(self.lvl_ids[lvl - 1]) * self.nc ** self.ndim

    @property
    def box_len(self) -> int:
        return (self.nc + 2) ** self.ndim

    def permuted_ids(self, rng: np.random.Generator) -> "Tree":
        """Return the same tree with box ids shuffled (ids are allocation order in the
        reference and get recycled, afivo/src/m_af_core.f90:885-922, so nothing may
        depend on them being spatially ordered).  Level-1 ids are kept because the
        coarse grid relies on them (afivo/src/m_af_core.f90:481)."""
        n = self.highest_id
        n1 = len(self.lvl_ids[0])
        perm = np.arange(n + 1)
        perm[n1 + 1:] = n1 + 1 + rng.permutation(n - n1)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(n + 1)

        def remap(a):
            a = np.asarray(a)
            out = a.copy()
            m = a > 0
            out[m] = perm[a[m]]
            return out

        def rows(a):
            return a[inv]

        return dataclasses.replace(
            self,
            lvl_ids=[remap(a) for a in self.lvl_ids],
            lvl=rows(self.lvl),
            ix=rows(self.ix),
            parent=remap(rows(self.parent)),
            children=remap(rows(self.children)),
            neighbors=remap(rows(self.neighbors)),
            neighbor_mat=remap(rows(self.neighbor_mat)),
            r_min=rows(self.r_min),
            dr=rows(self.dr),
        )


def _upsample(mask: np.ndarray) -> np.ndarray:
    for ax in range(mask.ndim):
        mask = np.repeat(mask, 2, axis=ax)
    return mask


def _downsample_any(mask: np.ndarray) -> np.ndarray:
    nd = mask.ndim
    shp = []
    for s in mask.shape:
        shp += [s // 2, 2]
    m = mask.reshape(shp)
    return m.any(axis=tuple(range(1, 2 * nd, 2)))


def _shift_or(mask: np.ndarray, periodic: Sequence[bool]) -> np.ndarray:
    """mask OR its six/four face-shifted copies (numpy axes are reversed dims)."""
    out = mask.copy()
    nd = mask.ndim
    for ax in range(nd):
        dim = nd - 1 - ax  # numpy axis ax <-> spatial dim
        for sh in (-1, 1):
            r = np.roll(mask, sh, axis=ax)
            if not periodic[dim]:
                sl = [slice(None)] * nd
                sl[ax] = 0 if sh == 1 else -1
                r[tuple(sl)] = False
            out |= r
    return out


RefineFn = Callable[[int, np.ndarray, np.ndarray], np.ndarray]


def build_tree(
    ndim: int,
    nc: int,
    coarse_grid_size: Sequence[int],
    max_lvl: int,
    refine_fn: Optional[RefineFn] = None,
    *,
    r_max: Optional[Sequence[float]] = None,
    r_min: Optional[Sequence[float]] = None,
    periodic: Optional[Sequence[bool]] = None,
    coord_t: int = AF_XYZ,
) -> Tree:
    """Build a 2:1 balanced tree.

    ``refine_fn(lvl, ix, centre)`` gets, for all existing boxes of level ``lvl``
    (1 <= lvl < max_lvl), their 1-based spatial indices ``ix`` (n, D) and centre
    coordinates (n, D), and returns a bool array: refine this box.  ``None`` means
    refine everything (uniform grid, ``af_refine_up_to_lvl``).
    """
    cgs = np.asarray(coarse_grid_size, dtype=np.int64)
    assert cgs.shape == (ndim,) and np.all(cgs % nc == 0) and nc % 2 == 0
    per = np.zeros(ndim, bool) if periodic is None else np.asarray(periodic, bool)
    rb = np.zeros(ndim) if r_min is None else np.asarray(r_min, float)
    rmax = np.ones(ndim) if r_max is None else np.asarray(r_max, float)
    dr_base = (rmax - rb) / cgs  # af_init: (r_max - r_min) / grid_size
    nb1 = cgs // nc  # level-1 boxes per dim

    # ---- which boxes exist / are refined, as dense per-level grids (numpy axis order z,y,x)
    def grid_shape(l):
        return tuple(int(n) << (l - 1) for n in nb1[::-1])

    exists = [None] * (max_lvl + 2)
    refined = [None] * (max_lvl + 2)
    exists[1] = np.ones(grid_shape(1), bool)
    for l in range(1, max_lvl):
        if refine_fn is None:
            ref = exists[l].copy()
        else:
            pos = np.argwhere(exists[l])  # (n, D) in (z,y,x)
            ixs = pos[:, ::-1] + 1
            dr_l = dr_base * 0.5 ** (l - 1)
            ctr = rb + (ixs - 0.5) * dr_l * nc
            flag = np.asarray(refine_fn(l, ixs, ctr), bool)
            ref = np.zeros(grid_shape(l), bool)
            ref[tuple(pos[flag].T)] = True
        refined[l] = ref
        exists[l + 1] = _upsample(ref)
    refined[max_lvl] = np.zeros(grid_shape(max_lvl), bool)
    # 2:1 balance over faces, finest to coarsest (ensure_two_one_balance)
    for l in range(max_lvl - 1, 1, -1):
        must_exist = _shift_or(refined[l], per)
        refined[l - 1] |= _downsample_any(must_exist)
    for l in range(1, max_lvl):
        exists[l + 1] = _upsample(refined[l])
    highest_lvl = max(l for l in range(1, max_lvl + 1) if exists[l].any())

    # ---- enumerate boxes level by level in the reference's list order
    cdix = child_dix(ndim)
    nch = 1 << ndim
    order_ix = []  # per level: (n, D) 1-based ix in list order
    g = np.indices(grid_shape(1)).reshape(ndim, -1).T  # lexicographic: x fastest since last axis
    order_ix.append(g[:, ::-1].astype(np.int64) + 1)
    for l in range(1, highest_lvl):
        ixl = order_ix[-1]
        is_par = refined[l][tuple((ixl[:, ::-1] - 1).T)]
        par = ixl[is_par]
        ch = (2 * par[:, None, :] - 1) + cdix[None, :, :]
        order_ix.append(ch.reshape(-1, ndim))

    counts = [len(a) for a in order_ix]
    n = int(sum(counts))
    starts = np.concatenate([[1], 1 + np.cumsum(counts)])
    lvl_ids = [np.arange(starts[i], starts[i] + counts[i], dtype=np.int32) for i in range(highest_lvl)]

    lvl = np.zeros(n + 1, np.int32)
    ix = np.zeros((n + 1, ndim), np.int32)
    parent = np.zeros(n + 1, np.int32)
    children = np.zeros((n + 1, nch), np.int32)
    neighbors = np.zeros((n + 1, 2 * ndim), np.int32)
    nmat = np.zeros((n + 1, 3 ** ndim), np.int32)
    rmin_a = np.zeros((n + 1, ndim))
    dr_a = np.zeros((n + 1, ndim))

    idgrid = []
    for li in range(highest_lvl):
        l = li + 1
        ids = lvl_ids[li]
        ixl = order_ix[li]
        lvl[ids] = l
        ix[ids] = ixl
        gr = np.zeros(grid_shape(l), np.int32)
        gr[tuple((ixl[:, ::-1] - 1).T)] = ids
        idgrid.append(gr)
        dr_a[ids] = dr_base * 0.5 ** (l - 1)
        if l == 1:
            rmin_a[ids] = rb + (ixl - 1) * dr_base * nc
        else:
            pg = idgrid[li - 1]
            pix = (ixl + 1) // 2
            pid = pg[tuple((pix[:, ::-1] - 1).T)]
            parent[ids] = pid
            c = ((ixl - 1) & 1)
            cidx = np.zeros(len(ids), np.int64)
            for d in range(ndim):
                cidx += c[:, d] << d
            children[pid, cidx] = ids
            # add_children: r_min = r_min_p + 0.5 * dr_p * dix * n_cell
            rmin_a[ids] = rmin_a[pid] + 0.5 * dr_a[pid] * c * nc

    # ---- neighbours: geometric lookup on the padded id grid
    offs = np.array(
        [[((m // 3 ** d) % 3) - 1 for d in range(ndim)] for m in range(3 ** ndim)], dtype=np.int64
    )
    for li in range(highest_lvl):
        ids = lvl_ids[li]
        ixl = order_ix[li]
        gr = idgrid[li]
        shape_xyz = np.array(gr.shape[::-1])
        for m in range(3 ** ndim):
            q = ixl - 1 + offs[m]
            outside = np.zeros(len(ids), bool)
            for d in range(ndim):
                if per[d]:
                    q[:, d] %= shape_xyz[d]
                else:
                    outside |= (q[:, d] < 0) | (q[:, d] >= shape_xyz[d])
            qc = np.clip(q, 0, shape_xyz - 1)
            val = gr[tuple(qc[:, ::-1].T)]
            val = np.where(outside, AF_PHYS_BOUNDARY, val)
            nmat[ids, m] = val
    centre = (3 ** ndim) // 2
    ndix = neighb_dix(ndim)
    for nb in range(2 * ndim):
        m = centre + int(sum(ndix[nb, d] * 3 ** d for d in range(ndim)))
        neighbors[:, nb] = nmat[:, m]
    neighbors[0] = 0
    nmat[0] = 0

    return Tree(
        ndim=ndim, nc=nc, coord_t=coord_t, coarse_grid_size=cgs.astype(np.int32), periodic=per,
        r_base=rb, dr_base=dr_base, highest_lvl=highest_lvl, highest_id=n, lvl_ids=lvl_ids,
        lvl=lvl, ix=ix, parent=parent, children=children, neighbors=neighbors, neighbor_mat=nmat,
        r_min=rmin_a, dr=dr_a,
    )


# ---- named workloads --------------------------------

def uniform_tree(ndim: int, nc: int, coarse: int, max_lvl: int, **kw) -> Tree:
    """S1: ``poisson_benchmark nc coarse max_lvl`` (afivo/examples/poisson_benchmark.f90:72-75,161-171)."""
    return build_tree(ndim, nc, [coarse] * ndim, max_lvl, None, **kw)


def corner_refined_tree(ndim: int, nc: int, coarse: int, max_lvl: int) -> Tree:
    """Tree refined towards the low corner, as in afivo/tests/test_ghostcell.f90:54-76
    (refine a box iff its lowest corner is the domain origin)."""
    def fn(l, ixs, ctr):
        return np.all(ixs == 1, axis=1)
    return build_tree(ndim, nc, [coarse] * ndim, max_lvl, fn)


def channel_tree(nc: int = 8, coarse: int = 8, max_lvl: int = 9, uniform_lvls: int = 3, width: float = 2.0) -> Tree:
    """S2 stand-in for ``programs/standard_3d``: levels 1..uniform_lvls uniform, then refine
    boxes whose centre lies within ``width`` box lengths (width * 0.5^(l-1)) of the segment
    (0.5,0.5,0.35)-(0.5,0.5,0.65): a streamer-channel-like refinement down to max_lvl."""
    a = np.array([0.5, 0.5, 0.35])
    b = np.array([0.5, 0.5, 0.65])

    def fn(l, ixs, ctr):
        if l < uniform_lvls:
            return np.ones(len(ixs), bool)
        ab = b - a
        t = np.clip(((ctr - a) @ ab) / (ab @ ab), 0.0, 1.0)
        d = np.linalg.norm(ctr - (a + t[:, None] * ab), axis=1)
        return d < width * 0.5 ** (l - 1)

    return build_tree(3, nc, [coarse] * 3, max_lvl, fn)


def shell_tree(nc: int = 16, coarse: int = 16, uniform_lvls: int = 6, margin: float = 0.46875) -> Tree:
    """S3: levels 1..uniform_lvls uniform, one more level on every box whose centre satisfies
    max|x-0.5| < margin (all but the outermost box layer): a closed refinement boundary."""
    def fn(l, ixs, ctr):
        if l < uniform_lvls:
            return np.ones(len(ixs), bool)
        return np.max(np.abs(ctr - 0.5), axis=1) < margin

    return build_tree(3, nc, [coarse] * 3, uniform_lvls + 1, fn)
