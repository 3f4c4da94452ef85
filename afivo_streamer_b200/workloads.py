"""Synthetic inputs for the multigrid path: boundary-condition tables and right-hand sides.

In the reference, physical-boundary ghost cells are produced by a per-face callback
``sides_bc(box, nb, iv, coords, bc_val, bc_type)`` (afivo/src/m_af_types.f90:401-420) whose
result never depends on phi (afivo/src/m_af_ghostcell.f90:615-652, src/m_field.f90:590-670,
src/m_photoi_helmh.f90:210-228).  Across the C ABI the callback is therefore evaluated on the
host once per solve and shipped as a table: one row per physical face.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import numpy as np

from .tree import AF_PHYS_BOUNDARY, Tree

# afivo/src/m_af_types.f90:58-69
AF_BC_DIRICHLET = -10
AF_BC_NEUMANN = -11
AF_BC_CONTINUOUS = -12
AF_BC_DIRICHLET_COPY = -13


@dataclasses.dataclass
class BCTable:
    ids: np.ndarray  # (n,) box id
    nbs: np.ndarray  # (n,) neighbour direction 1..2D
    types: np.ndarray  # (n,) af_bc_*
    vals: np.ndarray  # (n, nc^(D-1)); index a + (b-1)*nc over the transverse dims in increasing order


def face_coords(tree: Tree, ids: np.ndarray, nb: int) -> np.ndarray:
    """af_get_face_coords (afivo/src/m_af_types.f90:1215-1253): (n, nface, D)."""
    nd, nc = tree.ndim, tree.nc
    d = (nb - 1) // 2
    low = (nb % 2) == 1
    td = [q for q in range(nd) if q != d]
    nface = nc ** (nd - 1)
    out = np.zeros((len(ids), nface, nd))
    rmin = tree.r_min[ids]
    dr = tree.dr[ids]
    out[:, :, d] = (rmin[:, d] if low else rmin[:, d] + nc * dr[:, d])[:, None]
    a = np.arange(nface) % nc
    b = np.arange(nface) // nc
    out[:, :, td[0]] = (rmin[:, td[0]] + 0.5 * dr[:, td[0]])[:, None] + a[None, :] * dr[:, td[0]][:, None]
    if nd == 3:
        out[:, :, td[1]] = (rmin[:, td[1]] + 0.5 * dr[:, td[1]])[:, None] + b[None, :] * dr[:, td[1]][:, None]
    return out


def bc_table(tree: Tree, fn: Callable) -> BCTable:
    """``fn(nb, coords)`` -> (bc_type, values) with values broadcastable to (n, nface)."""
    nd, nc = tree.ndim, tree.nc
    nface = nc ** (nd - 1)
    all_ids = np.concatenate(tree.lvl_ids)
    ids_l, nbs_l, ty_l, val_l = [], [], [], []
    for nb in range(1, 2 * nd + 1):
        ids = all_ids[tree.neighbors[all_ids, nb - 1] == AF_PHYS_BOUNDARY]
        if len(ids) == 0:
            continue
        ty, vals = fn(nb, face_coords(tree, ids, nb))
        ids_l.append(ids)
        nbs_l.append(np.full(len(ids), nb, np.int32))
        ty_l.append(np.full(len(ids), ty, np.int32))
        val_l.append(np.broadcast_to(np.asarray(vals, float), (len(ids), nface)).copy())
    if not ids_l:
        return BCTable(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, nface)))
    return BCTable(np.concatenate(ids_l).astype(np.int32), np.concatenate(nbs_l), np.concatenate(ty_l),
                   np.concatenate(val_l))


def bc_dirichlet_zero(tree: Tree) -> BCTable:
    """af_bc_dirichlet_zero (afivo/src/m_af_ghostcell.f90:629-639)."""
    return bc_table(tree, lambda nb, c: (AF_BC_DIRICHLET, 0.0))


def bc_neumann_zero(tree: Tree) -> BCTable:
    """af_bc_neumann_zero (afivo/src/m_af_ghostcell.f90:616-626)."""
    return bc_table(tree, lambda nb, c: (AF_BC_NEUMANN, 0.0))


def bc_field_homogeneous(tree: Tree, voltage: float = 1.0) -> BCTable:
    """field_bc_homogeneous (src/m_field.f90:590-610): last dim Dirichlet 0 / voltage, others Neumann 0."""
    nd = tree.ndim

    def fn(nb, c):
        if (nb - 1) // 2 == nd - 1:
            return AF_BC_DIRICHLET, (0.0 if nb % 2 == 1 else voltage)
        return AF_BC_NEUMANN, 0.0

    return bc_table(tree, fn)


def bc_field_neumann(tree: Tree, voltage: float = 1.0) -> BCTable:
    """field_bc_neumann (src/m_field.f90:614-634; field_bc_type = "neumann"): last dim Dirichlet 0 at the low side and
    Neumann voltage / domain_len at the high side, others Neumann 0."""
    nd = tree.ndim
    length = float(tree.coarse_grid_size[nd - 1] * tree.dr_base[nd - 1])

    def fn(nb, c):
        if (nb - 1) // 2 == nd - 1:
            return (AF_BC_DIRICHLET, 0.0) if nb % 2 == 1 else (AF_BC_NEUMANN, voltage / length)
        return AF_BC_NEUMANN, 0.0

    return bc_table(tree, fn)


def bc_field_all_neumann(tree: Tree) -> BCTable:
    """field_bc_all_neumann (src/m_field.f90:637-647; "all_neumann", used with electrodes that fix the potential)."""
    return bc_neumann_zero(tree)


def bc_field_all_dirichlet(tree: Tree) -> BCTable:
    """field_bc_all_dirichlet (src/m_field.f90:650-669; the coaxial electrode set-up): Dirichlet 0 everywhere, except
    zero flux on the axis of a cylindrical domain."""
    cyl = tree.coord_t == 2

    def fn(nb, c):
        return (AF_BC_NEUMANN, 0.0) if (cyl and nb == 1) else (AF_BC_DIRICHLET, 0.0)

    return bc_table(tree, fn)


def bc_helmholtz(tree: Tree) -> BCTable:
    """photoi_helmh_bc (src/m_photoi_helmh.f90:210-228): last dim Dirichlet 0, others Neumann 0."""
    return bc_field_homogeneous(tree, 0.0)


def bc_dirichlet_function(tree: Tree, func: Callable[[np.ndarray], np.ndarray]) -> BCTable:
    """Dirichlet values from an analytic solution, like sides_bc of afivo/examples/poisson_basic.f90:219-235."""
    return bc_table(tree, lambda nb, c: (AF_BC_DIRICHLET, func(c)))


# ---- cell-centred fields -----------------------------------------------------

def cell_centres(tree: Tree, ids: np.ndarray, ghosts: bool = True) -> np.ndarray:
    """af_r_cc (afivo/src/m_af_types.f90:1035-1040) for all cells of the boxes: (n, [nz,] ny, nx, D)."""
    nd, nc = tree.ndim, tree.nc
    rng = np.arange(0, nc + 2) if ghosts else np.arange(1, nc + 1)
    ax = [tree.r_min[ids, d][:, None] + (rng[None, :] - 0.5) * tree.dr[ids, d][:, None] for d in range(nd)]
    m = len(rng)
    out = np.empty((len(ids),) + (m,) * nd + (nd,))
    for d in range(nd):
        shp = [len(ids)] + [1] * nd
        shp[nd - d] = m  # numpy axis for spatial dim d (x is last)
        out[..., d] = ax[d].reshape(shp)
    return out


def box_array(tree: Tree, n: int) -> np.ndarray:
    """Zero cc storage for n boxes in Fortran element order (x fastest): (n, [nc+2,] nc+2, nc+2)."""
    return np.zeros((n,) + (tree.nc + 2,) * tree.ndim)


def interior(tree: Tree):
    return (slice(None),) + (slice(1, tree.nc + 1),) * tree.ndim


def random_rhs_on_leaves(tree: Tree, seed: int = 12345):
    """S1r right-hand side: uniform(-1, 1) on the interior cells of all leaves, drawn in
    (level, box-list order, k, j, i) order.  Returns (ids, data)."""
    rng = np.random.default_rng(seed)
    ids = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    data = box_array(tree, len(ids))
    data[interior(tree)] = rng.uniform(-1.0, 1.0, size=(len(ids),) + (tree.nc,) * tree.ndim)
    return ids, data


def constant_rhs_on_leaves(tree: Tree, value: float = 1.0):
    """S1: rhs == 1 (afivo/examples/poisson_benchmark.f90:173-180)."""
    ids = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    data = box_array(tree, len(ids))
    data[interior(tree)] = value
    return ids, data


# ---- Gaussian manufactured solution (afivo/examples/m_gaussians.f90:54-105) ----

class Gaussians:
    """Sum of Gaussians exp(-|r-r0|^2/sigma^2) and its Laplacian."""

    def __init__(self, r0, sigma):
        self.r0 = np.atleast_2d(np.asarray(r0, float))
        self.sigma = np.broadcast_to(np.asarray(sigma, float), (len(self.r0),))

    def value(self, r):
        out = np.zeros(r.shape[:-1])
        for r0, s in zip(self.r0, self.sigma):
            out += np.exp(-np.sum((r - r0) ** 2, axis=-1) / s ** 2)
        return out

    def laplacian(self, r, cyl=False):
        nd = r.shape[-1]
        out = np.zeros(r.shape[:-1])
        for r0, s in zip(self.r0, self.sigma):
            xrel = (r - r0) / s
            d2 = np.sum(xrel ** 2, axis=-1)
            out += 4 / s ** 2 * (d2 - 0.5 * nd) * np.exp(-d2)
        return out
