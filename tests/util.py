"""Shared helpers of the parity tests: build the CPU oracle and the CUDA solver on identical
trees, boundary conditions and data."""
import numpy as np

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle


def all_ids(tree):
    return np.concatenate(tree.lvl_ids).astype(np.int32)


def interior_of(tree, a):
    return a[W.interior(tree)]


def make_pair(tree, bc_fn, *, seed=0, random_phi=True, **opts):
    """Return (oracle, mg) holding the same rhs (random on all boxes) and phi."""
    bc = W.bc_table(tree, bc_fn)
    orc = Oracle(tree, **opts)
    orc.set_bc(bc)
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc, **opts)
    M.mg_init(tree, mg)
    rng = np.random.default_rng(seed)
    ids = all_ids(tree)
    shape = (len(ids),) + (tree.nc + 2,) * tree.ndim
    rhs = rng.uniform(-1, 1, shape)
    phi = rng.uniform(-1, 1, shape) if random_phi else np.zeros(shape)
    tmp = rng.uniform(-1, 1, shape)
    for var, data in ((M.I_RHS, rhs), (M.I_PHI, phi), (M.I_TMP, tmp)):
        orc.set_cc(var, ids, data)
        mg.set_cc(var, ids, data)
    return orc, mg


def fill_all_ghosts(tree, orc, mg):
    for lvl in range(1, tree.highest_lvl + 1):
        orc.gc_lvl(lvl, M.I_PHI, True)
        mg.gc_lvl(lvl, M.I_PHI, True)


def assert_same_state(tree, orc, mg, *, exact=True, rtol=0.0, what=("phi", "tmp", "rhs")):
    ids = all_ids(tree)
    for name in what:
        var = {"phi": M.I_PHI, "tmp": M.I_TMP, "rhs": M.I_RHS}[name]
        a = orc.get_cc(var, ids).reshape((len(ids),) + (tree.nc + 2,) * tree.ndim)
        b = mg.get_cc(var, ids)
        if name == "rhs":  # ghost cells of rhs are never meaningful
            a, b = interior_of(tree, a), interior_of(tree, b)
        if exact:
            bad = np.argwhere(a != b)
            assert len(bad) == 0, f"{name}: {len(bad)} cells differ, first at {bad[0]} (box id {ids[bad[0][0]]}): " \
                                  f"{a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}"
        else:
            scale = np.max(np.abs(a))
            err = np.max(np.abs(a - b))
            assert err <= rtol * scale, f"{name}: rel max-norm error {err / scale:.3e} > {rtol:.1e}"


def bc_mixed(nb, coords):
    """field_bc_homogeneous-like (src/m_field.f90:590-610) but with non-trivial values: Dirichlet
    in the last dimension, Neumann elsewhere, values varying over the face."""
    d = (nb - 1) // 2
    vals = 0.3 * np.sin(3.0 * coords.sum(axis=-1)) + 0.1 * nb
    if d == coords.shape[-1] - 1:
        return W.AF_BC_DIRICHLET, vals
    return W.AF_BC_NEUMANN, vals


TREES = {
    "uniform_nc8_l3": lambda: T.uniform_tree(3, 8, 8, 3),
    "corner_nc8_l4": lambda: T.corner_refined_tree(3, 8, 8, 4),
    "uniform_nc16_l2": lambda: T.uniform_tree(3, 16, 16, 2),
    "corner_nc4_l4": lambda: T.corner_refined_tree(3, 4, 8, 4),
    "multibox_coarse_nc8": lambda: T.build_tree(3, 8, [16, 8, 24], 3,
                                                lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45),
    "permuted_ids": lambda: T.corner_refined_tree(3, 8, 8, 3).permuted_ids(np.random.default_rng(7)),
}

# 2D trees of the golden fixtures (Cartesian and cylindrical, SURVEY config C1)
TREES2D = {
    "xy2d_uniform_nc8_l4": lambda: T.uniform_tree(2, 8, 8, 4),
    "cyl2d_corner_nc8_l5": lambda: T.build_tree(2, 8, [8, 16], 5, lambda l, ix, c: np.all(ix == 1, axis=1),
                                                r_max=[1.0, 2.0], coord_t=T.AF_CYL),
}


def stencils_from_oracle(tree, orc):
    """What the Fortran shim ships with afmg_set_stencils: the stencils mg_set_operators_lvl stored for
    every box that is not a plain constant-Laplacian box (the oracle's builders stand in for the
    reference's host-side ones)."""
    out = []
    for lvl in range(1, tree.highest_lvl + 1):
        for bid in tree.lvl_ids[lvl - 1]:
            tag = orc.tag(bid)
            if tag == 0:
                continue
            stype, coeff, f, cyl = orc.op_stencil(bid)
            e = dict(box_id=int(bid), tag=tag, op=(stype, coeff), f=f, cyl=cyl)
            if lvl > 1:
                pst, pshape, pco = orc.prolong_stencil(bid)
                e["prolong"] = (pst, pshape, pco)
            out.append(e)
    return out
