"""CPU: the oracle's restatement of "field from potential" (mg_compute_phi_gradient, mg_box_lpllsf_gradient,
mg_box_field_norm, af_gc_interp) against known answers.  The reference holds no numeric vectors for these
routines; what it does hold is exactness for linear fields (afivo/examples/check_ghostcells.f90 test_gradient,
check_prolongation.f90), which pins index maps, weights and expression structure."""
import numpy as np
import pytest

from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import I_FLD, I_PHI, Oracle


def all_ids(tree):
    return np.concatenate(tree.lvl_ids).astype(np.int32)


TREES = {
    "3d_corner": lambda: T.corner_refined_tree(3, 8, 8, 4),
    "3d_multibox": lambda: T.build_tree(3, 4, [8, 4, 12], 3, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45),
    "2d_corner": lambda: T.corner_refined_tree(2, 8, 8, 5),
}


def fc_views(tree, fc):
    """fc(nc+1, ..., NDIM) per box -> list over dims of the defined faces"""
    nd, n1, nc = tree.ndim, tree.nc + 1, tree.nc
    a = fc.reshape((len(fc), nd) + (n1,) * nd)  # (box, dim, [z,] y, x)
    out = []
    for d in range(nd):
        sl = [slice(None), d] + [slice(0, nc)] * nd
        sl[2 + (nd - 1 - d)] = slice(0, n1)
        out.append(a[tuple(sl)])
    return out


@pytest.mark.parametrize("name", sorted(TREES))
def test_linear_potential_gives_constant_field_everywhere(name):
    tree = TREES[name]()
    nd = tree.ndim
    g = np.array([1.5, -0.75, 2.25][:nd])
    ids = all_ids(tree)
    r = W.cell_centres(tree, ids, ghosts=True)
    phi = r @ g + 0.3
    orc = Oracle(tree)
    orc.set_bc(W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0)))
    orc.mg_init()
    orc.set_cc(I_PHI, ids, phi)
    orc.compute_phi_gradient(-1.0, True)
    for d, v in enumerate(fc_views(tree, orc.get_fc(ids))):
        assert np.allclose(v, -g[d], rtol=0, atol=1e-10), (d, np.abs(v + g[d]).max())
    norm = np.linalg.norm(g)
    orc.gc_tree(I_FLD, True)
    fld = orc.get_cc(I_FLD, ids)
    # interior, sides (copy / af_bc_neumann_zero / af_gc_interp: weights sum to one), edges and corners
    assert np.allclose(fld, norm, rtol=0, atol=1e-10), np.abs(fld - norm).max()


@pytest.mark.parametrize("name", sorted(TREES))
def test_gc_interp_and_extrapolation_exact_for_linear_field(name):
    """check_ghostcells.f90 test_gradient: ghost cells of a linear field are exact (af_gc_interp on refinement
    boundaries, af_bc_continuous on the domain boundary, edge / corner extrapolation)."""
    tree = TREES[name]()
    nd = tree.ndim
    g = np.array([0.7, -1.1, 0.4][:nd])
    ids = all_ids(tree)
    exact = W.cell_centres(tree, ids, ghosts=True) @ g + 2.0
    start = np.zeros_like(exact)
    start[W.interior(tree)] = exact[W.interior(tree)]
    orc = Oracle(tree)
    orc.set_bc(W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0)))
    orc.mg_init()
    orc.set_cc(I_FLD, ids, start)
    orc.set_fld_bc(W.bc_table(tree, lambda nb, c: (W.AF_BC_CONTINUOUS, 0.0)))
    orc.gc_tree(I_FLD, True)
    got = orc.get_cc(I_FLD, ids).reshape(exact.shape)
    assert np.allclose(got, exact, rtol=0, atol=1e-12), np.abs(got - exact).max()


def test_lsf_gradient_planar_electrode():
    """A planar electrode at x = x0 held at V, phi = V inside and V + s (x - x0) outside: the plain difference
    across the cut face is s * dd, mg_box_lpllsf_gradient (m_af_multigrid.f90:2055-2137) restores s from the
    positive side, and faces between the cut cell and the electrode interior are left alone."""
    tree = T.uniform_tree(3, 8, 8, 2)
    x0, V, s = 0.4321, 0.7, 2.0
    ids = all_ids(tree)
    r = W.cell_centres(tree, ids, ghosts=True)
    x = r[..., 0]
    lsf = x - x0
    phi = np.where(lsf >= 0, V + s * lsf, V)
    nc = tree.nc
    c = lsf[:, 1:-1, 1:-1, 1:-1]
    dd = np.ones((len(ids), nc, nc, nc, 6))
    for m, b in enumerate([lsf[:, 1:-1, 1:-1, :-2], lsf[:, 1:-1, 1:-1, 2:]]):
        cut = c * b < 0
        dd[..., m] = np.where(cut, c / np.where(cut, c - b, 1.0), 1.0)  # mg_lsf_dist_linear (:1635-1647)
    has = np.any(dd < 1.0, axis=(1, 2, 3, 4))
    orc = Oracle(tree, lsf_boundary_value=V)
    orc.set_bc(W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0)))
    orc.set_lsf_distances(ids[has], dd[has].reshape(int(has.sum()), -1))
    orc.set_lsf_cc(ids[has], c[has].reshape(int(has.sum()), -1))
    orc.mg_init()
    orc.set_cc(I_PHI, ids, phi)
    orc.compute_phi_gradient(1.0, True)
    fx = fc_views(tree, orc.get_fc(ids))[0]  # (box, z, y, xface)
    leaves = np.array([tree.children[i][0] == 0 for i in ids])
    xf = np.concatenate([x[:, 1:-1, 1:-1, 1:2] - 0.5 * (x[:, 1:-1, 1:-1, 2:3] - x[:, 1:-1, 1:-1, 1:2]),
                         0.5 * (x[:, 1:-1, 1:-1, 1:-1] + x[:, 1:-1, 1:-1, 2:])], axis=-1)  # face positions
    dx = (x[:, 1, 1, 2] - x[:, 1, 1, 1])[:, None, None, None]
    outside = xf - 0.5 * dx > x0 - 1e-12      # low cell of the face has lsf >= 0 -> both cells outside
    inside = xf + 0.5 * dx < x0 + 1e-12       # both cells inside the electrode
    cutf = ~outside & ~inside
    lv = leaves[:, None, None, None]
    assert np.allclose(fx[outside & lv], s, atol=1e-10)
    assert np.allclose(fx[inside & lv], 0.0, atol=1e-10)
    assert cutf.any() and np.allclose(fx[cutf & lv], s, atol=1e-10), np.abs(fx[cutf & lv] - s).max()
    # boxes with children keep the plain gradient (:1876-1882): s * dd on the cut faces
    par = ~leaves
    assert par.any() and not np.allclose(fx[cutf & par[:, None, None, None]], s, atol=1e-6)
