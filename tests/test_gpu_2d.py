"""-m gpu: the 2D path: Cartesian and cylindrical trees, refinement
boundaries, explicit stencils (variable eps, level set).  Single operations bit-identical to the oracle,
whole cycles within 1e-10."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from util import all_ids, assert_same_state, bc_mixed, fill_all_ghosts, stencils_from_oracle

pytestmark = pytest.mark.gpu

CYL = T.AF_CYL


def bc_cyl(nb, coords):
    """field_bc_homogeneous-like in (r, z): Neumann-0 on the axis and the outer radius, Dirichlet in z."""
    d = (nb - 1) // 2
    if d == 1:
        return W.AF_BC_DIRICHLET, 0.25 * nb + 0.1 * np.cos(2.0 * coords[..., 0])
    return W.AF_BC_NEUMANN, np.zeros(coords.shape[:-1])


def eps2(r):
    return 1.0 + 0.5 * np.sin(2 * np.pi * r[..., 0]) * np.cos(2 * np.pi * r[..., 1])


def lsf_circle(r):
    return np.linalg.norm(r - np.array([0.45, 0.55]), axis=-1) - 0.2


def lsf_distances2(tree, lsf):
    ids = all_ids(tree)
    v = lsf(W.cell_centres(tree, ids, ghosts=True))  # (n, y, x)
    nc = tree.nc
    c = v[:, 1:-1, 1:-1]
    out = np.ones((len(ids), nc, nc, 4))
    for m, b in enumerate([v[:, 1:-1, :-2], v[:, 1:-1, 2:], v[:, :-2, 1:-1], v[:, 2:, 1:-1]]):
        cut = c * b < 0
        out[..., m] = np.where(cut, c / np.where(cut, c - b, 1.0), 1.0)
    has = np.any(out < 1.0, axis=(1, 2, 3))
    return ids[has], out[has].reshape(int(has.sum()), -1)


CASES = {
    "xy_uniform_nc8": (lambda: T.uniform_tree(2, 8, 8, 4), bc_mixed, {}),
    "xy_corner_nc8": (lambda: T.corner_refined_tree(2, 8, 8, 5), bc_mixed, {}),
    "xy_corner_nc16_sparse": (lambda: T.corner_refined_tree(2, 16, 16, 3), bc_mixed,
                              dict(prolongation_type=M.MG_PROLONG_SPARSE, use_corners=True)),
    "xy_multibox_nc4": (lambda: T.build_tree(2, 4, [8, 12], 4, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.35),
                        bc_mixed, {}),
    "cyl_uniform_nc8": (lambda: T.build_tree(2, 8, [8, 8], 4, None, coord_t=CYL), bc_cyl, {}),
    "cyl_corner_nc8_helmholtz": (lambda: T.build_tree(2, 8, [8, 16], 5, lambda l, ix, c: np.all(ix == 1, axis=1),
                                                      r_max=[1.0, 2.0], coord_t=CYL), bc_cyl, dict(helmholtz_lambda=30.0)),
    "cyl_channel_nc8": (lambda: T.build_tree(2, 8, [8, 8], 6, lambda l, ix, c: (c[:, 0] < 1.5 * 0.5 ** (l - 1)) & (np.abs(c[:, 1] - 0.5) < 0.3),
                                             coord_t=CYL), bc_cyl, {}),
    "xy_periodic_x_nc8": (lambda: T.uniform_tree(2, 8, 8, 4, periodic=[True, False]), bc_mixed, {}),
    "xy_periodic_x_refined_multibox_nc8": (lambda: T.build_tree(2, 8, [16, 8], 4, lambda l, ix, c: c[:, 1] < 0.55,
                                                               periodic=[True, False]), bc_mixed, {}),
    "xy_eps_corner_nc8": (lambda: T.corner_refined_tree(2, 8, 8, 4), bc_mixed, dict(eps=eps2)),
    "cyl_eps_uniform_nc8": (lambda: T.build_tree(2, 8, [8, 8], 3, None, coord_t=CYL), bc_cyl, dict(eps=eps2)),
    "xy_lsf_uniform_nc8": (lambda: T.uniform_tree(2, 8, 8, 4), bc_mixed, dict(lsf=lsf_circle, lsf_boundary_value=1.5)),
    "cyl_lsf_uniform_nc8": (lambda: T.build_tree(2, 8, [8, 8], 4, None, coord_t=CYL), bc_cyl,
                            dict(lsf=lambda r: np.linalg.norm(r - np.array([0.0, 0.5]), axis=-1) - 0.2, lsf_boundary_value=-0.5)),
    # mg_box_lpld_lsf_stencil: permittivity and electrode in the same boxes, Cartesian and cylindrical
    "xy_eps_lsf_corner_nc8": (lambda: T.corner_refined_tree(2, 8, 8, 4), bc_mixed,
                              dict(eps=eps2, lsf=lsf_circle, lsf_boundary_value=0.6)),
    "cyl_eps_lsf_uniform_nc8": (lambda: T.build_tree(2, 8, [8, 8], 4, None, coord_t=CYL), bc_cyl,
                                dict(eps=eps2, lsf=lambda r: np.linalg.norm(r - np.array([0.0, 0.5]), axis=-1) - 0.2,
                                     lsf_boundary_value=-0.5)),
}


def make_pair(tree, bc_fn, *, eps=None, lsf=None, seed=5, zero_phi=False, **opts):
    bc = W.bc_table(tree, bc_fn)
    orc = Oracle(tree, with_eps=eps is not None, **opts)
    orc.set_bc(bc)
    ids = all_ids(tree)
    if eps is not None:
        orc.set_cc(M.I_EPS, ids, eps(W.cell_centres(tree, ids, ghosts=True)))
    if lsf is not None:
        lids, dd = lsf_distances2(tree, lsf)
        assert len(lids) > 0
        orc.set_lsf_distances(lids, dd)
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc, **opts)
    M.mg_init(tree, mg)
    entries = stencils_from_oracle(tree, orc)
    if eps is not None or lsf is not None:
        assert entries
    if entries:
        mg.set_stencils(entries)
    rng = np.random.default_rng(seed)
    shape = (len(ids),) + (tree.nc + 2,) * 2
    for var in (M.I_RHS, M.I_PHI, M.I_TMP):
        data = rng.uniform(-1, 1, shape)
        if var == M.I_PHI and zero_phi:
            data[:] = 0.0
        orc.set_cc(var, ids, data)
        mg.set_cc(var, ids, data)
    return orc, mg


@pytest.mark.parametrize("name", sorted(CASES))
def test_single_operations_bit_exact(name):
    mk, bc_fn, kw = CASES[name]
    tree = mk()
    orc, mg = make_pair(tree, bc_fn, **kw)
    fill_all_ghosts(tree, orc, mg)
    assert_same_state(tree, orc, mg, what=("phi",))
    L = tree.highest_lvl
    for lvl in range(L, 1, -1):
        orc.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
        mg.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
        assert_same_state(tree, orc, mg)
        orc.update_coarse(lvl, True)
        mg.update_coarse(lvl, True)
        assert_same_state(tree, orc, mg)
    for lvl in range(2, L + 1):
        orc.correct_children(lvl - 1)
        orc.gc_lvl(lvl, M.I_PHI, True)
        mg.correct_children_gc(lvl - 1)
        assert_same_state(tree, orc, mg)
        orc.gsrb_boxes(lvl, M.MG_CYCLE_UP)
        mg.gsrb_boxes(lvl, M.MG_CYCLE_UP)
        assert_same_state(tree, orc, mg)
    for lvl in range(1, L + 1):
        orc.residual_lvl(lvl)
        mg.residual_lvl(lvl)
    assert_same_state(tree, orc, mg)
    assert orc.maxabs(M.I_TMP) == M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
    a, b = M.af_tree_sum_cc(tree, mg, M.I_PHI), orc.tree_sum(M.I_PHI)
    assert abs(a - b) <= 1e-12 * max(1.0, abs(b))
    orc.init_phi_rhs()
    mg.init_phi_rhs()
    assert_same_state(tree, orc, mg)
    # set_coarse_phi_rhs path of the FMG prologue
    for lvl in range(L, 1, -1):
        orc.update_coarse(lvl, False)
        mg.update_coarse(lvl, False)
    assert_same_state(tree, orc, mg)
    M.mg_destroy(mg)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cycles_match_oracle(name):
    mk, bc_fn, kw = CASES[name]
    tree = mk()
    orc, mg = make_pair(tree, bc_fn, zero_phi=True, **kw)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(4):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    orc.fas_fmg(True, True)
    M.mg_fas_fmg(tree, mg, True, True)
    ho.append(orc.maxabs(M.I_TMP))
    hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    ho, hg = np.array(ho), np.array(hg)
    assert ho[-1] < 0.1 * ho[0], ho
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)


def test_subtract_mean_2d_cyl():
    tree = T.build_tree(2, 8, [8, 8], 3, None, coord_t=CYL)
    orc, mg = make_pair(tree, bc_cyl, zero_phi=True, subtract_mean=True, helmholtz_lambda=5.0)
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)
