"""-m gpu: boxes with explicit stencils: variable / constant
epsilon (mg_box_lpld_stencil + mg_box_prolong_eps_stencil + mg_sides_rb_extrap) and level-set boxes
(mg_box_lsf_stencil with bc_correction).  The stencils are built by the oracle's restatement of the
reference's host-side builders and shipped through afmg_set_stencils, as the Fortran shim does.
Single operations must be bit-identical; whole cycles within 1e-10 (coarse solve: dense inverse on
the GPU vs banded LU in the oracle)."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from util import all_ids, assert_same_state, bc_mixed, fill_all_ghosts, stencils_from_oracle

pytestmark = pytest.mark.gpu


def eps_smooth(r):
    return 1.0 + 0.5 * np.sin(2 * np.pi * r[..., 0]) * np.cos(2 * np.pi * r[..., 1]) + 0.25 * r[..., 2]


def eps_jump(r):
    return np.where(r[..., 2] < 0.5, 1.0, 3.0)


def eps_const2(r):
    return np.full(r.shape[:-1], 2.0)


def lsf_sphere(r, centre=(0.45, 0.55, 0.5), radius=0.2):
    return np.linalg.norm(r - np.asarray(centre), axis=-1) - radius


def lsf_distances(tree, lsf):
    """all_distances(2*ndim, cells) per box with mg_lsf_dist_linear (m_af_multigrid.f90:1635-1647):
    la / (la - lb) where the level set changes sign towards the neighbour cell, else 1."""
    ids = all_ids(tree)
    r = W.cell_centres(tree, ids, ghosts=True)  # (n, z, y, x, 3)
    v = lsf(r)
    nc = tree.nc
    c = v[:, 1:-1, 1:-1, 1:-1]
    out = np.ones((len(ids), nc, nc, nc, 6))
    nbs = [v[:, 1:-1, 1:-1, :-2], v[:, 1:-1, 1:-1, 2:], v[:, 1:-1, :-2, 1:-1], v[:, 1:-1, 2:, 1:-1],
           v[:, :-2, 1:-1, 1:-1], v[:, 2:, 1:-1, 1:-1]]
    for m, b in enumerate(nbs):
        cut = c * b < 0
        out[..., m] = np.where(cut, c / np.where(cut, c - b, 1.0), 1.0)
    has = np.any(out < 1.0, axis=(1, 2, 3, 4))
    return ids[has], out[has].reshape(int(has.sum()), -1)


def make_pair(tree, *, eps=None, lsf=None, seed=3, **opts):
    bc = W.bc_table(tree, bc_mixed)
    orc = Oracle(tree, with_eps=eps is not None, **opts)
    orc.set_bc(bc)
    ids = all_ids(tree)
    if eps is not None:
        orc.set_cc(M.I_EPS, ids, eps(W.cell_centres(tree, ids, ghosts=True)))
    if lsf is not None:
        lids, dd = lsf_distances(tree, lsf)
        assert len(lids) > 0
        orc.set_lsf_distances(lids, dd)
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc, **opts)
    M.mg_init(tree, mg)
    entries = stencils_from_oracle(tree, orc)
    assert entries, "the case must contain boxes with explicit stencils"
    mg.set_stencils(entries)
    rng = np.random.default_rng(seed)
    shape = (len(ids),) + (tree.nc + 2,) * 3
    for var in (M.I_RHS, M.I_PHI, M.I_TMP):
        data = rng.uniform(-1, 1, shape)
        orc.set_cc(var, ids, data)
        mg.set_cc(var, ids, data)
    return orc, mg, entries


CASES = {
    "eps_smooth_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), dict(eps=eps_smooth)),
    "eps_jump_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 3), dict(eps=eps_jump)),
    "eps_const2_corner_nc4": (lambda: T.corner_refined_tree(3, 4, 8, 3), dict(eps=eps_const2)),
    "eps_smooth_uniform_nc16": (lambda: T.uniform_tree(3, 16, 16, 2), dict(eps=eps_smooth)),
    "lsf_sphere_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 3), dict(lsf=lsf_sphere, lsf_boundary_value=1.5)),
    "lsf_sphere_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), dict(lsf=lsf_sphere, lsf_boundary_value=-0.7)),
    # mg_box_lpld_lsf_stencil (m_af_multigrid.f90:1535-1623): permittivity and electrode in the same boxes
    "eps_lsf_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4),
                           dict(eps=eps_smooth, lsf=lsf_sphere, lsf_boundary_value=0.9)),
    # periodic domain with a non-separable coarse operator: the dense coarse solve with wrap-around couplings
    "eps_smooth_periodic_xy_nc8": (lambda: T.uniform_tree(3, 8, 8, 3, periodic=[True, True, False]), dict(eps=eps_smooth)),
    "ceps_lsf_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 3), dict(eps=eps_const2, lsf=lsf_sphere, lsf_boundary_value=-1.2)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_single_operations_bit_exact(name):
    mk, kw = CASES[name]
    tree = mk()
    orc, mg, entries = make_pair(tree, **kw)
    kinds = {(e["op"][0], e["tag"]) for e in entries}
    assert kinds, kinds
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
    M.mg_destroy(mg)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cycles_match_oracle(name):
    mk, kw = CASES[name]
    tree = mk()
    orc, mg, _ = make_pair(tree, **kw)
    ids = all_ids(tree)
    zeros = np.zeros((len(ids),) + (tree.nc + 2,) * 3)
    orc.set_cc(M.I_PHI, ids, zeros)
    mg.set_cc(M.I_PHI, ids, zeros)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(4):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    ho, hg = np.array(ho), np.array(hg)
    assert ho[-1] < 0.2 * ho[0], ho
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)


def test_lsf_boundary_value_can_change():
    tree = T.uniform_tree(3, 8, 8, 3)
    orc, mg, _ = make_pair(tree, lsf=lsf_sphere, lsf_boundary_value=1.0)
    orc.set_opts(lsf_boundary_value=2.5)
    orc.mg_init()
    mg.set_lsf_boundary_value(2.5)
    orc.fas_fmg(True, True)
    M.mg_fas_fmg(tree, mg, True, True)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)
