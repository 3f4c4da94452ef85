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


def lsf_prolong_distances(tree, lsf):
    """Distances of mg_box_prolong_lsf_stencil (m_af_multigrid.f90:1392-1482) with mg_lsf_dist_linear: from each
    fine cell centre to its ndim+1 coarse prolongation points (i_c1,j_c1,k_c1), (i_c2,..), (.., j_c2, ..), (.., k_c2);
    cells away from the surface (outside the root mask) are marked by dd[0] = -1."""
    nd, nc = tree.ndim, tree.nc
    ids = np.array([b for b in all_ids(tree) if tree.lvl[b] > 1], np.int32)
    a = W.cell_centres(tree, ids, ghosts=False)  # (n, [z,] y, x, nd)
    la = lsf(a)
    fine = np.arange(1, nc + 1)
    out = np.ones((len(ids),) + (nc,) * nd + (nd + 1,))
    pr = tree.parent[ids]
    off = ((tree.ix[ids] - 1) & 1) * (nc // 2)  # af_get_child_offset
    c1 = [off[:, d][:, None] + (fine[None, :] + 1) // 2 for d in range(nd)]       # (n, nc) per dim
    c2 = [c1[d] + 1 - 2 * (fine[None, :] & 1) for d in range(nd)]

    def coarse_point(sel):  # sel[d] in {1, 2}: which coarse index along dim d
        shape = (len(ids),) + (nc,) * nd + (nd,)
        r = np.empty(shape)
        for d in range(nd):
            cidx = (c1 if sel[d] == 1 else c2)[d]  # (n, nc) along dim d
            coord = tree.r_min[pr][:, d][:, None] + (cidx - 0.5) * tree.dr[pr][:, d][:, None]
            view = [1] * (nd + 1)
            view[0] = len(ids)
            view[nd - d] = nc  # arrays are ordered ([z,] y, x)
            r[..., d] = coord.reshape(view)
        return r

    sels = [[1] * nd] + [[2 if q == d else 1 for q in range(nd)] for d in range(nd)]
    for m, sel in enumerate(sels):
        lb = lsf(coarse_point(sel))
        cut = la * lb < 0
        out[..., m] = np.where(cut, np.maximum(la / np.where(cut, la - lb, 1.0), 1e-4), 1.0)
    norm_dr = np.linalg.norm(tree.dr[ids], axis=1).reshape((len(ids),) + (1,) * nd)
    out[..., 0] = np.where(np.abs(la) < 2 * norm_dr, out[..., 0], -1.0)
    return ids, out.reshape(len(ids), -1)


def make_pair(tree, *, eps=None, lsf=None, seed=3, custom_prolong=False, **opts):
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
        if custom_prolong:  # mg%lsf_use_custom_prolongation
            orc.set_lsf_prolong_distances(*lsf_prolong_distances(tree, lsf))
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
    # mg%lsf_use_custom_prolongation: mg_box_prolong_lsf_stencil (variable p234 weights that vanish behind the surface)
    "lsf_custom_prolong_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4),
                                      dict(lsf=lsf_sphere, lsf_boundary_value=0.8, custom_prolong=True)),
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


@pytest.mark.parametrize("name", ["eps_lsf_corner_nc8", "lsf_sphere_uniform_nc8", "eps_smooth_uniform_nc16"])
def test_fused_generic_halfsweep_bit_exact(name, monkeypatch):
    """AFMG_GSRB_FUSED_GEN=1: k_gsrb2g sweeps all boxes of a level that holds explicit-stencil boxes in one launch
    (opt-in: measured slower than the two launches side by side); same bits as the oracle"""
    monkeypatch.setenv("AFMG_GSRB_FUSED_GEN", "1")
    mk, kw = CASES[name]
    tree = mk()
    orc, mg, _ = make_pair(tree, **kw)
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(tree.highest_lvl, 1, -1):
        for cyc in (M.MG_CYCLE_DOWN, M.MG_CYCLE_UP):
            orc.gsrb_boxes(lvl, cyc)
            mg.gsrb_boxes(lvl, cyc)
            assert_same_state(tree, orc, mg)
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


def two_spheres(r):
    a = np.linalg.norm(r - np.array([0.3, 0.35, 0.4]), axis=-1) - 0.14
    b = np.linalg.norm(r - np.array([0.7, 0.65, 0.6]), axis=-1) - 0.12
    return np.minimum(a, b)


def two_sphere_potential(r, v1=1.7, v2=-0.6):
    """mg%lsf_boundary_function of a two-electrode set-up (rod_rod_get_potential, src/m_field.f90:802-825): the
    potential of whichever electrode is closer."""
    a = np.linalg.norm(r - np.array([0.3, 0.35, 0.4]), axis=-1) - 0.14
    b = np.linalg.norm(r - np.array([0.7, 0.65, 0.6]), axis=-1) - 0.12
    return np.where(a < b, v1, v2)


def test_lsf_boundary_function_two_electrodes():
    """Per-cell boundary values (afmg_set_lsf_boundary_values) in bc_correction, in the coarse-grid right-hand
    side and in the level-set gradient; then new potentials for the same boxes without re-shipping anything else."""
    tree = T.uniform_tree(3, 8, 8, 3)
    orc, mg, _ = make_pair(tree, lsf=two_spheres, lsf_boundary_value=9.9)  # the scalar must not be used
    ids = all_ids(tree)
    lids, dd = lsf_distances(tree, two_spheres)
    inner = (slice(None),) + (slice(1, -1),) * 3
    for v1, v2 in ((1.7, -0.6), (0.25, 3.0)):
        bv = two_sphere_potential(W.cell_centres(tree, lids, ghosts=True), v1, v2)[inner].reshape(len(lids), -1)
        assert len(np.unique(bv)) == 2
        orc.set_lsf_boundary_values(lids, bv)
        orc.mg_init()
        mg.set_lsf_boundary_values(lids, bv)
        zeros = np.zeros((len(ids),) + (tree.nc + 2,) * 3)
        rng = np.random.default_rng(11)
        for var, data in ((M.I_PHI, zeros), (M.I_RHS, rng.uniform(-1, 1, zeros.shape)), (M.I_TMP, rng.uniform(-1, 1, zeros.shape))):
            orc.set_cc(var, ids, data)  # identical state in all three variables (cycles leave unobservable
            mg.set_cc(var, ids, data)   # differences in tmp / rhs ghost cells behind)
        # single operations bit for bit
        fill_all_ghosts(tree, orc, mg)
        for lvl in range(tree.highest_lvl, 1, -1):
            orc.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
            mg.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
            orc.update_coarse(lvl, True)
            mg.update_coarse(lvl, True)
            assert_same_state(tree, orc, mg)
        # cycles
        orc.fas_fmg(True, False)
        M.mg_fas_fmg(tree, mg, True, False)
        for _ in range(3):
            orc.fas_vcycle(True)
            M.mg_fas_vcycle(tree, mg, True)
        ro, rg = orc.maxabs(M.I_TMP), M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        assert abs(ro - rg) <= 1e-6 * ro + 1e-9
        assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
        # the electrodes really sit at their own potentials: phi next to electrode 1 is close to v1
        phi = mg.get_cc(M.I_PHI, lids)[inner].reshape(len(lids), -1)
        cut = (dd.reshape(len(lids), -1, 6) < 1.0).any(axis=2)
        near1 = cut & (bv == v1)
        assert near1.any() and np.all(np.abs(phi[near1] - v1) < 0.75 * abs(v1 - v2))
        # field with the level-set correction
        vals = two_spheres(W.cell_centres(tree, lids, ghosts=True))[inner].reshape(len(lids), -1)
        orc.set_lsf_cc(lids, vals)
        mg.set_lsf_distances(lids, dd, vals)
        orc.compute_phi_gradient(-1.0, True)
        M.mg_compute_phi_gradient(tree, mg, -1.0, True)
        fa, fb = orc.get_fc(ids), mg.get_fc(ids)
        assert np.max(np.abs(fa - fb)) <= 1e-8 * np.max(np.abs(fa))
    # back to the scalar
    orc.set_lsf_boundary_values(np.zeros(0, np.int32), np.zeros((0, 8 ** 3)))
    orc.set_opts(lsf_boundary_value=0.4)
    orc.mg_init()
    mg.set_lsf_boundary_values(np.zeros(0, np.int32), np.zeros((0, 8 ** 3)))
    mg.set_lsf_boundary_value(0.4)
    orc.fas_fmg(True, True)
    M.mg_fas_fmg(tree, mg, True, True)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)
