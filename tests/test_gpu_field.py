"""-m gpu: field from potential on the device: mg_compute_phi_gradient with
mg_box_lpl_gradient (incl. the eps-weighted boundary faces of variable-eps boxes), mg_box_lpllsf_gradient,
mg_box_field_norm, mg_compute_field_norm and af_gc_tree of the norm (af_bc_neumann_zero + af_gc_interp).
Everything is box-local fp64 arithmetic in the reference's expression order: bit-identical to the oracle."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

from util import all_ids, bc_mixed, fill_all_ghosts
from util import make_pair as make_plain
import test_gpu_2d as G2
import test_gpu_stencils as G3

pytestmark = pytest.mark.gpu


def lsf_values(tree, ids, lsf):
    v = lsf(W.cell_centres(tree, ids, ghosts=True))
    sl = (slice(None),) + (slice(1, -1),) * tree.ndim
    return v[sl].reshape(len(ids), -1)


def build(name):
    kind, mk, kw = CASES[name]
    kw = dict(kw)
    bc_fn = kw.pop("bc", bc_mixed)
    tree = mk()
    if kind in ("plain3", "plain2"):
        orc, mg = make_plain(tree, bc_fn, **kw)
    elif kind == "st3":
        orc, mg, _ = G3.make_pair(tree, **kw)
    else:
        orc, mg = G2.make_pair(tree, bc_fn, **kw)
    ids = all_ids(tree)
    if kw.get("eps") is not None:
        mg.set_cc(M.I_EPS, ids, kw["eps"](W.cell_centres(tree, ids, ghosts=True)))
    if kw.get("lsf") is not None:
        # cells on both sides of the boundary carry distances < 1, so both outcomes of the reference's
        # `cc(IJK, i_lsf) >= 0` test occur
        lids, dd = (G3.lsf_distances if tree.ndim == 3 else G2.lsf_distances2)(tree, kw["lsf"])
        vals = lsf_values(tree, lids, kw["lsf"])
        assert (vals < 0).any() and (vals > 0).any()
        orc.set_lsf_cc(lids, vals)
        mg.set_lsf_distances(lids, dd, vals)
    return tree, orc, mg


CASES = {
    "uniform_nc8": ("plain3", lambda: T.uniform_tree(3, 8, 8, 3), {}),
    "corner_nc8_l4": ("plain3", lambda: T.corner_refined_tree(3, 8, 8, 4), {}),
    "corner_nc16_l3": ("plain3", lambda: T.corner_refined_tree(3, 16, 16, 3), {}),
    "multibox_nc4": ("plain3", lambda: T.build_tree(3, 4, [8, 4, 12], 3,
                                                    lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45), {}),
    "eps_smooth_corner_nc8": ("st3", lambda: T.corner_refined_tree(3, 8, 8, 4), dict(eps=G3.eps_smooth)),
    "eps_jump_uniform_nc8": ("st3", lambda: T.uniform_tree(3, 8, 8, 3), dict(eps=G3.eps_jump)),
    "lsf_sphere_corner_nc8": ("st3", lambda: T.corner_refined_tree(3, 8, 8, 4),
                              dict(lsf=G3.lsf_sphere, lsf_boundary_value=-0.7)),
    "lsf_sphere_uniform_nc8": ("st3", lambda: T.uniform_tree(3, 8, 8, 3), dict(lsf=G3.lsf_sphere, lsf_boundary_value=1.5)),
    "eps_lsf_corner_nc8": ("st3", lambda: T.corner_refined_tree(3, 8, 8, 4),
                           dict(eps=G3.eps_smooth, lsf=G3.lsf_sphere, lsf_boundary_value=0.9)),
    "xy_corner_nc8": ("plain2", lambda: T.corner_refined_tree(2, 8, 8, 5), {}),
    "cyl_channel_nc8": ("plain2", lambda: T.build_tree(2, 8, [8, 8], 6, lambda l, ix, c: (c[:, 0] < 1.5 * 0.5 ** (l - 1)) &
                                                      (np.abs(c[:, 1] - 0.5) < 0.3), coord_t=T.AF_CYL), dict(bc=G2.bc_cyl)),
    "xy_eps_corner_nc8": ("st2", lambda: T.corner_refined_tree(2, 8, 8, 4), dict(eps=G2.eps2)),
    "xy_lsf_corner_nc8": ("st2", lambda: T.corner_refined_tree(2, 8, 8, 4), dict(lsf=G2.lsf_circle, lsf_boundary_value=0.8)),
}


def same(a, b, what, ids):
    a = a.reshape(len(ids), -1)
    b = b.reshape(len(ids), -1)
    bad = np.argwhere(a != b)
    assert len(bad) == 0, f"{what}: {len(bad)} values differ, first at box id {ids[bad[0][0]]} offset {bad[0][1]}: " \
                          f"{a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_gradient_norm_ghostcells_bit_exact(name):
    tree, orc, mg = build(name)
    ids = all_ids(tree)
    fill_all_ghosts(tree, orc, mg)
    # gradient without norm, then the norm from fc
    orc.compute_phi_gradient(-1.0, False)
    M.mg_compute_phi_gradient(tree, mg, -1.0, False)
    same(orc.get_fc(ids), mg.get_fc(ids), "fc", ids)
    assert np.max(np.abs(mg.get_fc(ids))) > 0
    orc.compute_field_norm()
    M.mg_compute_field_norm(tree, mg)
    inner = W.interior(tree)
    shape = (len(ids),) + (tree.nc + 2,) * tree.ndim
    same(orc.get_cc(M.I_FLD, ids).reshape(shape)[inner], mg.get_cc(M.I_FLD, ids)[inner], "norm", ids)
    # fused gradient + norm with another factor
    orc.compute_phi_gradient(2.5, True)
    M.mg_compute_phi_gradient(tree, mg, 2.5, True)
    same(orc.get_fc(ids), mg.get_fc(ids), "fc (fused)", ids)
    same(orc.get_cc(M.I_FLD, ids).reshape(shape)[inner], mg.get_cc(M.I_FLD, ids)[inner], "norm (fused)", ids)
    # ghost cells of the norm: sides, edges, corners on all levels
    orc.gc_tree(M.I_FLD, True)
    M.af_gc_tree(tree, mg, M.I_FLD, True)
    same(orc.get_cc(M.I_FLD, ids), mg.get_cc(M.I_FLD, ids), "norm incl. ghost cells", ids)
    M.mg_destroy(mg)


@pytest.mark.parametrize("name", ["corner_nc8_l4", "lsf_sphere_corner_nc8", "xy_corner_nc8"])
def test_field_from_potential_after_solve(name):
    """The callers' sequence (field_compute, src/m_field.f90:448-528): solve, then field_from_potential."""
    tree, orc, mg = build(name)
    ids = all_ids(tree)
    zeros = np.zeros((len(ids),) + (tree.nc + 2,) * tree.ndim)
    orc.set_cc(M.I_PHI, ids, zeros)
    mg.set_cc(M.I_PHI, ids, zeros)
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    orc.compute_phi_gradient(-1.0, True)
    orc.gc_tree(M.I_FLD, True)
    M.field_from_potential(tree, mg, -1.0)
    a, b = orc.get_cc(M.I_FLD, ids), mg.get_cc(M.I_FLD, ids).reshape(len(ids), -1)
    assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(a))  # phi itself agrees to 1e-10 (coarse solve)
    fa, fb = orc.get_fc(ids), mg.get_fc(ids)
    assert np.max(np.abs(fa - fb)) <= 1e-9 * np.max(np.abs(fa))
    M.mg_destroy(mg)


def test_fc_round_trip_and_custom_bc():
    tree = T.corner_refined_tree(3, 8, 8, 3)
    orc, mg = make_plain(tree, bc_mixed)
    ids = all_ids(tree)
    rng = np.random.default_rng(5)
    fc = rng.uniform(-1, 1, (len(ids), mg.fc_len()))
    mg.set_fc(ids, fc)
    orc.set_fc(ids, fc)
    same(fc, mg.get_fc(ids), "fc round trip", ids)
    orc.compute_field_norm()
    M.mg_compute_field_norm(tree, mg)
    bc = W.bc_table(tree, bc_mixed)  # Dirichlet / Neumann with non-trivial values for the norm variable
    orc.set_fld_bc(bc)
    mg.set_fld_bc(bc)
    orc.gc_tree(M.I_FLD, True)
    M.af_gc_tree(tree, mg, M.I_FLD, True)
    same(orc.get_cc(M.I_FLD, ids), mg.get_cc(M.I_FLD, ids), "norm with custom bc", ids)
    M.mg_destroy(mg)


def test_gc_tree_phi_matches_level_loop():
    tree = T.corner_refined_tree(3, 8, 8, 4)
    orc, mg = make_plain(tree, bc_mixed)
    ids = all_ids(tree)
    orc.gc_tree(M.I_PHI, True)
    M.af_gc_tree(tree, mg, M.I_PHI, True)
    same(orc.get_cc(M.I_PHI, ids), mg.get_cc(M.I_PHI, ids), "phi ghost cells", ids)
    M.mg_destroy(mg)
