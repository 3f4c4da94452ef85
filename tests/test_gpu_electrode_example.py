"""-m gpu: afivo/examples/electrode_example.f90 in 2D (Cartesian, and cylindrical as with its command-line argument):
a rod electrode (0.4, 0.4)-(0.6, 0.6) of radius 0.02 at potential 1 in a grounded box, coarse grid 4 x 4 boxes of 8^2,
refined while lvl < 5 and r_min(1) < 0.5; 10 FMG cycles each followed by mg_compute_phi_gradient.  The problem is set
up by the library's own builders from the built-in rod level-set function (mg_set_operators_tree); the oracle gets
the same distances.  Checked: residual history, potential, field, and the physics (0 <= phi <= 1, phi = 1 inside)."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import stencils as S
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("coord", [T.AF_XYZ, T.AF_CYL])
def test_electrode_example_2d(coord):
    nc = 8
    t = T.build_tree(2, nc, [4 * nc] * 2, 5,
                     lambda l, ixs, ctr: (l < 5) & ((ixs[:, 0] - 1) * (0.25 / 2 ** (l - 1)) < 0.5), coord_t=coord)
    assert t.highest_lvl == 5
    el = S.electrode("rod", 2, rod_r0=(0.4, 0.4), rod_r1=(0.6, 0.6), rod_radius=0.02)
    bc = W.bc_dirichlet_zero(t)
    mg = M.mg_t(sides_bc=bc, lsf_boundary_value=1.0)
    M.mg_init(t, mg)
    entries, data = M.mg_set_operators_tree(t, mg, lsf=el)
    assert len(entries) == len(data.ids) > 0
    orc = Oracle(t, lsf_boundary_value=1.0)
    orc.set_bc(bc)
    orc.set_lsf_distances(data.ids, data.dd.reshape(len(data.ids), -1))
    orc.set_lsf_cc(data.ids, data.lsf_cells)
    orc.mg_init()
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    shape = (len(ids), nc + 2, nc + 2)
    inner = W.interior(t)
    res_o, res_g = [], []
    for it in range(10):
        orc.fas_fmg(True, it > 0)
        M.mg_fas_fmg(t, mg, True, it > 0)
        res_o.append(orc.maxabs(M.I_TMP))
        res_g.append(M.af_tree_maxabs_cc(t, mg, M.I_TMP))
        if coord == T.AF_XYZ:
            orc.compute_phi_gradient(1.0, True)
            M.mg_compute_phi_gradient(t, mg, 1.0, True)
    res_o, res_g = np.array(res_o), np.array(res_g)
    assert res_o[5] < 1e-6 * res_o[0]
    assert np.all(np.abs(res_g - res_o) <= 1e-10 * res_o[0] + 2e-8), (res_o, res_g)
    po = orc.get_cc(M.I_PHI, ids).reshape(shape)
    pg = mg.get_cc(M.I_PHI, ids)
    assert np.max(np.abs(po - pg)) <= 1e-10 * np.max(np.abs(po))
    if coord == T.AF_XYZ:
        fo = orc.get_cc(M.I_FLD, ids).reshape(shape)[inner]
        fg = mg.get_cc(M.I_FLD, ids)[inner]
        assert np.max(fo) > 10 and np.max(np.abs(fo - fg)) <= 1e-7 * np.max(fo)
    # physics on the leaves: maximum principle and the electrode interior at the electrode potential
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    phi = mg.get_cc(M.I_PHI, leaves)[inner]
    assert phi.min() > -1e-3 and phi.max() < 1 + 1e-3
    sel = np.isin(leaves, data.ids)
    pos = {int(b): n for n, b in enumerate(data.ids)}
    lsf = np.stack([data.lsf_cells[pos[int(b)]] for b in leaves[sel]]).reshape(-1, nc, nc)
    dr = t.dr[leaves[sel]][:, 0][:, None, None]
    inside = lsf < -dr
    assert inside.any() and np.max(np.abs(phi[sel][inside] - 1)) < 2e-3
    M.mg_destroy(mg)


def test_electrode_example_3d_large_coarse_grid():
    """afivo/examples/electrode_example.f90 in 3D: box_size 8, coarse grid 32^3 (64 level-1 boxes, :41-45), refined
    while lvl < 3 and r_min(1) < 0.5 (:84), rod electrode from (0.4, 0.4, 0.4) to (0.6, 0.6, 0.6), radius 0.02, at
    potential 1 in a grounded box.  The electrode crosses level-1 boxes, so the coarse operator has explicit stencils
    on 32 768 cells: beyond the dense inverse (8192), solved by the block-tridiagonal plane solver (k_cs_plane_*) --
    the case that returned AFMG_ERR_UNSUPPORTED in round 1.  Oracle: banded LU of the same matrix."""
    nc = 8
    t = T.build_tree(3, nc, [4 * nc] * 3, 3,
                     lambda l, ixs, ctr: (l < 3) & ((ixs[:, 0] - 1) * (0.25 / 2 ** (l - 1)) < 0.5))
    assert t.highest_lvl == 3 and len(t.lvl_ids[0]) == 64
    el = S.electrode("rod", 3, rod_r0=(0.4, 0.4, 0.4), rod_r1=(0.6, 0.6, 0.6), rod_radius=0.02)
    bc = W.bc_dirichlet_zero(t)
    mg = M.mg_t(sides_bc=bc, lsf_boundary_value=1.0)
    M.mg_init(t, mg)
    entries, data = M.mg_set_operators_tree(t, mg, lsf=el)
    assert any(t.lvl[e["box_id"]] == 1 for e in entries), "the electrode must reach the coarse grid"
    orc = Oracle(t, lsf_boundary_value=1.0)
    orc.set_bc(bc)
    orc.set_lsf_distances(data.ids, data.dd.reshape(len(data.ids), -1))
    orc.set_lsf_cc(data.ids, data.lsf_cells)
    orc.mg_init()
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    res_o, res_g = [], []
    for it in range(6):
        orc.fas_fmg(True, it > 0)
        M.mg_fas_fmg(t, mg, True, it > 0)
        res_o.append(orc.maxabs(M.I_TMP))
        res_g.append(M.af_tree_maxabs_cc(t, mg, M.I_TMP))
    res_o, res_g = np.array(res_o), np.array(res_g)
    assert res_o[-1] < 1e-5 * res_o[0], res_o
    assert np.all(np.abs(res_g - res_o) <= 1e-9 * res_o[0] + 1e-6 * res_o), (res_o, res_g)
    po = orc.get_cc(M.I_PHI, ids)
    pg = mg.get_cc(M.I_PHI, ids).reshape(po.shape)
    assert np.max(np.abs(po - pg)) <= 1e-10 * np.max(np.abs(po))
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    phi = mg.get_cc(M.I_PHI, leaves)[W.interior(t)]
    # maximum principle up to the discretisation of the electrode surface (0.4 % overshoot next to it at this resolution)
    assert phi.min() > -1e-3 and phi.max() < 1 + 1e-2 and phi.max() > 0.9
    M.mg_destroy(mg)
