"""-m gpu: solve from an afivo .dat file alone: a simulation state (tree, phi, rhs, stored
boundary conditions, operator / prolongation / level-set distance stencils) is written with the .dat writer,
read back, and solved on the GPU with nothing but the file; the oracle solves the original state."""
import numpy as np
import pytest

from afivo_streamer_b200 import datfile as D
from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from dat_util import make_dat
from util import all_ids, bc_mixed
import test_gpu_2d as G2
import test_gpu_stencils as G3

pytestmark = pytest.mark.gpu

CASES = {
    "plain_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), {}),
    "eps_smooth_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 3), dict(eps=G3.eps_smooth)),
    "lsf_sphere_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), dict(lsf=G3.lsf_sphere, lsf_boundary_value=-0.7)),
    "cyl_lsf_2d": (lambda: T.build_tree(2, 8, [8, 8], 4, lambda l, ix, c: np.linalg.norm(c - 0.5, axis=1) < 0.4,
                                         coord_t=T.AF_CYL), dict(lsf=G2.lsf_circle, lsf_boundary_value=0.8, bc=G2.bc_cyl)),
    "permuted_ids_with_gaps": (lambda: T.corner_refined_tree(3, 8, 8, 3).permuted_ids(np.random.default_rng(11)), {}),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_solve_from_dat(tmp_path, name):
    mk, kw = CASES[name]
    kw = dict(kw)
    tree = mk()
    ids = all_ids(tree)
    bc = W.bc_table(tree, kw.pop("bc", bc_mixed))
    eps, lsf = kw.pop("eps", None), kw.pop("lsf", None)
    orc = Oracle(tree, with_eps=eps is not None, **kw)
    orc.set_bc(bc)
    extra, lsf_dd = {}, None
    if eps is not None:
        e = np.zeros((tree.highest_id + 1, tree.box_len))
        e[ids] = eps(W.cell_centres(tree, ids, ghosts=True)).reshape(len(ids), -1)
        orc.set_cc(M.I_EPS, ids, e[ids])
        extra["eps"] = e
    if lsf is not None:
        lsf_dd = (G3.lsf_distances if tree.ndim == 3 else G2.lsf_distances2)(tree, lsf)
        orc.set_lsf_distances(*lsf_dd)
        v = np.zeros((tree.highest_id + 1, tree.box_len))
        v[ids] = lsf(W.cell_centres(tree, ids, ghosts=True)).reshape(len(ids), -1)
        extra["lsf"] = v
        sl = (slice(None),) + (slice(1, -1),) * tree.ndim
        shape = (len(lsf_dd[0]),) + (tree.nc + 2,) * tree.ndim
        orc.set_lsf_cc(lsf_dd[0], v[lsf_dd[0]].reshape(shape)[sl].reshape(len(lsf_dd[0]), -1))
    orc.mg_init()
    rid, rhs = W.random_rhs_on_leaves(tree)
    orc.set_cc(M.I_RHS, rid, rhs)
    path = str(tmp_path / "state.dat")
    D.write_tree(path, make_dat(tree, orc, bc, extra_cc=extra, lsf_dd=lsf_dd))

    dat = D.read_tree(path)
    tree2, mg = M.mg_from_dat(dat, eps="eps" if eps is not None else None, **kw)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree2, mg, True, False)
    for _ in range(3):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree2, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree2, mg, True)
    ho, hg = np.array(ho), np.array(hg)
    assert ho[-1] < 0.3 * ho[0]
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    a, b = orc.get_cc(M.I_PHI, ids), mg.get_cc(M.I_PHI, ids).reshape(len(ids), -1)
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(a))
    # and the field from it, level-set correction included
    orc.compute_phi_gradient(-1.0, True)
    M.mg_compute_phi_gradient(tree2, mg, -1.0, True)
    fa, fb = orc.get_fc(ids), mg.get_fc(ids)
    assert np.max(np.abs(fa - fb)) <= 1e-8 * np.max(np.abs(fa))
    M.mg_destroy(mg)
