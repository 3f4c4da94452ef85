"""-m gpu: a permittivity + electrode problem set up entirely by the product's own host-side builders
(mg_set_operators_tree -> afmg_build_box_* -> afmg_set_stencils; no oracle-built stencils on the way in) solves like
the oracle given the same permittivity and level-set distances."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

import test_gpu_2d as G2
import test_gpu_stencils as G3
from util import all_ids, assert_same_state, bc_mixed

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nd", [3, 2])
def test_problem_built_by_the_library_solves_like_the_oracle(nd):
    if nd == 3:
        tree, eps, lsf, bc_fn = T.corner_refined_tree(3, 8, 8, 4), G3.eps_smooth, G3.lsf_sphere, bc_mixed
    else:
        tree, eps, lsf, bc_fn = T.corner_refined_tree(2, 8, 8, 4), G2.eps2, G2.lsf_circle, bc_mixed
    bc = W.bc_table(tree, bc_fn)
    ids = all_ids(tree)
    e = eps(W.cell_centres(tree, ids, ghosts=True))
    eps_cc = np.zeros((tree.highest_id + 1,) + e.shape[1:])
    eps_cc[ids] = e
    mg = M.mg_t(sides_bc=bc, lsf_boundary_value=0.9)
    M.mg_init(tree, mg)
    entries, data = M.mg_set_operators_tree(tree, mg, eps_cc=eps_cc, lsf=lsf)
    assert any(en["tag"] & 1 for en in entries) and any(en["tag"] & 2 for en in entries)
    orc = Oracle(tree, with_eps=True, lsf_boundary_value=0.9)
    orc.set_bc(bc)
    orc.set_cc(M.I_EPS, ids, e)
    orc.set_lsf_distances(data.ids, data.dd.reshape(len(data.ids), -1))
    orc.mg_init()
    rng = np.random.default_rng(21)
    shape = (len(ids),) + (tree.nc + 2,) * nd
    rhs = rng.uniform(-1, 1, shape)
    for s in (orc, mg):
        s.set_cc(M.I_RHS, ids, rhs)
        s.set_cc(M.I_PHI, ids, np.zeros(shape))
    lvl = tree.highest_lvl
    orc.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)  # one smoother call: the shipped coefficients are the oracle's, bit for bit
    mg.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
    assert_same_state(tree, orc, mg, exact=True, what=("phi",))
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(2):
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)
