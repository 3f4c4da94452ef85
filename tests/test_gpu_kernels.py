"""-m gpu: every CUDA kernel of the path against the CPU oracle on the same inputs, bit-exact
(integer/index work and, thanks to -fmad=false, all fp64 arithmetic except the coarse solve)."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M

from util import TREES, all_ids, assert_same_state, bc_mixed, fill_all_ghosts, make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(params=sorted(TREES))
def pair(request):
    tree = TREES[request.param]()
    orc, mg = make_pair(tree, bc_mixed, seed=11)
    yield tree, orc, mg
    M.mg_destroy(mg)


def test_upload_download_roundtrip(pair):
    tree, orc, mg = pair
    ids = all_ids(tree)[::-1].copy()
    rng = np.random.default_rng(3)
    data = rng.normal(size=(len(ids),) + (tree.nc + 2,) * 3)
    mg.set_cc(M.I_TMP, ids, data)
    back = mg.get_cc(M.I_TMP, ids)
    assert np.array_equal(back, data)


def test_gc_lvl_bit_exact(pair):
    tree, orc, mg = pair
    for corners in (False, True):
        for lvl in range(1, tree.highest_lvl + 1):
            orc.gc_lvl(lvl, M.I_PHI, corners)
            mg.gc_lvl(lvl, M.I_PHI, corners)
    assert_same_state(tree, orc, mg, what=("phi",))


@pytest.mark.parametrize("type_cycle", [M.MG_CYCLE_DOWN, M.MG_CYCLE_UP])
def test_gsrb_boxes_bit_exact(pair, type_cycle):
    tree, orc, mg = pair
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(tree.highest_lvl, 1, -1):
        orc.gsrb_boxes(lvl, type_cycle)
        mg.gsrb_boxes(lvl, type_cycle)
    assert_same_state(tree, orc, mg, what=("phi",))


def test_gsrb_use_corners(pair):
    tree, orc, mg = pair
    M.mg_destroy(mg)
    orc, mg = make_pair(tree, bc_mixed, seed=5, use_corners=True)
    fill_all_ghosts(tree, orc, mg)
    lvl = tree.highest_lvl
    orc.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
    mg.gsrb_boxes(lvl, M.MG_CYCLE_DOWN)
    assert_same_state(tree, orc, mg, what=("phi",))
    M.mg_destroy(mg)


@pytest.mark.parametrize("with_tmp", [True, False])
def test_update_coarse_bit_exact(pair, with_tmp):
    tree, orc, mg = pair
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(tree.highest_lvl, 1, -1):
        orc.update_coarse(lvl, with_tmp)
        mg.update_coarse(lvl, with_tmp)
    assert_same_state(tree, orc, mg)


def test_correct_children_bit_exact(pair):
    tree, orc, mg = pair
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(1, tree.highest_lvl):
        orc.correct_children(lvl)
        mg.correct_children(lvl)
    assert_same_state(tree, orc, mg)


def test_correct_children_then_gc_bit_exact(pair):
    """correct_children + af_gc_lvl as fused in the cycles (push from the prolongation kernel)."""
    tree, orc, mg = pair
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(1, tree.highest_lvl):
        orc.correct_children(lvl)
        orc.gc_lvl(lvl + 1, M.I_PHI, True)
        mg.correct_children_gc(lvl)
    assert_same_state(tree, orc, mg)


def test_residual_and_maxabs(pair):
    tree, orc, mg = pair
    fill_all_ghosts(tree, orc, mg)
    for lvl in range(1, tree.highest_lvl + 1):
        orc.residual_lvl(lvl)
        mg.residual_lvl(lvl)
    assert_same_state(tree, orc, mg, what=("tmp",))
    assert M.af_tree_maxabs_cc(tree, mg, M.I_TMP) == orc.maxabs(M.I_TMP)
    assert M.af_tree_maxabs_cc(tree, mg, M.I_PHI) == orc.maxabs(M.I_PHI)
    a, b = M.af_tree_sum_cc(tree, mg, M.I_PHI), orc.tree_sum(M.I_PHI)
    assert abs(a - b) <= 1e-12 * max(1.0, abs(b))


def test_init_phi_rhs_bit_exact(pair):
    tree, orc, mg = pair
    orc.init_phi_rhs()
    mg.init_phi_rhs()
    assert_same_state(tree, orc, mg)


def test_coarse_solve(pair):
    tree, orc, mg = pair
    orc.solve_coarse_grid()
    mg.solve_coarse_grid()
    # direct solves by different algorithms (banded LU vs fast diagonalisation): equal to round-off
    assert_same_state(tree, orc, mg, exact=False, rtol=1e-12, what=("phi",))
