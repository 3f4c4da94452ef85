"""-m gpu: photoi_helmh_compute on the device (src/m_photoi_helmh.f90:162-204): the three
Bourdon Helmholtz modes solved back to back with a shared right-hand side, i_photo = -sum c_n phi_n on the
leaves.  The oracle side runs the same loop with one CPU solver per mode."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from util import all_ids

pytestmark = pytest.mark.gpu

# Bourdon-3 (src/m_photoi_helmh.f90:113-124) at 1 bar, 20 % O2, on a 2 cm domain scaled to the unit cube
LAMBDAS = np.array([4147.85, 10950.93, 66755.67]) * 0.2 * 0.02
COEFFS = np.array([1117314.935, 28692377.5, 2748842283.0]) * (0.2 * 0.02) ** 2


def oracle_photoi(tree, rhs_ids, rhs, max_fmg, max_rel):
    ids = all_ids(tree)
    bc = W.bc_table(tree, M.photoi_helmh_bc)
    leaves = np.array([tree.children[i][0] == 0 for i in ids])
    photo = np.zeros((len(ids), tree.box_len))
    ncyc, res = [], []
    max_rhs = None
    for lam, c in zip(LAMBDAS, COEFFS):
        orc = Oracle(tree, helmholtz_lambda=lam ** 2, prolongation_type=M.MG_PROLONG_LINEAR)
        orc.set_bc(bc)
        orc.mg_init()
        orc.set_cc(M.I_RHS, rhs_ids, rhs)
        if max_rhs is None:
            max_rhs = max(orc.maxabs(M.I_RHS), np.sqrt(np.finfo(float).eps))
        n = 0
        for n in range(1, max_fmg + 1):
            orc.fas_fmg(True, True)
            r = orc.maxabs(M.I_TMP)
            if r / max_rhs < max_rel:
                break
        ncyc.append(n)
        res.append(r)
        phi = orc.get_cc(M.I_PHI, ids)
        photo[leaves] = photo[leaves] - c * phi[leaves]
    return photo, np.array(ncyc), np.array(res), leaves


@pytest.mark.parametrize("name,mk", [
    ("corner_nc8_3d", lambda: T.corner_refined_tree(3, 8, 8, 4)),
    ("uniform_nc16_3d", lambda: T.uniform_tree(3, 16, 16, 2)),
    ("cyl_nc8_2d", lambda: T.build_tree(2, 8, [8, 8], 5, lambda l, ix, c: (c[:, 0] < 1.5 * 0.5 ** (l - 1)) &
                                       (np.abs(c[:, 1] - 0.5) < 0.3), coord_t=T.AF_CYL)),
])
@pytest.mark.parametrize("max_rel", [1e-2, 1e-7])
def test_photoi_helmh_compute_matches_oracle(name, mk, max_rel):
    tree = mk()
    ids = all_ids(tree)
    rhs_ids, rhs = W.random_rhs_on_leaves(tree)
    rhs = rhs * 1.0e3
    want, ncyc_o, res_o, leaves = oracle_photoi(tree, rhs_ids, rhs, 10, max_rel)
    bc = W.bc_table(tree, M.photoi_helmh_bc)
    mgs = []
    for lam in LAMBDAS:
        mg = M.mg_t(sides_bc=bc, helmholtz_lambda=lam ** 2, prolongation_type=M.MG_PROLONG_LINEAR)
        M.mg_init(tree, mg)
        mgs.append(mg)
    mgs[0].set_cc(M.I_RHS, rhs_ids, rhs)
    ncyc, res = M.photoi_helmh_compute(tree, mgs, COEFFS, 10, max_rel)
    assert list(ncyc) == list(ncyc_o), (ncyc, ncyc_o)
    assert np.all(np.abs(res - res_o) <= 1e-6 * res_o + 1e-9 * np.abs(rhs).max()), (res, res_o)
    got = mgs[0].get_cc(M.I_PHOTO, ids).reshape(len(ids), -1)
    scale = np.max(np.abs(want))
    assert scale > 0
    assert np.max(np.abs(got - want)) <= 1e-10 * scale, np.max(np.abs(got - want)) / scale
    assert np.all(got[~leaves] == 0.0)  # only leaves are accumulated (:192-201)
    # a second call starts from the previous modes (have_guess = T) and must reproduce the oracle's second call too
    ncyc2, _ = M.photoi_helmh_compute(tree, mgs, COEFFS, 10, max_rel)
    assert np.all(ncyc2 <= ncyc)
    for mg in mgs:
        M.mg_destroy(mg)
