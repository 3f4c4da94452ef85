"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU
oracle): the oracle must keep reproducing them bit for bit (CPU), the CUDA path must match them to
the north-star tolerance 1e-10 (GPU)."""
import glob
import os

import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import workloads as W

from util import TREES, TREES2D, bc_mixed

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"), allow_pickle=True)
    opts = {k: v for k, v in g["opts"]} if g["opts"].size else {}
    return g, opts


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    g, opts = _load(name)
    tree, hist, leaves, phi = make_golden.run_case(name, opts)
    assert tree.n_boxes == int(g["n_boxes"])
    assert np.array_equal(hist, g["residual_history"])
    sel = np.isin(leaves, g["leaf_ids"])
    assert np.array_equal(phi[sel], g["phi"])
    assert np.array_equal(phi.sum(axis=1), g["phi_box_sums"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(name):
    g, opts = _load(name)
    tree = (TREES.get(name) or TREES2D[name])()
    mg = M.mg_t(sides_bc=W.bc_table(tree, bc_mixed), **opts)
    M.mg_init(tree, mg)
    ids, rhs = W.random_rhs_on_leaves(tree)
    mg.set_cc(M.I_RHS, ids, rhs)
    hist = []
    M.mg_fas_fmg(tree, mg, True, False)
    hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(len(g["residual_history"]) - 1):
        M.mg_fas_vcycle(tree, mg, True)
        hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    ref = g["residual_history"]
    assert np.all(np.abs(np.array(hist) - ref) <= 1e-9 * ref[0] + 1e-6 * ref)
    phi = mg.get_cc(M.I_PHI, g["leaf_ids"]).reshape(len(g["leaf_ids"]), -1)
    assert np.max(np.abs(phi - g["phi"])) <= 1e-10 * float(g["phi_maxabs"])
    M.mg_destroy(mg)
