"""Generates tests/golden/*.npz from the CPU oracle (oracle/afmg_oracle.cpp).

The Fortran reference cannot be run in this image (no Fortran compiler, Hypre not vendored), so
these fixtures pin the ORACLE: residual histories and potentials of small seeded problems.  They
guard the oracle against regressions (CPU tests) and give the CUDA path a committed target
(-m gpu tests).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from afivo_streamer_b200 import workloads as W  # noqa: E402
from oracle.oracle import I_PHI, I_RHS, I_TMP, Oracle  # noqa: E402
from util import TREES, TREES2D, all_ids, bc_mixed  # noqa: E402

CASES = {
    "corner_nc8_l4": dict(),
    "uniform_nc16_l2": dict(),
    "multibox_coarse_nc8": dict(helmholtz_lambda=250.0),
    "xy2d_uniform_nc8_l4": dict(),
    "cyl2d_corner_nc8_l5": dict(helmholtz_lambda=30.0),
}


def run_case(name, opts, n_v=3):
    tree = (TREES.get(name) or TREES2D[name])()
    bc = W.bc_table(tree, bc_mixed)
    orc = Oracle(tree, **opts)
    orc.set_bc(bc)
    orc.mg_init()
    ids, rhs = W.random_rhs_on_leaves(tree)
    orc.set_cc(I_RHS, ids, rhs)
    hist = []
    orc.fas_fmg(True, False)
    hist.append(orc.maxabs(I_TMP))
    for _ in range(n_v):
        orc.fas_vcycle(True)
        hist.append(orc.maxabs(I_TMP))
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    phi = orc.get_cc(I_PHI, leaves)
    return tree, np.array(hist), leaves, phi


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, opts in CASES.items():
        tree, hist, leaves, phi = run_case(name, opts)
        # keep fixtures small: full potential of the first 8 and last 8 leaves + norms of the rest
        keep = np.concatenate([np.arange(min(8, len(leaves))), np.arange(max(0, len(leaves) - 8), len(leaves))])
        keep = np.unique(keep)
        np.savez_compressed(os.path.join(here, f"{name}.npz"), residual_history=hist, leaf_ids=leaves[keep],
                            phi=phi[keep], phi_box_sums=phi.sum(axis=1), phi_maxabs=np.abs(phi).max(),
                            n_boxes=tree.n_boxes, opts=np.array(sorted(opts.items()), dtype=object))
        print(name, hist)
