"""Solve the Poisson problem stored in an afivo .dat file (af_write_tree, version 3) on the GPU:
    python tools/solve_dat.py sim_000100.dat [--ndim 3] [--phi phi] [--rhs rhs] [--eps eps] [--cycles 5]
Prints the residual max-norm per cycle the way field_compute tests it (src/m_field.f90:491-524) and the
maximum of the field norm computed on the device."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afivo_streamer_b200 import datfile as D  # noqa: E402
from afivo_streamer_b200 import mg as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--ndim", type=int, default=None)
    ap.add_argument("--phi", default="phi")
    ap.add_argument("--rhs", default="rhs")
    ap.add_argument("--eps", default=None)
    ap.add_argument("--cycles", type=int, default=5)
    ap.add_argument("--lsf-boundary-value", type=float, default=0.0)
    a = ap.parse_args()
    dat = D.read_tree(a.file, a.ndim)
    t = dat.tree
    print(f"{a.file}: NDIM={dat.ndim} n_cell={t.nc} levels={t.highest_lvl} boxes={t.n_boxes} variables={dat.cc_names}")
    tree, mg = M.mg_from_dat(dat, phi=a.phi, rhs=a.rhs, eps=a.eps, lsf_boundary_value=a.lsf_boundary_value)
    M.mg_fas_fmg(tree, mg, True, True)
    print(f"FMG      residual {M.af_tree_maxabs_cc(tree, mg, M.I_TMP):.6e}")
    for i in range(a.cycles):
        M.mg_fas_vcycle(tree, mg, True)
        print(f"V-cycle {i + 1} residual {M.af_tree_maxabs_cc(tree, mg, M.I_TMP):.6e}")
    M.field_from_potential(tree, mg, -1.0)
    ids = dat.ids_in_use()
    leaves = ids[~tree.has_children(ids)]
    print(f"max |E| on leaves {np.max(mg.get_cc(M.I_FLD, leaves)):.6e}")
    M.mg_destroy(mg)


if __name__ == "__main__":
    main()
