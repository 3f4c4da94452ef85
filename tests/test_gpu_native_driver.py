"""-m gpu: the reference's own benchmark program (afivo/examples/poisson_benchmark.f90) as a native C++
executable on top of the C ABI (tools/poisson_benchmark.cpp): same command line and output, the tree built in
the reference's conventions without any Python in the loop."""
import os
import re
import subprocess

import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_poisson_benchmark_native_driver():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    out = subprocess.run([os.path.join(ROOT, "tools", "poisson_benchmark_3d"), "8", "8", "3", "0.05"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    txt = out.stdout
    assert "Running poisson_benchmark_3d" in txt and "Max refinement lvl:  3" in txt
    assert int(re.search(r"Number of boxes used:\s+(\d+)", txt).group(1)) == 73  # 1 + 8 + 64
    per_it = float(re.search(r"Per iteration:\s+([0-9.E+-]+) seconds", txt).group(1))
    assert 0 < per_it < 0.05
    res = float(re.search(r"Residual max-norm after one more V-cycle:\s+([0-9.E+-]+)", txt).group(1))
    # the same problem through the Python mirror, iterated to its fixed point: the native driver's tree, boundary
    # conditions and right-hand side must describe the same system (its residual sits at the same rounding floor)
    tree = T.uniform_tree(3, 8, 8, 3)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, mg)
    ids, rhs = W.constant_rhs_on_leaves(tree, 1.0)
    mg.set_cc(M.I_RHS, ids, rhs)
    M.mg_fas_fmg(tree, mg, False, False)
    for _ in range(30):
        M.mg_fas_fmg(tree, mg, False, True)
    M.mg_fas_vcycle(tree, mg, True)
    floor = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
    M.mg_destroy(mg)
    assert res < 1e-9 and floor < 1e-9 and res < 50 * floor + 1e-13, (res, floor)


def test_poisson_basic_native_driver_matches_oracle():
    """afivo/examples/poisson_basic.f90 as a native executable (tools/poisson_basic.cpp): the adaptively refined tree
    (3 x 1 x 1 domain, refinement where dr^2 |rhs| > 1e-3, 139 boxes), Dirichlet values from the analytic solution and
    the per-cycle table "max residual | max error" must agree with the oracle run on the tree of the Python builder."""
    from oracle.oracle import Oracle
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s"])
    out = subprocess.run([os.path.join(ROOT, "tools", "poisson_basic_3d")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    txt = out.stdout
    rows = re.findall(r"^\s*(\d+)\s+([0-9.]+E[+-]\d+)\s*([0-9.]+E[+-]\d+)\s*$", txt, flags=re.M)
    assert len(rows) == 10, txt
    got_res = np.array([float(r[1]) for r in rows])
    got_err = np.array([float(r[2]) for r in rows])

    g = W.Gaussians([[0.1, 0.1, 0.1], [0.75, 0.75, 0.75]], 0.04)
    nc = 16

    def refine(l, ixs, ctr):
        dr = 3.0 / 48 * 0.5 ** (l - 1)
        off = (np.arange(nc) - (nc - 1) / 2) * dr
        gz, gy, gx = np.meshgrid(off, off, off, indexing="ij")
        pts = ctr[:, None, :] + np.stack([gx, gy, gz], axis=-1).reshape(1, -1, 3)
        return dr * dr * np.max(np.abs(g.laplacian(pts)), axis=1) > 1e-3

    t = T.build_tree(3, nc, [48, 16, 16], 4, refine, r_max=[3.0, 1.0, 1.0])
    assert int(re.search(r"Number of boxes used:\s+(\d+)", txt).group(1)) == t.n_boxes == 139
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_function(t, g.value))
    o.mg_init()
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    ctr = W.cell_centres(t, leaves, ghosts=True)
    o.set_cc(M.I_RHS, leaves, g.laplacian(ctr))
    res, err = [], []
    for it in range(10):
        o.fas_fmg(True, it > 0)
        res.append(o.maxabs(M.I_TMP))
        phi = o.get_cc(M.I_PHI, leaves).reshape(ctr.shape[:-1])
        err.append(np.max(np.abs(phi - g.value(ctr))[W.interior(t)]))
    res, err = np.array(res), np.array(err)
    # printed with 6 significant digits; the residual rhs - L(phi) itself is only defined up to the rounding of its
    # seven terms of size |phi| / dr^2 (finest dr = 3 / 48 / 8)
    floor = 16 * np.finfo(float).eps * 7 / (3.0 / 48 / 8) ** 2
    assert np.all(np.abs(got_res - res) <= 2e-5 * res + floor), (got_res, res, floor)
    assert np.all(np.abs(got_err - err) <= 2e-5 * err), (got_err, err)
    # the known behaviour of the example: monotone residual, error plateau at the discretisation level
    assert np.all(got_res[1:7] < 0.1 * got_res[:6]) and abs(got_err[-1] - got_err[-2]) < 1e-6 * got_err[-1]
