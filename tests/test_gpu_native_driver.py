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
