"""CPU: the C++ host-side mirror (include/afmg.hpp) compiles with -pedantic -Werror and builds the same tree,
array by array, as the Python builder that follows afivo's conventions (ids, ix, parent, children in af_child_dix
order, neighbors, neighbor_mat, level lists)."""
import os
import subprocess

import numpy as np
import pytest

from afivo_streamer_b200 import tree as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nc,coarse,lvl", [(8, 8, 3), (4, 8, 3), (16, 16, 2)])
def test_cpp_tree_matches_python_builder(tmp_path, nc, coarse, lvl):
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    exe = str(tmp_path / "cpp_tree_dump")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_tree_dump.cpp"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, str(nc), str(coarse), str(lvl)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    t = T.uniform_tree(3, nc, coarse, lvl)
    hl, hid = map(int, lines[0].split())
    assert (hl, hid) == (t.highest_lvl, t.highest_id)
    for l in range(hl):
        vals = list(map(int, lines[1 + l].split()[1:]))
        assert vals[0] == len(t.lvl_ids[l]) and vals[1:] == list(map(int, t.lvl_ids[l]))
    boxes = np.array([list(map(int, ln.split()[1:])) for ln in lines[1 + hl:1 + hl + hid]])
    ids = np.arange(1, hid + 1)
    assert np.array_equal(boxes[:, 0], t.lvl[ids])
    assert np.array_equal(boxes[:, 1:4], t.ix[ids])
    assert np.array_equal(boxes[:, 4], t.parent[ids])
    assert np.array_equal(boxes[:, 5:13], t.children[ids])
    assert np.array_equal(boxes[:, 13:19], t.neighbors[ids])
    assert np.array_equal(boxes[:, 19:46], t.neighbor_mat[ids])
    assert lines[-1] in ("device present", "error -2")  # AFMG_ERR_CUDA without a GPU: no CPU fallback
