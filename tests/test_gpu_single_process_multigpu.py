"""-m gpu, needs >= 2 visible devices: afmg_opts.n_gpus -- ONE process, one handle, N GPUs (one host thread per GPU
inside the library, peer access instead of CUDA IPC).  This is the mode a single-process caller such as the reference
(one OpenMP process, afivo/documentation/parallelization.md) can use without torchrun / MPI.  Every result must be
bit-identical to the single-GPU solve, as in the multi-process mode (tools/mgpu_check.py)."""
import os

import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

from util import all_ids, bc_mixed

pytestmark = pytest.mark.gpu


def n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(n_devices() < 2, reason="needs two GPUs on one node")


@pytest.fixture(autouse=True)
def split_small_levels():
    """the test trees are small: let the partition split every level so that halos really cross GPUs"""
    old = os.environ.get("AFMG_MIN_SPLIT_BOXES")
    os.environ["AFMG_MIN_SPLIT_BOXES"] = "8"
    yield
    os.environ.pop("AFMG_MIN_SPLIT_BOXES", None)
    if old is not None:
        os.environ["AFMG_MIN_SPLIT_BOXES"] = old


def solve(tree, bc, ids, rhs, n_gpus, **opts):
    mg = M.mg_t(sides_bc=bc, n_gpus=n_gpus, device=0, **opts)
    M.mg_init(tree, mg)
    mg.set_cc(M.I_RHS, ids, rhs)
    hist = []
    M.mg_fas_fmg(tree, mg, True, False)
    hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(3):
        M.mg_fas_vcycle(tree, mg, True)
        hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    boxes = all_ids(tree)
    out = {"phi": mg.get_cc(M.I_PHI, boxes), "tmp": mg.get_cc(M.I_TMP, boxes), "sum": M.af_tree_sum_cc(tree, mg, M.I_PHI),
           "checksum": mg.checksum(M.I_PHI), "inner": mg.get_cc_interior(M.I_PHI, boxes)}
    M.field_from_potential(tree, mg, -1.0)
    out["fld"] = mg.get_cc(M.I_FLD, boxes)
    owners = mg.owners(boxes)
    M.mg_destroy(mg)
    return np.array(hist), out, owners


@needs2
@pytest.mark.parametrize("name", ["multibox_nc8", "shell_nc8", "uniform_nc16_l3", "channel_nc8"])
def test_one_process_two_gpus_bit_identical(name):
    tree = {"multibox_nc8": lambda: T.build_tree(3, 8, [16, 8, 24], 3, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45),
            "shell_nc8": lambda: T.shell_tree(8, 8, 4, 0.35),
            "uniform_nc16_l3": lambda: T.uniform_tree(3, 16, 16, 3), "channel_nc8": lambda: T.channel_tree(8, 8, 6, 3)}[name]()
    bc = W.bc_table(tree, bc_mixed)
    ids, rhs = W.random_rhs_on_leaves(tree)
    h1, o1, _ = solve(tree, bc, ids, rhs, 0)
    h2, o2, owners = solve(tree, bc, ids, rhs, 2)
    assert set(np.unique(owners)) == {0, 1}, "both GPUs must own boxes"
    assert np.array_equal(h1, h2)
    for k in ("phi", "tmp", "inner", "fld"):
        assert np.array_equal(o1[k], o2[k]), k
    assert o1["sum"] == o2["sum"] and o1["checksum"] == o2["checksum"]


@needs2
def test_one_process_two_gpus_helmholtz_modes():
    lambdas = np.array([4147.85, 10950.93, 66755.67]) * 0.2 * 0.02
    coeffs = np.array([1117314.935, 28692377.5, 2748842283.0]) * (0.2 * 0.02) ** 2
    tree = T.uniform_tree(3, 8, 8, 4)
    bc = W.bc_table(tree, M.photoi_helmh_bc)
    ids, rhs = W.random_rhs_on_leaves(tree)
    boxes = all_ids(tree)
    res = []
    for ng in (0, 2):
        mgs = []
        for lam in lambdas:
            mg = M.mg_t(sides_bc=bc, n_gpus=ng, device=0, helmholtz_lambda=lam ** 2, prolongation_type=M.MG_PROLONG_LINEAR)
            M.mg_init(tree, mg)
            mgs.append(mg)
        mgs[0].set_cc(M.I_RHS, ids, rhs * 1.0e3)
        ncyc, r = M.photoi_helmh_compute(tree, mgs, coeffs, 10, 1.0e-2)
        res.append((list(ncyc), np.array(r), mgs[0].get_cc(M.I_PHOTO, boxes)))
        for mg in mgs:
            M.mg_destroy(mg)
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2]) and np.abs(res[1][2]).max() > 0


@needs2
def test_each_gpu_maps_only_the_boxes_it_owns():
    """the slot space is reserved as virtual addresses; physical memory sits under the owned slot ranges only, so two
    GPUs hold about half of the 256^3 tree each (2 MB granularity) -- and still solve it bit-identically"""
    tree = T.uniform_tree(3, 16, 16, 5)
    bc = W.bc_dirichlet_zero(tree)
    ids, rhs = W.constant_rhs_on_leaves(tree, 1.0)
    out = []
    for ng in (0, 2):
        mg = M.mg_t(sides_bc=bc, n_gpus=ng, device=0)
        M.mg_init(tree, mg)
        mapped, full = mg.slab_bytes()
        if ng == 2:
            assert len(mapped) == 2 and np.all(mapped < 0.62 * full), (mapped, full)
            assert mapped.sum() < 1.1 * full[0]
        else:
            assert mapped[0] == full[0]
        mg.set_cc(M.I_RHS, ids, rhs)
        M.mg_fas_fmg(tree, mg, True, False)
        M.mg_fas_vcycle(tree, mg, True)
        out.append((M.af_tree_maxabs_cc(tree, mg, M.I_TMP), mg.checksum(M.I_PHI)))
        mg.clear(M.I_TMP)  # afmg_clear touches the owned records only
        M.mg_destroy(mg)
    assert out[0] == out[1]


def test_n_gpus_beyond_the_visible_devices_is_an_error_not_a_fallback():
    tree = T.uniform_tree(3, 8, 8, 2)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree), n_gpus=n_devices() + 1 if n_devices() < 8 else 9, device=0)
    with pytest.raises(M.AfmgError):
        M.mg_init(tree, mg)
