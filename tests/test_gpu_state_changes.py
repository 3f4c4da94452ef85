"""-m gpu: things that change BETWEEN solves while the cached CUDA graphs of the cycles stay alive (the time loop of
the streamer code: the applied voltage changes every step, src/m_field.f90:481-487, 590-610), plus the data-path
helpers added for the end-to-end path (interior download, page-locked buffers, bitwise checksum)."""
import ctypes as C

import numpy as np
import pytest

from afivo_streamer_b200 import _lib
from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from test_gpu_stencils import lsf_sphere, make_pair
from util import all_ids, assert_same_state

pytestmark = pytest.mark.gpu


def test_lsf_boundary_value_changes_between_cached_cycles():
    """mg%lsf_boundary_value = current_voltage before every solve: the device reads it from memory, so the graphs
    captured for the first solve must give the right answer for the later values."""
    tree = T.uniform_tree(3, 8, 8, 3)
    orc, mg, _ = make_pair(tree, lsf=lsf_sphere, lsf_boundary_value=1.0)
    for volt in (1.0, 2.5, -0.75, 2.5):
        orc.set_opts(lsf_boundary_value=volt)
        orc.mg_init()
        mg.set_lsf_boundary_value(volt)
        n0 = mg.kernel_launches()
        orc.fas_fmg(True, True)
        M.mg_fas_fmg(tree, mg, True, True)
        for _ in range(2):
            orc.fas_vcycle(True)
            M.mg_fas_vcycle(tree, mg, True)
        assert mg.kernel_launches() > n0
        ro, rg = orc.maxabs(M.I_TMP), M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        assert abs(ro - rg) <= 1e-6 * ro + 1e-9
        assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)


@pytest.mark.parametrize("ndim", [2, 3])
def test_bc_values_and_types_change_between_cached_cycles(ndim):
    """afmg_set_bc before every solve (field_bc_homogeneous returns the current voltage).  New VALUES keep the graphs;
    new TYPES rebuild the coarse solver, whose buffers are reallocated: the cached graphs must not survive that."""
    tree = T.corner_refined_tree(ndim, 8, 8, 3)
    ids = all_ids(tree)
    rng = np.random.default_rng(5)
    rhs = rng.uniform(-1, 1, (len(ids),) + (tree.nc + 2,) * ndim)

    def bc_a(volt):
        return W.bc_field_homogeneous(tree, volt)

    def bc_b(volt):  # Dirichlet on the x faces instead, Neumann elsewhere
        return W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, volt if nb == 2 else 0.0) if (nb - 1) // 2 == 0
                          else (W.AF_BC_NEUMANN, 0.0))

    orc = Oracle(tree)
    orc.set_bc(bc_a(1.0))
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc_a(1.0))
    M.mg_init(tree, mg)
    orc.set_cc(M.I_RHS, ids, rhs)
    mg.set_cc(M.I_RHS, ids, rhs)
    for bc in (bc_a(1.0), bc_a(3.0), bc_b(2.0), bc_b(-1.0), bc_a(0.5)):
        orc.set_bc(bc)
        orc.mg_init()
        mg.set_bc(bc)
        orc.fas_fmg(True, True)
        M.mg_fas_fmg(tree, mg, True, True)
        for _ in range(3):
            orc.fas_vcycle(True)
            M.mg_fas_vcycle(tree, mg, True)
        ro, rg = orc.maxabs(M.I_TMP), M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        assert abs(ro - rg) <= 1e-6 * ro + 1e-9
        assert_same_state(tree, orc, mg, exact=False, rtol=1e-10, what=("phi",))
    M.mg_destroy(mg)


@pytest.mark.parametrize("ndim,nc", [(3, 16), (3, 8), (2, 8)])
def test_interior_download_and_checksum(ndim, nc):
    tree = T.corner_refined_tree(ndim, nc, nc, 3)
    ids = all_ids(tree)
    rng = np.random.default_rng(9)
    data = rng.uniform(-1, 1, (len(ids),) + (nc + 2,) * ndim)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, mg)
    mg.set_cc(M.I_PHI, ids, data)
    sub = ids[::2]
    got = mg.get_cc_interior(M.I_PHI, sub)
    assert np.array_equal(got, data[::2][W.interior(tree)].reshape(got.shape))
    s, x = mg.checksum(M.I_PHI)
    bits = np.ascontiguousarray(data).view(np.uint64).reshape(-1)
    with np.errstate(over="ignore"):
        assert s == int(np.add.reduce(bits, dtype=np.uint64))
    assert x == int(np.bitwise_xor.reduce(bits))
    M.mg_destroy(mg)


def test_page_locked_buffers_and_chunked_transfers():
    """afmg_host_alloc'ed buffers through upload / download in several chunks (the staging area holds two chunks of
    128 MB: 4681 boxes of 18^3 doubles = 218 MB go up and down in two, with the copy of one chunk overlapping the
    kernel of the other)."""
    tree = T.uniform_tree(3, 16, 16, 5)
    ids = all_ids(tree)
    n = len(ids) * tree.box_len
    L = _lib.lib()
    p_up, p_dn = L.afmg_host_alloc(n * 8), L.afmg_host_alloc(n * 8)
    assert p_up and p_dn
    up = np.ctypeslib.as_array(C.cast(p_up, C.POINTER(C.c_double)), shape=(n,))
    dn = np.ctypeslib.as_array(C.cast(p_dn, C.POINTER(C.c_double)), shape=(n,))
    up[:] = np.random.default_rng(4).uniform(-1, 1, n)
    dn[:] = 0.0
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, mg)
    mg.upload_ptr(M.I_TMP, ids, p_up)
    mg.download_ptr(M.I_TMP, ids, p_dn)
    assert np.array_equal(up, dn)
    inner = np.empty(len(ids) * 16 ** 3)
    mg.download_interior_ptr(M.I_TMP, ids, inner.ctypes.data)
    assert np.array_equal(inner.reshape(len(ids), 16, 16, 16), up.reshape((len(ids), 18, 18, 18))[:, 1:-1, 1:-1, 1:-1])
    M.mg_destroy(mg)
    del up, dn
    L.afmg_host_free(p_up)
    L.afmg_host_free(p_dn)


@pytest.mark.parametrize("ndim,nc", [(3, 16), (3, 8), (2, 8)])
def test_field_set_rhs_on_device_has_the_references_bits(ndim, nc):
    """field_set_rhs (src/m_field.f90:406-444): cc(:, i_rhs) = 0, then += q_n * cc(:, species_n) for n = 1 .. on the
    leaves.  numpy restates the loop (same order, no FMA); the device result must be identical bit for bit, from host
    memory and from densities already resident on the device."""
    import torch
    tree = T.corner_refined_tree(ndim, nc, nc, 3)
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    rng = np.random.default_rng(21)
    fac = -1.602176634e-19 / 8.8541878128e-12  # -UC_elem_charge / UC_eps0
    charges = np.array([1.0, -1.0, 2.0, -1.0]) * fac
    dens = [rng.uniform(0, 1e18, (len(leaves), tree.box_len)) for _ in charges]
    dens[2][:, ::3] = 0.0
    want = np.zeros_like(dens[0])
    for q, d in zip(charges, dens):
        want = want + q * d
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, mg)
    mg.set_cc(M.I_RHS, leaves, rng.uniform(-1, 1, want.shape))  # stale values must not survive
    mg.field_set_rhs(leaves, charges, dens)
    got = mg.get_cc(M.I_RHS, leaves).reshape(want.shape)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    # densities resident on the device
    mg.set_cc(M.I_RHS, leaves, rng.uniform(-1, 1, want.shape))
    dev = [torch.as_tensor(d, device="cuda") for d in dens]
    mg.field_set_rhs(leaves, charges, [int(t.data_ptr()) for t in dev], on_device=True)
    got = mg.get_cc(M.I_RHS, leaves).reshape(want.shape)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    M.mg_destroy(mg)
