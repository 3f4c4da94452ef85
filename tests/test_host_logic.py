"""CPU: host-side logic (tree builder conventions, device layout index maps) and the C ABI
surface (library loads, exports every symbol include/afmg.h declares, fails loudly without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from afivo_streamer_b200 import _lib
from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "afmg.h")).read()
    declared = set(re.findall(r"\b(afmg_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"afmg_handle", "afmg_opts", "afmg_tree"}
    assert declared, "no declarations found"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in afmg.h but not exported by libafmg.so"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


@pytest.mark.parametrize("ndim,nc", [(3, 4), (3, 8), (3, 16), (2, 8), (2, 16)])
def test_device_layout_is_a_bijection(ndim, nc):
    L = _lib.lib()
    n = nc + 2
    rng = range(n)
    offs = [L.afmg_layout_offset(ndim, nc, i, j, k) for k in (rng if ndim == 3 else [0]) for j in rng for i in rng]
    assert sorted(offs) == list(range(L.afmg_layout_box_len(ndim, nc)))


def test_device_layout_colour_blocks():
    """interior cells of colour (i+j+k)&1 fill the first nc^3/2 entries of their colour block, rows of
    nc/2 cells contiguous in i (DESIGN.md data layout)."""
    L = _lib.lib()
    nc = 8
    col = nc * nc * nc // 2 + 6 * nc * nc // 2
    for (i, j, k) in [(1, 1, 1), (2, 1, 1), (3, 1, 1), (1, 2, 1), (8, 8, 8), (7, 8, 8)]:
        off = L.afmg_layout_offset(3, nc, i, j, k)
        c = (i + j + k) & 1
        assert off == c * col + ((k - 1) * nc + (j - 1)) * (nc // 2) + ((i - 1) >> 1)
    # x-low ghost face cell (0, j, k): colour (j+k)&1, face segment 0
    off = L.afmg_layout_offset(3, nc, 0, 3, 5)
    assert off == ((3 + 5) & 1) * col + nc ** 3 // 2 + (5 - 1) * (nc // 2) + ((3 - 1) >> 1)


def test_create_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    t = T.uniform_tree(3, 8, 8, 2)
    mg = M.mg_t(sides_bc=M.af_bc_dirichlet_zero)
    with pytest.raises(M.AfmgError) as e:
        M.mg_init(t, mg)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_mg_init_requires_sides_bc():
    """mg_init: error stop 'sides_bc not set' (m_af_multigrid.f90:50-51)."""
    with pytest.raises(M.AfmgError):
        M.mg_init(T.uniform_tree(3, 8, 8, 1), M.mg_t())


@pytest.mark.parametrize("ndim", [2, 3])
def test_tree_conventions(ndim):
    t = T.corner_refined_tree(ndim, 8, 8, 4)
    nch = 1 << ndim
    cd = T.child_dix(ndim)
    for id_ in range(1, t.highest_id + 1):
        ch = t.children[id_]
        if ch[0] == 0:
            assert np.all(ch == 0)
            continue
        for c in range(nch):  # ix_c = 2*ix_p - 1 + af_child_dix (m_af_core.f90:1197-1201)
            assert np.array_equal(t.ix[ch[c]], 2 * t.ix[id_] - 1 + cd[c])
            assert t.parent[ch[c]] == id_ and t.lvl[ch[c]] == t.lvl[id_] + 1
    # neighbours are mutual, reversed direction (af_neighb_rev)
    for id_ in range(1, t.highest_id + 1):
        for nb in range(2 * ndim):
            o = t.neighbors[id_, nb]
            if o > 0:
                assert t.neighbors[o, nb ^ 1] == id_
            elif o == 0:  # refinement boundary: the parent's neighbour exists (2:1 balance)
                assert t.neighbors[t.parent[id_], nb] > 0
    # level lists: children of parents in list order
    for l in range(1, t.highest_lvl):
        par = t.parents(l)
        assert np.array_equal(t.lvl_ids[l], t.children[par].reshape(-1))
    centre = 3 ** ndim // 2
    assert np.array_equal(t.neighbor_mat[1:, centre], np.arange(1, t.highest_id + 1))


def test_bc_table_matches_face_coordinates():
    t = T.uniform_tree(3, 8, 8, 2)
    bc = W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, c[..., 0] + 10 * c[..., 1] + 100 * c[..., 2]))
    # box 2 (first child, low corner), face lowx (nb=1): x = 0, y,z at cell centres, y fastest
    row = [q for q in range(len(bc.ids)) if bc.ids[q] == 2 and bc.nbs[q] == 1][0]
    dr = t.dr[2, 0]
    expect = np.array([[10 * (a + 0.5) * dr + 100 * (b + 0.5) * dr for a in range(8)] for b in range(8)]).reshape(-1)
    assert np.allclose(bc.vals[row], expect)


def test_cell_update_count_matches_definition():
    import bench
    t = T.uniform_tree(3, 16, 16, 5)
    assert bench.cell_updates_vcycle(t) == 4 * 16 ** 3 * (8 + 64 + 512 + 4096)


def test_morton_key_matches_the_reference_answers():
    """afivo/tests/answers/test_morton_3d: index (15, 2047, 2047) -> 7362801663; test_morton_2d: (15, 2047) ->
    2796287 (x in the lowest bit).  The slot order of a level and the multi-GPU cuts are built on this key."""
    L = _lib.lib()
    assert L.afmg_morton_key(3, 15, 2047, 2047) == 7362801663
    assert L.afmg_morton_key(2, 15, 2047, 0) == 2796287
    # Z-order: the eight children of a box (2 ix - 1 + af_child_dix) are consecutive, in af_child_dix order
    base = L.afmg_morton_key(3, 2 * 5, 2 * 3, 2 * 9)
    keys = [L.afmg_morton_key(3, 2 * 5 + (c & 1), 2 * 3 + ((c >> 1) & 1), 2 * 9 + ((c >> 2) & 1)) for c in range(8)]
    assert keys == [base + c for c in range(8)]


@pytest.mark.parametrize("ndim", [2, 3])
def test_tree_builder_reproduces_the_reference_refinement_counts_in_a_periodic_domain(ndim):
    """afivo/tests/test_reduction.f90 + answers/test_reduction_2d, _3d: a fully periodic domain of length 2 pi, one
    coarse box of 8^D, refined where all(r_min < 0.4); the 2:1 balance reaches across the periodic boundary (the
    low-side neighbours of the corner box are the wrapped high-side boxes).  highest_id before the i-th
    af_adjust_refinement and max(sum(box%ix)) after it are the reference's printed answers."""
    from afivo_streamer_b200 import tree as T
    highest_id = {3: [1, 9, 17, 49, 105, 273, 713, 2033, 8545], 2: [1, 5, 9, 21, 37, 81, 153, 281, 649, 1881]}[ndim]
    max_ix_sum = {3: [6, 6, 8, 12, 20, 40, 76, 148], 2: [4, 4, 6, 10, 18, 36, 70, 138, 274]}[ndim]
    dlen = 2 * np.arccos(-1.0)

    def refine(l, ixs, ctr):
        r_min = (ixs - 1) * (dlen / 2 ** (l - 1))
        return np.all(r_min < 0.4, axis=1) & (l < 10)

    for lvl, want in enumerate(highest_id, start=1):
        t = T.build_tree(ndim, 8, [8] * ndim, lvl, refine, r_max=[dlen] * ndim, periodic=[True] * ndim)
        assert t.highest_id == want, (lvl, t.highest_id, want)
        ids = np.concatenate(t.lvl_ids)
        assert t.ix[ids].sum(axis=1).min() == ndim  # the reference's "min" column: 3 (2 in 2D)
        if lvl >= 2:
            assert t.ix[ids].sum(axis=1).max() == max_ix_sum[lvl - 2], lvl


def test_mg_use_checks_initialisation_like_the_reference():
    from afivo_streamer_b200 import mg as M
    from afivo_streamer_b200 import tree as T
    from afivo_streamer_b200 import workloads as W
    t = T.uniform_tree(3, 8, 8, 2)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(t))
    with pytest.raises(_lib.AfmgError, match="initialized is false"):  # m_af_multigrid.f90:122
        M.mg_use(t, mg)


def test_photoi_helmh_parameter_sets_are_the_references():
    """src/m_photoi_helmh.f90:80-136: Luque scales by (frac_O2 / 0.2) p, Bourdon by frac_O2 p, coefficients by the
    square; the reference's error stops become exceptions."""
    from afivo_streamer_b200 import mg as M
    lam, cf = M.photoi_helmh_parameters("Bourdon-3")
    assert np.allclose(lam, np.array([4147.85, 10950.93, 66755.67]) * 0.2, rtol=1e-15)
    assert np.allclose(cf, np.array([1117314.935, 28692377.5, 2748842283.0]) * 0.04, rtol=1e-15)
    lam, cf = M.photoi_helmh_parameters("Bourdon-2", frac_O2=0.1, gas_pressure=0.5)
    assert np.allclose(lam, np.array([7305.62, 44081.25]) * 0.05) and np.allclose(cf, np.array([11814508.38, 998607256.0]) * 0.0025)
    lam, cf = M.photoi_helmh_parameters("Luque", frac_O2=0.2, gas_pressure=2.0)
    assert np.allclose(lam, np.array([4425.38, 750.06]) * 2) and np.allclose(cf, np.array([337557.38, 19972.14]) * 4)
    lam, cf = M.photoi_helmh_parameters("custom", gas_pressure=0.5, lambdas=[100.0], coeffs=[8.0])
    assert lam[0] == 50.0 and cf[0] == 2.0
    for bad in (dict(author="Luque", eta=0.5), dict(author="Bourdon-3", frac_O2=0.0), dict(author="custom"),
                dict(author="Zheleznyak")):
        with pytest.raises(ValueError):
            M.photoi_helmh_parameters(**bad)


def test_field_boundary_condition_callbacks_of_the_streamer_code():
    """field_bc_homogeneous / _neumann / _all_neumann / _all_dirichlet (src/m_field.f90:590-669) as rows of afmg_set_bc."""
    from afivo_streamer_b200 import tree as T
    from afivo_streamer_b200 import workloads as W
    t3 = T.build_tree(3, 8, [8, 8, 16], 2, None, r_max=[1.0, 1.0, 2.0])
    t2 = T.build_tree(2, 8, [8, 8], 2, None, coord_t=T.AF_CYL)

    def rows(bc, nb):
        m = bc.nbs == nb
        return set(bc.types[m]), set(np.unique(bc.vals[m]))

    h = W.bc_field_homogeneous(t3, 5.0)
    assert rows(h, 5) == ({W.AF_BC_DIRICHLET}, {0.0}) and rows(h, 6) == ({W.AF_BC_DIRICHLET}, {5.0})
    assert all(rows(h, nb) == ({W.AF_BC_NEUMANN}, {0.0}) for nb in (1, 2, 3, 4))
    n = W.bc_field_neumann(t3, 5.0)
    assert rows(n, 5) == ({W.AF_BC_DIRICHLET}, {0.0}) and rows(n, 6) == ({W.AF_BC_NEUMANN}, {2.5})  # voltage / domain_len(3)
    a = W.bc_field_all_neumann(t3)
    assert set(a.types) == {W.AF_BC_NEUMANN} and not a.vals.any()
    d3, d2 = W.bc_field_all_dirichlet(t3), W.bc_field_all_dirichlet(t2)
    assert set(d3.types) == {W.AF_BC_DIRICHLET}
    assert rows(d2, 1) == ({W.AF_BC_NEUMANN}, {0.0}) and all(rows(d2, nb)[0] == {W.AF_BC_DIRICHLET} for nb in (2, 3, 4))


def test_field_residual_threshold_is_the_references_formula():
    """src/m_field.f90:467-480."""
    from afivo_streamer_b200 import mg as M
    from afivo_streamer_b200 import tree as T
    t = T.build_tree(3, 8, [8, 8, 16], 3, None, r_max=[1e-2, 1e-2, 2e-2])  # min dr = 1.25e-3 / 4
    min_dr = 1.25e-3 / 4
    assert M.field_residual_threshold(t, 0.0, 0.0) == 1e-6                                  # min_residual
    assert M.field_residual_threshold(t, 1e10, 0.0) == 1e10 * 1e-4                          # rhs term
    assert M.field_residual_threshold(t, 0.0, -4e4) == pytest.approx(1e-10 * 4e4 / (2e-2 * min_dr), rel=1e-14)
    assert M.field_residual_threshold(t, 0.0, 4e4, use_electrode=True) == pytest.approx(1e-8 * 4e4 / (2e-2 * min_dr), rel=1e-14)
