"""CPU: property tests (hypothesis) of the library's pure host functions: the multi-GPU partition rule, the Morton
key, and invariants of the stencil builders that hold for any permittivity / distance field."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from afivo_streamer_b200 import _lib
from afivo_streamer_b200 import mg as M


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 8), st.lists(st.integers(0, 300), min_size=1, max_size=9))
def test_partition_covers_every_level_once_and_keeps_sibling_groups_together(n_ranks, groups):
    counts = [1] + [8 * g for g in groups]  # level 1: the coarse box; finer levels: whole sibling groups
    cuts = M.partition(n_ranks, counts)
    assert cuts.shape == (len(counts), n_ranks + 1)
    for l, n in enumerate(counts):
        c = cuts[l]
        assert c[0] == 0 and c[-1] == n and np.all(np.diff(c) >= 0)
        if l == 0:
            assert c[1] == n  # the coarse grid is rank 0's
        else:
            assert np.all(c % 8 == 0)  # cuts at sibling-group boundaries
            sizes = np.diff(c) // 8
            assert sizes.max() - sizes.min() <= 1  # balanced to one group


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2 ** 20 - 1), st.integers(0, 2 ** 20 - 1), st.integers(0, 2 ** 20 - 1))
def test_morton_key_interleaves_bits_with_x_lowest(x, y, z):
    L = _lib.lib()
    k3 = L.afmg_morton_key(3, x, y, z)
    k2 = L.afmg_morton_key(2, x, y, 0)
    for b in range(20):
        assert (k3 >> (3 * b)) & 1 == (x >> b) & 1 and (k3 >> (3 * b + 1)) & 1 == (y >> b) & 1
        assert (k3 >> (3 * b + 2)) & 1 == (z >> b) & 1
        assert (k2 >> (2 * b)) & 1 == (x >> b) & 1 and (k2 >> (2 * b + 1)) & 1 == (y >> b) & 1
    # children of one parent are consecutive: the key of (2x + a, 2y + b, 2z + c) is 8 * key(x, y, z) + (a | b<<1 | c<<2)
    if max(x, y, z) < 2 ** 19:
        assert L.afmg_morton_key(3, 2 * x + 1, 2 * y, 2 * z + 1) == 8 * k3 + 5


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 3), st.sampled_from([4, 8]), st.integers(0, 2 ** 31 - 1), st.floats(2.0, 50.0))
def test_permittivity_stencil_invariants(nd, nc, seed, contrast):
    """mg_box_lpld_stencil for any positive eps: off-diagonal weights positive and between the two face-adjacent
    idr2 * eps values (harmonic mean), the row sums to zero, symmetric across a shared face."""
    L = _lib.lib()
    rng = np.random.default_rng(seed)
    eps = 1.0 + (contrast - 1.0) * rng.random((nc + 2,) * nd)
    dr = np.full(nd, 1.0 / nc)
    ncell = nc ** nd
    v, f = np.zeros((ncell, 2 * nd + 1)), np.zeros(ncell)
    stype, has_f, cyl = C.c_int32(), C.c_int32(), C.c_int32()
    rc = L.afmg_build_box_operator(nd, nc, 1, 2, _dp(dr), None, _dp(np.ascontiguousarray(eps).reshape(-1)), None,
                                   _dp(v), _dp(f), C.byref(stype), C.byref(has_f), C.byref(cyl))
    assert rc == 0 and stype.value == 2 and not has_f.value
    assert np.all(v[:, 1:] > 0)
    assert np.max(np.abs(v.sum(axis=1))) <= 1e-12 * np.max(np.abs(v[:, 0]))
    vv = v.reshape((nc,) * nd + (2 * nd + 1,))
    inner = (slice(1, -1),) * nd
    idr2 = nc * nc
    for d in range(nd):  # array axes are (k, j, i): dimension d is axis nd - 1 - d
        ax = nd - 1 - d
        lo = [slice(1, -1)] * nd
        hi = [slice(1, -1)] * nd
        lo[ax], hi[ax] = slice(0, -2), slice(2, None)
        e0, em, ep = eps[inner], eps[tuple(lo)], eps[tuple(hi)]
        for w, other in ((vv[..., 1 + 2 * d], em), (vv[..., 2 + 2 * d], ep)):
            assert np.all(w >= idr2 * np.minimum(e0, other) * (1 - 1e-12))
            assert np.all(w <= idr2 * np.maximum(e0, other) * (1 + 1e-12))
        # the weight a cell gives its high neighbour equals the weight that neighbour gives back
        a = np.moveaxis(vv[..., 2 + 2 * d], ax, 0)[:-1]
        b = np.moveaxis(vv[..., 1 + 2 * d], ax, 0)[1:]
        assert np.array_equal(a, b)


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 3), st.integers(0, 2 ** 31 - 1))
def test_level_set_stencil_invariants(nd, seed):
    """mg_box_lsf_stencil for any distances in (0, 1]: rows sum to f (the weight moved to the right-hand side),
    f <= 0, entries towards a boundary are zero, and cells without a boundary carry the plain Laplacian."""
    L = _lib.lib()
    nc = 4
    rng = np.random.default_rng(seed)
    ncell = nc ** nd
    dd = np.where(rng.random((ncell, 2 * nd)) < 0.3, rng.uniform(1e-4, 0.999, (ncell, 2 * nd)), 1.0)
    dr = np.full(nd, 0.25)
    v, f = np.zeros((ncell, 2 * nd + 1)), np.zeros(ncell)
    stype, has_f, cyl = C.c_int32(), C.c_int32(), C.c_int32()
    rc = L.afmg_build_box_operator(nd, nc, 1, 1, _dp(dr), None, None, _dp(dd.reshape(-1)), _dp(v), _dp(f),
                                   C.byref(stype), C.byref(has_f), C.byref(cyl))
    assert rc == 0 and stype.value == 2 and has_f.value and not cyl.value
    assert np.all(f <= 0) and np.all(v[:, 1:][dd < 1] == 0) and np.all(v[:, 1:][dd >= 1] > 0)
    assert np.max(np.abs(v.sum(axis=1) - f)) <= 1e-11 * np.max(np.abs(v[:, 0]))
    plain = np.all(dd >= 1, axis=1)
    if plain.any():
        assert np.allclose(v[plain, 1:], 16.0, rtol=1e-15) and np.allclose(v[plain, 0], -2 * nd * 16.0, rtol=1e-15)
        assert np.all(f[plain] == 0)
