"""CPU: known answers for the oracle's variable-coefficient and level-set paths (SURVEY 8 rows a9, a17,
a22, a25), which the GPU parity tests of tests/test_gpu_stencils.py rely on."""
import numpy as np

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from util import all_ids, bc_mixed


def solve(tree, eps=None, lsf_dd=None, lsf_value=0.0, n_v=3):
    orc = Oracle(tree, with_eps=eps is not None, lsf_boundary_value=lsf_value)
    orc.set_bc(W.bc_table(tree, bc_mixed))
    ids = all_ids(tree)
    if eps is not None:
        orc.set_cc(M.I_EPS, ids, eps(W.cell_centres(tree, ids, ghosts=True)))
    if lsf_dd is not None:
        orc.set_lsf_distances(*lsf_dd)
    orc.mg_init()
    i, r = W.random_rhs_on_leaves(tree)
    orc.set_cc(M.I_RHS, i, r)
    orc.fas_fmg(True, False)
    for _ in range(n_v):
        orc.fas_vcycle(True)
    return orc, orc.get_cc(M.I_PHI, ids)


def test_eps_one_is_the_plain_laplacian_bit_for_bit():
    tree = T.corner_refined_tree(3, 8, 8, 3)
    o0, p0 = solve(tree)
    o1, p1 = solve(tree, eps=lambda r: np.ones(r.shape[:-1]))
    assert {o1.tag(b) for b in all_ids(tree)} == {0}  # mg_normal_box (m_af_multigrid.f90:1100-1145)
    assert np.array_equal(p0, p1)


def test_constant_eps_box_tags_and_stencil():
    tree = T.uniform_tree(3, 8, 8, 2)
    orc, _ = solve(tree, eps=lambda r: np.full(r.shape[:-1], 2.0), n_v=0)
    for b in all_ids(tree):
        assert orc.tag(b) == 4  # mg_ceps_box
        stype, c, f, _ = orc.op_stencil(b)
        dr = tree.dr[b, 0]
        assert stype == 1 and f is None  # constant after stencil_try_constant
        np.testing.assert_allclose(c[1:], 2.0 / dr ** 2, rtol=1e-15)  # harmonic mean of equal eps
        np.testing.assert_allclose(c[0], -12.0 / dr ** 2, rtol=1e-15)


def test_variable_eps_harmonic_mean_and_prolongation():
    tree = T.uniform_tree(3, 8, 8, 2)
    eps = lambda r: 1.0 + r[..., 0]
    orc, _ = solve(tree, eps=eps, n_v=0)
    b = int(tree.lvl_ids[1][0])
    assert orc.tag(b) == 2  # mg_veps_box
    stype, v, f, _ = orc.op_stencil(b)
    assert stype == 2 and f is None
    e = eps(W.cell_centres(tree, np.array([b]), ghosts=True))[0]
    dr = tree.dr[b, 0]
    a0, am, ap = e[3, 3, 3], e[3, 3, 2], e[3, 3, 4]  # cell (i,j,k) = (3,3,3)
    cell = (3 - 1) + 8 * ((3 - 1) + 8 * (3 - 1))
    np.testing.assert_allclose(v[cell, 1], 2 * a0 * am / (a0 + am) / dr ** 2, rtol=1e-14)  # mg_box_lpld_stencil :1493-1532
    np.testing.assert_allclose(v[cell, 2], 2 * a0 * ap / (a0 + ap) / dr ** 2, rtol=1e-14)
    np.testing.assert_allclose(v[cell, 0], -v[cell, 1:].sum(), rtol=1e-14)
    pst, shape, pv = orc.prolong_stencil(b)
    assert shape == 2 and pst == 2  # variable af_stencil_p234 (mg_box_prolong_eps_stencil :1308-1388)
    np.testing.assert_allclose(pv.sum(axis=1), 1.0, rtol=1e-13)  # weights of an interpolation


def test_lsf_stencil_coefficients():
    tree = T.uniform_tree(3, 8, 8, 1)
    dd = np.ones((1, 8 ** 3, 6))
    cell = (4 - 1) + 8 * ((4 - 1) + 8 * (4 - 1))
    dd[0, cell, 1] = 0.25  # boundary at a quarter of the way to the +x neighbour
    orc, _ = solve(tree, lsf_dd=(np.array([1], np.int32), dd.reshape(1, -1)), lsf_value=2.0, n_v=0)
    assert orc.tag(1) == 1  # mg_lsf_box
    stype, v, f, _ = orc.op_stencil(1)
    dr2 = tree.dr[1, 0] ** 2
    # mg_box_lsf_stencil (:1782-1854): 1 / (0.5 dr^2 (d- + d+) d+-), boundary coupling moved into f
    cm = 1 / (0.5 * dr2 * 1.25 * 1.0)
    cp = 1 / (0.5 * dr2 * 1.25 * 0.25)
    np.testing.assert_allclose(v[cell, 1], cm, rtol=1e-14)
    assert v[cell, 2] == 0.0
    np.testing.assert_allclose(f[cell], -cp, rtol=1e-14)
    np.testing.assert_allclose(v[cell, 0], -(cm + cp + 4 / dr2), rtol=1e-14)
    other = 0
    np.testing.assert_allclose(v[other, 1:], 1 / dr2, rtol=1e-14)
    assert f[other] == 0.0


def test_constant_eps_scales_the_solution():
    tree = T.uniform_tree(3, 8, 8, 3)
    # Dirichlet-0 / Neumann-0 problem: eps * laplace(phi) = rhs  =>  phi(eps=2) = phi(eps=1) / 2
    def run(eps):
        orc = Oracle(tree, with_eps=eps is not None)
        orc.set_bc(W.bc_field_homogeneous(tree, 0.0))
        ids = all_ids(tree)
        if eps is not None:
            orc.set_cc(M.I_EPS, ids, np.full((len(ids),) + (10,) * 3, eps))
        orc.mg_init()
        i, r = W.random_rhs_on_leaves(tree)
        orc.set_cc(M.I_RHS, i, r)
        orc.fas_fmg(True, False)
        for _ in range(8):
            orc.fas_vcycle(True)
        return orc.get_cc(M.I_PHI, ids)
    p1, p2 = run(None), run(2.0)
    assert np.max(np.abs(p1 - 2 * p2)) <= 1e-9 * np.max(np.abs(p1))


def test_lpld_lsf_stencil_reduces_to_its_two_parents():
    """mg_box_lpld_lsf_stencil (m_af_multigrid.f90:1535-1623) pinned by its two limits: with a constant
    permittivity 2 next to an electrode it is 2 x mg_box_lsf_stencil (:1782-1854) (the harmonic mean of equal
    values, exact in binary), and in boxes of the same run without a boundary cell the tag falls back to
    mg_veps_box / mg_ceps_box and mg_box_lpld_stencil."""
    import test_gpu_stencils as G3
    tree = T.uniform_tree(3, 8, 8, 3)
    lsf_dd = G3.lsf_distances(tree, G3.lsf_sphere)
    o_lsf, _ = solve(tree, lsf_dd=lsf_dd, lsf_value=0.5, n_v=0)
    o_both, _ = solve(tree, eps=lambda r: np.full(r.shape[:-1], 2.0), lsf_dd=lsf_dd, lsf_value=0.5, n_v=0)
    seen = set()
    for b in all_ids(tree):
        tag = o_both.tag(b)
        seen.add(tag)
        if tag == 5:  # mg_ceps_box + mg_lsf_box
            assert o_lsf.tag(b) == 1
            s1, v1, f1, _ = o_lsf.op_stencil(b)
            s2, v2, f2, _ = o_both.op_stencil(b)
            assert s1 == s2 == 2
            assert np.array_equal(2.0 * v1, v2) and np.array_equal(2.0 * f1, f2)
        else:
            assert tag == 4 and o_lsf.tag(b) == 0
    assert seen == {4, 5}
    # variable permittivity: away from boundary cells the coefficients are those of mg_box_lpld_stencil
    eps = lambda r: 1.0 + 0.5 * r[..., 0] + 0.25 * r[..., 1] * r[..., 2]
    o_eps, _ = solve(tree, eps=eps, n_v=0)
    o_mix, _ = solve(tree, eps=eps, lsf_dd=lsf_dd, lsf_value=0.5, n_v=0)
    lids, dd = lsf_dd
    dd = dd.reshape(len(lids), -1, 6)
    checked = 0
    for b, d in zip(lids, dd):
        assert o_mix.tag(b) == 3  # mg_veps_box + mg_lsf_box
        _, v_mix, f_mix, _ = o_mix.op_stencil(b)
        _, v_eps, _, _ = o_eps.op_stencil(b)
        free = (d >= 1.0).all(axis=1)
        assert np.array_equal(v_mix[free], v_eps[free]) and np.all(f_mix[free] == 0.0)
        cut = ~free
        assert np.all(f_mix[cut] < 0)  # f = -(sum of the boundary legs), bc_correction = f * V
        assert np.all(v_mix[cut][:, 1:][d[cut] < 1.0] == 0.0)  # boundary legs moved to the right-hand side
        checked += int(cut.sum())
    assert checked > 0
    # and the solve with the mixed stencils converges
    o, _ = solve(tree, eps=eps, lsf_dd=lsf_dd, lsf_value=0.5, n_v=0)
    i, r = W.random_rhs_on_leaves(tree)
    o.set_cc(M.I_RHS, i, r)
    o.fas_fmg(True, False)
    r0 = o.maxabs(M.I_TMP)
    for _ in range(4):
        o.fas_vcycle(True)
    assert o.maxabs(M.I_TMP) < 1e-3 * r0


def test_harmonic_mean_operator_is_exact_for_a_flux_continuous_piecewise_linear_potential():
    """mg_box_lpld_stencil (m_af_multigrid.f90:1493-1532): with the permittivity jumping from 1 to 3 at the cell face
    z = 0.5, the potential a z (z < 0.5), a / 2 + (a / 3)(z - 1/2) (z > 0.5) has a continuous flux eps dphi/dz; the
    harmonic-mean coefficient 2 a0 a / (a0 + a) makes the discrete operator vanish on it exactly, in the two cell
    layers next to the interface as well."""
    tree = T.uniform_tree(3, 8, 8, 3)
    ids = all_ids(tree)
    orc = Oracle(tree, with_eps=True)
    orc.set_bc(W.bc_table(tree, bc_mixed))
    r = W.cell_centres(tree, ids, ghosts=True)
    z = r[..., 2]
    orc.set_cc(M.I_EPS, ids, np.where(z < 0.5, 1.0, 3.0))
    orc.mg_init()
    a = 1.7
    phi = np.where(z < 0.5, a * z, 0.5 * a + (a / 3.0) * (z - 0.5))
    orc.set_cc(M.I_PHI, ids, phi)      # exact values in the ghost cells too
    orc.set_cc(M.I_RHS, ids, np.zeros_like(phi))
    for lvl in range(1, tree.highest_lvl + 1):
        orc.residual_lvl(lvl)
    res = orc.get_cc(M.I_TMP, ids).reshape(phi.shape)[W.interior(tree)]
    scale = a / np.min(tree.dr[ids]) ** 2
    assert np.max(np.abs(res)) < 1e-12 * scale, np.max(np.abs(res)) / scale
    tags = {orc.tag(b) for b in ids}
    assert 2 in tags  # boxes straddling the interface are mg_veps_box


def test_level_set_operator_is_exact_for_a_linear_potential_through_the_electrode_value():
    """mg_box_lsf_stencil (m_af_multigrid.f90:1782-1854) + bc_correction = f * V (:1171-1174): next to a planar
    electrode at x = x0 the neighbour value is replaced by V at distance dd h with the non-uniform second-difference
    weights 1 / (h^2/2 (dd1 + dd2) dd); a potential that is linear and equals V on the plane is annihilated exactly."""
    tree = T.uniform_tree(3, 8, 8, 3)
    ids = all_ids(tree)
    x0, V, s = 0.4321, 0.7, 2.0
    r = W.cell_centres(tree, ids, ghosts=True)
    lsf = r[..., 0] - x0
    nc = tree.nc
    c = lsf[:, 1:-1, 1:-1, 1:-1]
    dd = np.ones((len(ids), nc, nc, nc, 6))
    for m, b in enumerate([lsf[:, 1:-1, 1:-1, :-2], lsf[:, 1:-1, 1:-1, 2:]]):
        cut = c * b < 0
        dd[..., m] = np.where(cut, c / np.where(cut, c - b, 1.0), 1.0)
    has = np.any(dd < 1.0, axis=(1, 2, 3, 4))
    orc = Oracle(tree, lsf_boundary_value=V)
    orc.set_bc(W.bc_table(tree, bc_mixed))
    orc.set_lsf_distances(ids[has], dd[has].reshape(int(has.sum()), -1))
    orc.mg_init()
    phi = V + s * lsf
    orc.set_cc(M.I_PHI, ids, phi)
    orc.set_cc(M.I_RHS, ids, np.zeros_like(phi))
    for lvl in range(1, tree.highest_lvl + 1):
        orc.residual_lvl(lvl)
    res = orc.get_cc(M.I_TMP, ids).reshape(phi.shape)[W.interior(tree)]
    scale = s / np.min(tree.dr[ids]) ** 2
    assert has.any() and np.max(np.abs(res)) < 1e-11 * scale, np.max(np.abs(res)) / scale


def _cyl_dielectric_problem():
    """afivo/examples/poisson_cyl_dielectric.f90:27-28,128-172: manufactured Gaussian (sigma 0.1 at r = 0, z = 0.25) in
    a cylindrical domain whose corner r, z < 0.5 has eps = 100; rhs = eps * lpl_cyl(g) plus the surface charge that
    the jump of eps times the normal gradient puts on every face, split over the two adjacent cells with the
    weights eps_other / (eps_a + eps_b)."""
    r0, s = np.array([0.0, 0.25]), 0.1

    def g(r):
        return np.exp(-np.sum((r - r0) ** 2, axis=-1) / s ** 2)

    def lap_cyl(r):  # gauss_laplacian_cyl, afivo/examples/m_gaussians.f90:107-120
        x = (r - r0) / s
        return 4 / s ** 2 * (np.sum(x ** 2, axis=-1) - 1 - 0.5 * (r[..., 0] - r0[0]) / r[..., 0]) * g(r)

    def grad(r):     # gauss_gradient, :76-89
        return -2 * (r - r0) / s ** 2 * g(r)[..., None]

    def eps_f(r):
        return np.where((r[..., 0] < 0.5) & (r[..., 1] < 0.5), 100.0, 1.0)

    def make_rhs(t, ids):
        c = W.cell_centres(t, ids, ghosts=True)  # (n, z, r, 2)
        e = eps_f(c)
        rhs = lap_cyl(c) * e
        dr = t.dr[ids]
        fx = 0.5 * (c[:, 1:-1, :-1] + c[:, 1:-1, 1:])  # r faces between cells i and i+1, i = 0 .. nc
        q = (e[:, 1:-1, 1:] - e[:, 1:-1, :-1]) * grad(fx)[..., 0] / dr[:, 0][:, None, None]
        w = e[:, 1:-1, 1:] / (e[:, 1:-1, :-1] + e[:, 1:-1, 1:])
        rhs[:, 1:-1, 1:] += w * q
        rhs[:, 1:-1, :-1] += (1 - w) * q
        fy = 0.5 * (c[:, :-1, 1:-1] + c[:, 1:, 1:-1])
        q = (e[:, 1:, 1:-1] - e[:, :-1, 1:-1]) * grad(fy)[..., 1] / dr[:, 1][:, None, None]
        w = e[:, 1:, 1:-1] / (e[:, :-1, 1:-1] + e[:, 1:, 1:-1])
        rhs[:, 1:, 1:-1] += w * q
        rhs[:, :-1, 1:-1] += (1 - w) * q
        return rhs

    def solve_error(t):
        ids = all_ids(t)
        o = Oracle(t, with_eps=True)
        o.set_bc(W.bc_table(t, lambda nb, c: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, g(c))))
        o.set_cc(M.I_EPS, ids, eps_f(W.cell_centres(t, ids, ghosts=True)))
        o.mg_init()
        leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
        o.set_cc(M.I_RHS, leaves, make_rhs(t, leaves))
        res = []
        for it in range(10):
            o.fas_fmg(True, it > 0)
            res.append(o.maxabs(M.I_TMP))
        assert res[-1] < 1e-8 * res[0], res
        c = W.cell_centres(t, leaves, ghosts=True)
        phi = o.get_cc(M.I_PHI, leaves).reshape(c.shape[:-1])
        return o, np.max(np.abs(phi - g(c))[W.interior(t)])

    return lap_cyl, solve_error


def test_poisson_cyl_dielectric_manufactured_solution_second_order():
    """Uniform grids: the error of the converged solution falls by 4 per halving of the spacing across a jump of
    eps by a factor 100 (harmonic-mean operator + cylindrical form)."""
    _, solve_error = _cyl_dielectric_problem()
    errs = [solve_error(T.build_tree(2, 8, [8, 8], lv, None, coord_t=T.AF_CYL))[1] for lv in (4, 5)]
    assert errs[0] < 2e-2 and 3.5 < errs[0] / errs[1] < 4.5, errs


def test_poisson_cyl_dielectric_on_the_adaptive_tree():
    """The example's adaptively refined tree (7 levels): variable-eps, constant-eps and plain boxes side by side,
    refinement boundaries with mg_sides_rb_extrap and mg_box_prolong_eps_stencil on the jump."""
    lap_cyl, solve_error = _cyl_dielectric_problem()
    nc = 8

    def refine(l, ixs, ctr):
        dr = 1.0 / (nc * 2 ** (l - 1))
        off = (np.arange(nc) - (nc - 1) / 2) * dr
        gy, gx = np.meshgrid(off, off, indexing="ij")
        pts = ctr[:, None, :] + np.stack([gx, gy], axis=-1).reshape(1, -1, 2)
        return (dr * dr * np.max(np.abs(lap_cyl(pts)), axis=1) > 1e-3) & (l < 7)

    t = T.build_tree(2, nc, [nc, nc], 7, refine, coord_t=T.AF_CYL)
    assert t.highest_lvl == 7
    o, err = solve_error(t)
    assert {o.tag(int(b)) for b in all_ids(t)} == {0, 2, 4}
    assert err < 3e-4, err
