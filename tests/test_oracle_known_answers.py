"""CPU: pins the oracle against the known answers the reference's own tests / examples hold for this
path.  The reference ships no numeric golden vectors for the multigrid."""
import numpy as np
import pytest

from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import I_PHI, I_RHS, I_TMP, MG_CYCLE_DOWN, Oracle

from util import all_ids


def test_box_count_five_full_levels():
    """afivo/tests/answers/test_refinement_3d: 4681 == (1 - 8^5) / (1 - 8)."""
    t = T.uniform_tree(3, 8, 8, 5)
    assert t.n_boxes == 4681 == (1 - 8 ** 5) // (1 - 8)
    assert [len(a) for a in t.lvl_ids] == [1, 8, 64, 512, 4096]


@pytest.mark.parametrize("ndim", [2, 3])
def test_zero_field_neumann_keeps_zero_ghosts(ndim):
    """afivo/tests/test_ghostcell.f90:14-30,54-76: all-zero field, Neumann-0, corner-refined 4-level tree."""
    t = T.corner_refined_tree(ndim, 8, 8, 4)
    o = Oracle(t)
    o.set_bc(W.bc_neumann_zero(t))
    with pytest.raises(RuntimeError, match="code 4"):
        o.mg_init()  # all-Neumann Poisson: singular coarse operator; ghost cells do not need it
    for lvl in range(1, t.highest_lvl + 1):
        o.gc_lvl(lvl, I_PHI, True)
    assert np.all(o.get_cc(I_PHI, all_ids(t)) == 0.0)


def _linear(c):
    return 0.5 + c[..., 0] * 1.25 - 0.75 * c[..., 1] + (0.3 * c[..., 2] if c.shape[-1] == 3 else 0.0)


@pytest.mark.parametrize("ndim", [2, 3])
def test_ghost_cells_exact_for_linear_field(ndim):
    """afivo/examples/check_ghostcells.f90 (test_gradient): ghost cells of a linear field are exact,
    for same-level copies, refinement boundaries (mg_sides_rb), Dirichlet faces, edges and corners."""
    t = T.corner_refined_tree(ndim, 8, 8, 4)
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_function(t, _linear))
    o.mg_init()
    ids = all_ids(t)
    exact = _linear(W.cell_centres(t, ids, ghosts=True))
    data = np.zeros_like(exact)
    data[W.interior(t)] = exact[W.interior(t)]
    o.set_cc(I_PHI, ids, data)
    for lvl in range(1, t.highest_lvl + 1):
        o.gc_lvl(lvl, I_PHI, True)
    got = o.get_cc(I_PHI, ids).reshape(exact.shape)
    assert np.max(np.abs(got - exact)) < 1e-13


@pytest.mark.parametrize("ndim", [2, 3])
def test_prolongation_exact_for_linear_field(ndim):
    """afivo/examples/check_prolongation.f90: (bi/tri)linear prolongation reproduces a linear field."""
    t = T.uniform_tree(ndim, 8, 8, 3)
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_function(t, _linear))
    o.mg_init()
    ids = all_ids(t)
    exact = _linear(W.cell_centres(t, ids, ghosts=True))
    phi = exact.copy()
    phi[t.lvl[ids] > 1] = 0.0          # children start from zero, tmp = 0: correction = parent phi
    o.set_cc(I_PHI, ids, phi)
    o.correct_children(1)
    got = o.get_cc(I_PHI, ids).reshape(exact.shape)
    lvl2 = t.lvl[ids] == 2
    assert np.max(np.abs(got[lvl2][W.interior(t)] - exact[lvl2][W.interior(t)])) < 1e-13


def test_restriction_is_the_2d_average():
    """af_restrict_box (m_af_restrict.f90:120-133)."""
    t = T.uniform_tree(3, 8, 8, 2)
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_zero(t))
    o.mg_init()
    ids, rhs = W.random_rhs_on_leaves(t)
    o.set_cc(I_RHS, ids, rhs)
    o.init_phi_rhs()
    parent = o.get_cc(I_RHS, np.array([1], np.int32)).reshape(10, 10, 10)[1:9, 1:9, 1:9]
    fine = np.zeros((16, 16, 16))
    for q, cid in enumerate(t.children[1]):
        ox, oy, oz = [(t.ix[cid, d] - 1) * 8 for d in range(3)]
        fine[oz:oz + 8, oy:oy + 8, ox:ox + 8] = rhs[list(ids).index(cid)][1:9, 1:9, 1:9]
    avg = fine.reshape(8, 2, 8, 2, 8, 2).mean(axis=(1, 3, 5))
    assert np.max(np.abs(avg - parent)) < 1e-15


def test_poisson_basic_gaussians_converge():
    """afivo/examples/poisson_basic.f90:40-63,104-119,143-165: two Gaussians, Dirichlet = analytic
    solution, refine where dr^2 |rhs| > 1e-3; FMG cycles: residual falls, error plateaus at the
    discretisation level."""
    g = W.Gaussians([[0.1, 0.1, 0.1], [0.75, 0.75, 0.75]], 0.04)
    nc = 8

    def refine(l, ixs, ctr):
        dr = 1.0 / (nc * 2 ** (l - 1))
        # afivo/examples/poisson_basic.f90:143-165: refine if dr^2 * max |rhs| over the box cells > 1e-3
        off = (np.arange(nc) - (nc - 1) / 2) * dr
        gz, gy, gx = np.meshgrid(off, off, off, indexing="ij")
        pts = ctr[:, None, :] + np.stack([gx, gy, gz], axis=-1).reshape(1, -1, 3)
        m = np.max(np.abs(g.laplacian(pts)), axis=1)
        return dr * dr * m > 1e-3

    t = T.build_tree(3, nc, [nc] * 3, 5, refine)
    assert t.highest_lvl >= 4
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_function(t, g.value))
    o.mg_init()
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    ctr = W.cell_centres(t, leaves, ghosts=True)
    rhs = g.laplacian(ctr)
    o.set_cc(I_RHS, leaves, rhs)
    res, err = [], []
    for it in range(6):
        o.fas_fmg(True, it > 0)
        res.append(o.maxabs(I_TMP))
        phi = o.get_cc(I_PHI, leaves).reshape(ctr.shape[:-1])
        err.append(np.max(np.abs(phi - g.value(ctr))[W.interior(t)]))
    assert all(res[i + 1] < 0.5 * res[i] for i in range(4)), res
    assert res[-1] < 1e-5 * res[0]
    assert err[-1] < 2e-2 and abs(err[-1] - err[-2]) < 1e-3 * err[-1], err  # plateau


def test_helmholtz_lambda_enters_the_centre_coefficient():
    """mg_box_lpl_stencil (m_af_multigrid.f90:1262): c(1) = -sum(c(2:)) - lambda."""
    t = T.uniform_tree(3, 8, 8, 2)
    o = Oracle(t, helmholtz_lambda=1.0e3)
    o.set_bc(W.bc_helmholtz(t))
    o.mg_init()
    stype, c, f, cyl = o.op_stencil(2)
    idr2 = 1.0 / t.dr[2] ** 2
    assert stype == 1 and f is None and not cyl
    assert np.allclose(c[1:], np.repeat(idr2, 2)) and c[0] == -np.sum(c[1:]) - 1.0e3


def test_results_do_not_depend_on_thread_count():
    """the cycle is a deterministic dataflow, independent of box order / threads."""
    t = T.corner_refined_tree(3, 8, 8, 3)
    out = []
    for nthreads in (1, 4):
        o = Oracle(t)
        o.set_num_threads(nthreads)
        o.set_bc(W.bc_field_homogeneous(t, 1.0))
        o.mg_init()
        ids, rhs = W.random_rhs_on_leaves(t)
        o.set_cc(I_RHS, ids, rhs)
        o.fas_fmg(True, False)
        o.fas_vcycle(True)
        out.append(o.get_cc(I_PHI, all_ids(t)))
    o.set_num_threads(o.num_threads())
    assert np.array_equal(out[0], out[1])


def test_gsrb_colour_convention():
    """stencil_gsrb_357 (m_af_stencil.f90:962): half-sweep n updates cells with (i+j+k+n) even."""
    t = T.uniform_tree(3, 8, 8, 1)
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_zero(t))
    o.mg_init()
    ids = all_ids(t)
    rng = np.random.default_rng(1)
    phi = rng.normal(size=(1, 10, 10, 10))
    o.set_cc(I_PHI, ids, phi)
    o.set_cc(I_RHS, ids, rng.normal(size=(1, 10, 10, 10)))
    o.box_gsrb_lvl(1, 1)
    new = o.get_cc(I_PHI, ids).reshape(10, 10, 10)
    k, j, i = np.indices((10, 10, 10))
    changed = new != phi[0]
    inner = (i >= 1) & (i <= 8) & (j >= 1) & (j <= 8) & (k >= 1) & (k <= 8)
    assert np.all(changed[inner & ((i + j + k) % 2 == 1)])
    assert not np.any(changed[~(inner & ((i + j + k) % 2 == 1))])


def test_discrete_eigenfunction_of_the_cell_centred_laplacian():
    """Analytic pin of stencil + Dirichlet ghost cells + coarse solve: on a cell-centred grid with ghost = 2 b - phi_1
    (bc_to_gc, m_af_ghostcell.f90:192-214, b = 0) the product of sines sin(pi x) sin(2 pi y) sin(3 pi z) sampled at
    the cell centres is an exact eigenvector of the 7-point operator (mg_box_lpl_stencil, m_af_multigrid.f90:1246-1264)
    with eigenvalue -(4 / h^2) (sin^2(pi h / 2) + sin^2(2 pi h / 2) + sin^2(3 pi h / 2)).  So with rhs = lambda_h * phi
    the residual of phi vanishes on every level, a V-cycle leaves phi unchanged, and the coarse-grid solve returns it."""
    t = T.uniform_tree(3, 8, 8, 3)
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_zero(t))
    o.mg_init()
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    r = W.cell_centres(t, ids, ghosts=True)
    modes = np.array([1.0, 2.0, 3.0])
    phi = np.prod(np.sin(np.pi * modes * r), axis=-1)
    lam = np.zeros(len(ids))
    for q, b in enumerate(ids):
        h = t.dr[b, 0]
        lam[q] = -(4 / h ** 2) * np.sum(np.sin(np.pi * modes * h / 2) ** 2)
    rhs = lam[:, None, None, None] * phi
    o.set_cc(I_PHI, ids, phi)
    o.set_cc(I_RHS, ids, rhs)
    for lvl in range(1, t.highest_lvl + 1):
        o.gc_lvl(lvl, I_PHI, True)      # Dirichlet ghost cells reproduce the odd extension of the sines
        o.residual_lvl(lvl)
    tmp = o.get_cc(I_TMP, ids).reshape(phi.shape)[W.interior(t)]
    scale = np.max(np.abs(rhs))
    assert np.max(np.abs(tmp)) < 1e-12 * scale, np.max(np.abs(tmp)) / scale
    # coarse grid: the direct solve of the BC-folded level-1 matrix returns the eigenvector
    one = ids[:1]
    o.set_cc(I_PHI, one, np.zeros_like(phi[:1]))
    o.solve_coarse_grid()
    got = o.get_cc(I_PHI, one).reshape(phi[:1].shape)[W.interior(t)]
    assert np.max(np.abs(got - phi[:1][W.interior(t)])) < 1e-12
    # on the finest level the eigenvector with its own rhs is a fixed point of the smoother
    leaves = t.leaves(t.highest_lvl).astype(np.int32)
    before = o.get_cc(I_PHI, leaves).copy()
    o.gsrb_boxes(t.highest_lvl, 1)
    after = o.get_cc(I_PHI, leaves)
    inner = W.interior(t)
    shape = (len(leaves),) + (t.nc + 2,) * 3
    assert np.max(np.abs(after.reshape(shape)[inner] - before.reshape(shape)[inner])) < 1e-13


def test_cylindrical_operator_is_exact_for_r2_plus_z2():
    """The conservative cylindrical 5-point form (cc_cyl, m_af_stencil.f90:886-925, af_cyl_flux_factors
    m_af_types.f90:1199-1211) differentiates r^2 and z^2 exactly: [(r + h/2)(2 r h + h^2) - (r - h/2)(2 r h - h^2)]
    / (r h^2) = 4 and the z part gives 2, so L(r^2 + z^2) = 6 in every cell whose stencil does not touch a domain
    boundary ghost cell -- on the axis too, where the inner flux factor vanishes."""
    t = T.build_tree(2, 8, [8, 8], 4, lambda l, ix, c: (c[:, 0] < 0.6) & (np.abs(c[:, 1] - 0.5) < 0.3), coord_t=T.AF_CYL)
    o = Oracle(t)
    o.set_bc(W.bc_table(t, lambda nb, c: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, (c ** 2).sum(axis=-1))))
    o.mg_init()
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    r = W.cell_centres(t, ids, ghosts=True)
    phi = (r ** 2).sum(axis=-1)          # exact values in ALL cells incl. ghost cells: no ghost fill needed
    rhs = np.full_like(phi, 6.0)
    o.set_cc(I_PHI, ids, phi)
    o.set_cc(I_RHS, ids, rhs)
    for lvl in range(1, t.highest_lvl + 1):
        o.residual_lvl(lvl)
    res = o.get_cc(I_TMP, ids).reshape(phi.shape)[W.interior(t)]
    assert np.max(np.abs(res)) < 1e-9, np.max(np.abs(res))  # 1/h^2 = 4096^... rounding of O(1e3) terms
    # and with ghost cells filled by the library's own rules the cells next to the axis stay exact (Neumann-0 there
    # multiplies a vanishing flux factor), those next to refinement boundaries only to second order
    lv1 = t.lvl_ids[0].astype(np.int32)
    o.gc_lvl(1, I_PHI, True)
    o.residual_lvl(1)
    res1 = o.get_cc(I_TMP, lv1).reshape((len(lv1), t.nc + 2, t.nc + 2))
    assert np.max(np.abs(res1[:, 2:-2, 1:-2])) < 1e-9   # rows away from the z boundaries, columns from the axis on


def test_helmholtz_with_neumann_everywhere_has_the_constant_solution():
    """Sign convention and Neumann folding in one known answer: lpl(phi) - lambda * phi = f (m_af_types.f90:595-597,
    mg_box_lpl_stencil c(1) = -sum(c(2:)) - lambda) with zero-flux boundaries on all faces (stencil_handle_boundaries:
    diag += c_nb, m_coarse_solver.f90:466-476) and constant f has the constant solution phi = -f / lambda; one FMG
    cycle on a refined tree must land on it to rounding."""
    t = T.corner_refined_tree(3, 8, 8, 4)
    lam, f = 50.0, 3.0
    o = Oracle(t, helmholtz_lambda=lam)
    o.set_bc(W.bc_neumann_zero(t))
    o.mg_init()
    ids, rhs = W.constant_rhs_on_leaves(t, f)
    o.set_cc(I_RHS, ids, rhs)
    o.fas_fmg(True, False)
    all_ids = np.concatenate(t.lvl_ids).astype(np.int32)
    phi = o.get_cc(I_PHI, all_ids)
    assert np.max(np.abs(phi + f / lam)) < 1e-13, np.max(np.abs(phi + f / lam))
    assert o.maxabs(I_TMP) < 1e-11


def _neumann_tree(ndim, coord_t=T.AF_XYZ):
    """afivo/examples/poisson_neumann.f90:88-97: refine while lvl <= 4 and all(r_min < 0.25)."""
    nc = 8
    dr1 = 1.0 / nc

    def refine(l, ixs, ctr):
        rmin = (ixs - 1) * (nc * dr1 / 2 ** (l - 1))
        return (l <= 4) & np.all(rmin < 0.25, axis=1)

    return T.build_tree(ndim, nc, [nc] * ndim, 5, refine, coord_t=coord_t)


@pytest.mark.parametrize("ndim", [2, 3])
def test_poisson_neumann_linear_solution_is_reproduced_exactly(ndim):
    """afivo/examples/poisson_neumann.f90 (Cartesian): rhs = 0, Dirichlet 0 at low x, Neumann 1 at high x, Neumann 0
    elsewhere -> phi = x.  A linear potential is exact for the 2nd-order operator, the Neumann / Dirichlet ghost
    cells and the refinement-boundary interpolation, so the multigrid converges to it up to rounding."""
    t = _neumann_tree(ndim)
    assert t.highest_lvl == 5

    def sides(nb, c):
        if nb == 1:
            return W.AF_BC_DIRICHLET, 0.0
        return W.AF_BC_NEUMANN, 1.0 if nb == 2 else 0.0

    o = Oracle(t)
    o.set_bc(W.bc_table(t, sides))
    o.mg_init()
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    x = W.cell_centres(t, leaves, ghosts=True)[..., 0]
    err = []
    for it in range(10):
        o.fas_fmg(True, it > 0)
        phi = o.get_cc(I_PHI, leaves).reshape(x.shape)
        err.append(np.max(np.abs(phi - x)[W.interior(t)]))
    assert err[-1] < 1e-12, err
    assert o.maxabs(I_TMP) < 1e-9


def test_poisson_neumann_cylindrical_r_squared():
    """afivo/examples/poisson_neumann.f90 with cylindrical = T: rhs = 4, zero flux on the axis, d(phi)/dr = 2 at
    r = 1, Dirichlet r^2 in z -> phi = r^2.  Operator and Neumann ghost cells (central difference of a quadratic) are
    exact for r^2, so a uniform grid reproduces it to rounding; on the example's corner-refined tree only the linear
    refinement-boundary interpolation is inexact and the error stays at the 1e-4 level."""
    def sides(nb, c):
        if nb == 1:
            return W.AF_BC_NEUMANN, 0.0
        if nb == 2:
            return W.AF_BC_NEUMANN, 2.0
        return W.AF_BC_DIRICHLET, c[..., 0] ** 2

    errs = []
    for t in (T.build_tree(2, 8, [8, 8], 4, None, coord_t=T.AF_CYL), _neumann_tree(2, T.AF_CYL)):
        o = Oracle(t)
        o.set_bc(W.bc_table(t, sides))
        o.mg_init()
        ids, rhs = W.constant_rhs_on_leaves(t, 4.0)
        o.set_cc(I_RHS, ids, rhs)
        for it in range(10):
            o.fas_fmg(True, it > 0)
        r = W.cell_centres(t, ids, ghosts=True)[..., 0]
        phi = o.get_cc(I_PHI, ids).reshape(r.shape)
        errs.append(np.max(np.abs(phi - r ** 2)[W.interior(t)]))
        assert o.maxabs(I_TMP) < 1e-8
    assert errs[0] < 1e-11 and 0 < errs[1] < 1e-3, errs


@pytest.mark.parametrize("ndim", [2, 3])
def test_poisson_helmholtz_gaussians_second_order(ndim):
    """afivo/examples/poisson_helmholtz.f90:18,32-34,119-128,145-161: lpl(phi) - lambda phi = lpl(g) - lambda g with
    lambda = 1e3, two Gaussians (sigma 0.04) at 0.25 and 0.75, Dirichlet = analytic.  On uniform grids the error of
    the converged solution falls by ~4 per halving of the spacing (2nd order), and it is smaller than for lambda = 0
    (the Helmholtz term damps it)."""
    g = W.Gaussians([[0.25] * ndim, [0.75] * ndim], 0.04)
    lam = 1.0e3
    errs = {}
    lo, hi = 4, 5  # 64 and 128 cells per side: sigma / h = 2.6 and 5.1 (coarser grids are pre-asymptotic)
    for lam_, lvls in ((lam, lo), (lam, hi), (0.0, lo)):
        t = T.uniform_tree(ndim, 8, 8, lvls)
        o = Oracle(t, helmholtz_lambda=lam_)
        o.set_bc(W.bc_dirichlet_function(t, g.value))
        o.mg_init()
        leaves = t.leaves(t.highest_lvl).astype(np.int32)
        ctr = W.cell_centres(t, leaves, ghosts=True)
        o.set_cc(I_RHS, leaves, g.laplacian(ctr) - lam_ * g.value(ctr))
        for it in range(5):
            o.fas_fmg(True, it > 0)
        assert o.maxabs(I_TMP) < 1e-6 * np.max(np.abs(g.laplacian(ctr)))
        phi = o.get_cc(I_PHI, leaves).reshape(ctr.shape[:-1])
        errs[(lam_, lvls)] = np.max(np.abs(phi - g.value(ctr))[W.interior(t)])
    ratio = errs[(lam, lo)] / errs[(lam, hi)]
    assert 3.0 < ratio < 5.0, (ratio, errs)
    assert errs[(lam, lo)] < errs[(0.0, lo)]


def test_poisson_cyl_analytic_gaussian_charge_on_the_axis():
    """afivo/examples/poisson_cyl_analytic.f90:17-25,104-169,189-208: a Gaussian charge on the axis of a cylindrical
    domain has the potential Q erf(d / (sqrt(2) sigma)) / (4 pi eps0 d); Neumann-0 on the axis, Dirichlet = analytic
    elsewhere, refined where dr^2 |rhs| > 0.1.  The 2D cylindrical operator therefore has to reproduce a genuinely
    three-dimensional (1/d) potential: the relative error falls from 3e-3 (6 levels) to 1e-4 (8 levels)."""
    from scipy.special import erf
    L = 1.25e-2
    sigma = 4e-4 * np.sqrt(0.5)
    src = np.array([0.0, 0.5]) * L
    eps0 = 8.85e-12
    Q = 3e18 * 1.6022e-19 * sigma ** 3 * np.sqrt(2 * np.pi) ** 3
    nc = 8

    def rhs_f(r):
        return -Q * np.exp(-np.sum((r - src) ** 2, axis=-1) / (2 * sigma ** 2)) / (sigma ** 3 * np.sqrt(2 * np.pi) ** 3 * eps0)

    def sol(r):
        d = np.linalg.norm(r - src, axis=-1)
        small = d < np.sqrt(np.finfo(float).eps)
        safe = np.where(small, 1.0, d)
        return np.where(small, np.sqrt(2 / np.pi) / sigma, erf(safe * np.sqrt(0.5) / sigma) / safe) * Q / (4 * np.pi * eps0)

    rel = []
    for max_lvl in (6, 8):
        def refine(l, ixs, ctr):
            dr = L / (nc * 2 ** (l - 1))
            off = (np.arange(nc) - (nc - 1) / 2) * dr
            gy, gx = np.meshgrid(off, off, indexing="ij")
            pts = ctr[:, None, :] + np.stack([gx, gy], axis=-1).reshape(1, -1, 2)
            return (dr * dr * np.max(np.abs(rhs_f(pts)), axis=1) > 1e-1) & (l < max_lvl)

        t = T.build_tree(2, nc, [nc, nc], max_lvl, refine, r_max=[L, L], coord_t=T.AF_CYL)
        assert t.highest_lvl == max_lvl
        o = Oracle(t)
        o.set_bc(W.bc_table(t, lambda nb, c: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, sol(c))))
        o.mg_init()
        leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
        ctr = W.cell_centres(t, leaves, ghosts=True)
        o.set_cc(I_RHS, leaves, rhs_f(ctr))
        res = []
        for it in range(10):
            o.fas_fmg(True, it > 0)
            res.append(o.maxabs(I_TMP))
        assert res[-1] < 1e-7 * res[0], res
        phi = o.get_cc(I_PHI, leaves).reshape(ctr.shape[:-1])
        rel.append(np.max(np.abs(phi - sol(ctr))[W.interior(t)]) / np.max(sol(ctr)))
    assert rel[0] < 5e-3 and rel[1] < 3e-4 and rel[1] < 0.1 * rel[0], rel


@pytest.mark.parametrize("ndim", [2, 3])
def test_implicit_diffusion_in_a_periodic_domain_follows_the_discrete_decay(ndim):
    """afivo/examples/helmholtz_variable_stencil.f90:9-21,40-47,66-71,113-128: backward-Euler diffusion steps
    (lpl - 1/(D dt)) phi_new = -phi_old / (D dt) in a fully periodic domain of length 2 pi, started from
    1 + cos(x) cos(y).  cos(x) cos(y) is an eigenvector of the periodic 5/7-point operator with eigenvalue
    -mu = -(4 / h^2) * 2 sin^2(h / 2), so every converged step multiplies its amplitude by exactly 1 / (1 + D dt mu)
    and leaves the constant 1 alone; mg_update_operator_stencil's role (lambda changes with dt) is played by a new
    oracle per dt."""
    dlen = 2 * np.arccos(-1.0)
    D = 1.0
    t = T.build_tree(ndim, 8, [8] * ndim, 3, None, r_max=[dlen] * ndim, periodic=[True] * ndim)
    h = t.dr[t.leaves(3)[0], 0]
    mu = (4 / h ** 2) * 2 * np.sin(h / 2) ** 2
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    leaves = t.leaves(3).astype(np.int32)
    rr = W.cell_centres(t, ids, ghosts=True)
    mode = np.cos(rr[..., 0]) * np.cos(rr[..., 1])
    for k_factor in (1, 4):
        dt = 0.1 / k_factor
        o = Oracle(t, helmholtz_lambda=1 / (D * dt))
        o.set_bc(W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0)))  # no physical faces: an empty table
        o.mg_init()
        o.set_cc(I_PHI, ids, 1 + mode)
        amp = 1.0
        sel = np.isin(ids, leaves)
        for step in range(3):
            phi_old = o.get_cc(I_PHI, leaves)
            o.set_cc(I_RHS, leaves, -phi_old / (dt * D))
            for it in range(4):
                o.fas_fmg(True, True)
            amp /= 1 + D * dt * mu
            phi = o.get_cc(I_PHI, leaves).reshape(mode[sel].shape)
            err = np.max(np.abs(phi - (1 + amp * mode[sel]))[W.interior(t)])
            assert err < 1e-9, (k_factor, step, err)
        # the continuum decay exp(-2 D t) is matched to first order in dt and second order in h
        assert abs(amp - np.exp(-2 * D * 3 * dt)) < 0.6 * dt


@pytest.mark.parametrize("ndim,seed", [(2, 1), (2, 2), (2, 3), (3, 4), (3, 5)])
def test_ghost_cells_and_prolongation_exact_for_linear_field_on_random_trees(ndim, seed):
    """afivo/examples/check_ghostcells.f90 / check_prolongation.f90 refine and derefine at random and assert that ghost
    cells and prolongation reproduce a linear field exactly; here: random 2:1-balanced trees (level 1 refined, each finer box
    with probability 0.4), all levels, sides + edges + corners, then a full correct_children pass."""
    rng = np.random.default_rng(seed)
    t = T.build_tree(ndim, 8, [16] * ndim if ndim == 2 else [8] * ndim, 4 if ndim == 2 else 3,
                     lambda l, ixs, ctr: (rng.random(len(ixs)) < 0.4) | (l == 1))
    assert t.highest_lvl >= 2
    o = Oracle(t)
    o.set_bc(W.bc_dirichlet_function(t, _linear))
    o.mg_init()
    ids = all_ids(t)
    exact = _linear(W.cell_centres(t, ids, ghosts=True))
    inner = W.interior(t)
    data = np.zeros_like(exact)
    data[inner] = exact[inner]
    o.set_cc(I_PHI, ids, data)
    for lvl in range(1, t.highest_lvl + 1):
        o.gc_lvl(lvl, I_PHI, True)
    got = o.get_cc(I_PHI, ids).reshape(exact.shape)
    assert np.max(np.abs(got - exact)) < 1e-13
    # prolongation: children of level-1 parents start from zero, tmp = 0, so the correction is the parent's phi
    phi = exact.copy()
    lvl2 = t.lvl[ids] == 2
    phi[lvl2] = 0.0
    o.set_cc(I_PHI, ids, phi)
    o.set_cc(I_TMP, ids, np.zeros_like(phi))
    o.correct_children(1)
    got = o.get_cc(I_PHI, ids).reshape(exact.shape)
    assert lvl2.any() and np.max(np.abs(got[lvl2][inner] - exact[lvl2][inner])) < 1e-13
