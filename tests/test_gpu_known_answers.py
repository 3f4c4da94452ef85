"""-m gpu: analytic known answers evaluated on the CUDA path itself (no oracle in the loop): what the reference's
discretisation must return by construction -- the discrete sine eigenvector, the constant Helmholtz / Neumann
solution, exactness of the cylindrical form for r^2 + z^2, and a constant field from a linear potential with exact
ghost cells of its norm (afivo/examples/check_ghostcells.f90 style)."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

pytestmark = pytest.mark.gpu


def all_ids(tree):
    return np.concatenate(tree.lvl_ids).astype(np.int32)


def test_discrete_sine_eigenvector():
    """sin(pi x) sin(2 pi y) sin(3 pi z) at the cell centres is an eigenvector of the 7-point operator with Dirichlet-0
    ghost cells (ghost = -phi_1), eigenvalue -(4 / h^2) sum sin^2(m pi h / 2): zero residual on every level, fixed point
    of the half-sweeps, and what the coarse-grid solve returns."""
    t = T.uniform_tree(3, 8, 8, 3)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(t))
    M.mg_init(t, mg)
    ids = all_ids(t)
    r = W.cell_centres(t, ids, ghosts=True)
    modes = np.array([1.0, 2.0, 3.0])
    phi = np.prod(np.sin(np.pi * modes * r), axis=-1)
    h = t.dr[ids, 0]
    lam = -(4 / h ** 2) * np.sum(np.sin(np.pi * modes[None, :] * h[:, None] / 2) ** 2, axis=1)
    rhs = lam[:, None, None, None] * phi
    mg.set_cc(M.I_PHI, ids, phi)
    mg.set_cc(M.I_RHS, ids, rhs)
    for lvl in range(1, t.highest_lvl + 1):
        mg.gc_lvl(lvl, M.I_PHI, True)
        mg.residual_lvl(lvl)
    inner = W.interior(t)
    tmp = mg.get_cc(M.I_TMP, ids)[inner]
    scale = np.max(np.abs(rhs))
    assert np.max(np.abs(tmp)) < 1e-12 * scale, np.max(np.abs(tmp)) / scale
    leaves = t.leaves(t.highest_lvl).astype(np.int32)
    before = mg.get_cc(M.I_PHI, leaves)[inner]
    mg.gsrb_boxes(t.highest_lvl, M.MG_CYCLE_DOWN)
    after = mg.get_cc(M.I_PHI, leaves)[inner]
    assert np.max(np.abs(after - before)) < 1e-13
    mg.set_cc(M.I_PHI, ids[:1], np.zeros_like(phi[:1]))
    mg.solve_coarse_grid()
    got = mg.get_cc(M.I_PHI, ids[:1])[inner]
    assert np.max(np.abs(got - phi[:1][inner])) < 1e-11
    M.mg_destroy(mg)


def test_helmholtz_neumann_constant_solution():
    """lpl(phi) - lambda phi = f with zero-flux boundaries everywhere and constant f: phi = -f / lambda after one FMG."""
    t = T.corner_refined_tree(3, 8, 8, 4)
    lam, f = 50.0, 3.0
    mg = M.mg_t(sides_bc=W.bc_neumann_zero(t), helmholtz_lambda=lam)
    M.mg_init(t, mg)
    ids, rhs = W.constant_rhs_on_leaves(t, f)
    mg.set_cc(M.I_RHS, ids, rhs)
    M.mg_fas_fmg(t, mg, True, False)
    phi = mg.get_cc(M.I_PHI, all_ids(t))
    assert np.max(np.abs(phi + f / lam)) < 1e-12, np.max(np.abs(phi + f / lam))
    assert M.af_tree_maxabs_cc(t, mg, M.I_TMP) < 1e-10
    M.mg_destroy(mg)


def test_cylindrical_operator_exact_for_r2_plus_z2():
    """[(r + h/2)(2 r h + h^2) - (r - h/2)(2 r h - h^2)] / (r h^2) = 4, d2/dz2 z^2 = 2: L(r^2 + z^2) = 6 in every cell,
    the axis included (its inner flux factor vanishes)."""
    t = T.build_tree(2, 8, [8, 8], 4, lambda l, ix, c: (c[:, 0] < 0.6) & (np.abs(c[:, 1] - 0.5) < 0.3), coord_t=T.AF_CYL)
    bc = W.bc_table(t, lambda nb, c: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, (c ** 2).sum(axis=-1)))
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(t, mg)
    ids = all_ids(t)
    phi = (W.cell_centres(t, ids, ghosts=True) ** 2).sum(axis=-1)  # exact values in the ghost cells too
    mg.set_cc(M.I_PHI, ids, phi)
    mg.set_cc(M.I_RHS, ids, np.full_like(phi, 6.0))
    for lvl in range(1, t.highest_lvl + 1):
        mg.residual_lvl(lvl)
    res = mg.get_cc(M.I_TMP, ids)[W.interior(t)]
    assert np.max(np.abs(res)) < 1e-9, np.max(np.abs(res))
    M.mg_destroy(mg)


@pytest.mark.parametrize("mk", [lambda: T.corner_refined_tree(3, 8, 8, 4), lambda: T.corner_refined_tree(2, 8, 8, 5)])
def test_linear_potential_gives_constant_field_and_exact_norm_ghost_cells(mk):
    t = mk()
    nd = t.ndim
    g = np.array([1.5, -0.75, 2.25][:nd])
    ids = all_ids(t)
    phi = W.cell_centres(t, ids, ghosts=True) @ g + 0.3
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(t))
    M.mg_init(t, mg)
    mg.set_cc(M.I_PHI, ids, phi)
    M.field_from_potential(t, mg, -1.0)
    n1, nc = t.nc + 1, t.nc
    fc = mg.get_fc(ids).reshape((len(ids), nd) + (n1,) * nd)
    for d in range(nd):
        sl = [slice(None), d] + [slice(0, nc)] * nd
        sl[2 + (nd - 1 - d)] = slice(0, n1)
        assert np.allclose(fc[tuple(sl)], -g[d], rtol=0, atol=1e-10)
    fld = mg.get_cc(M.I_FLD, ids)
    assert np.allclose(fld, np.linalg.norm(g), rtol=0, atol=1e-10), np.abs(fld - np.linalg.norm(g)).max()
    M.mg_destroy(mg)


@pytest.mark.parametrize("ndim", [2, 3])
def test_poisson_neumann_linear_solution(ndim):
    """afivo/examples/poisson_neumann.f90 (Cartesian): rhs = 0, Dirichlet 0 at low x, Neumann 1 at high x, Neumann 0
    elsewhere, refined while lvl <= 4 and all(r_min < 0.25): phi = x, reproduced to rounding."""
    nc = 8
    t = T.build_tree(ndim, nc, [nc] * ndim, 5,
                     lambda l, ixs, ctr: (l <= 4) & np.all((ixs - 1) * (1.0 / 2 ** (l - 1)) < 0.25, axis=1))
    assert t.highest_lvl == 5
    bc = W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0) if nb == 1 else (W.AF_BC_NEUMANN, 1.0 if nb == 2 else 0.0))
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(t, mg)
    leaves = np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)
    x = W.cell_centres(t, leaves, ghosts=True)[..., 0]
    for it in range(10):
        M.mg_fas_fmg(t, mg, True, it > 0)
    err = np.max(np.abs(mg.get_cc(M.I_PHI, leaves) - x)[W.interior(t)])
    assert err < 1e-11, err
    assert M.af_tree_maxabs_cc(t, mg, M.I_TMP) < 1e-9
    M.mg_destroy(mg)
