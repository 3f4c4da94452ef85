"""-m gpu: whole FMG / V-cycles against the oracle.  Tolerance from BASELINE.json north_star: fp64
potential <= 1e-10 relative max-norm after the same number of cycles, same per-cycle residual
history."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from util import TREES, all_ids, bc_mixed

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def run_both(tree, bc_fn, n_v=4, have_guess=False, **opts):
    bc = W.bc_table(tree, bc_fn)
    orc = Oracle(tree, **opts)
    orc.set_bc(bc)
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc, **opts)
    M.mg_init(tree, mg)
    ids, rhs = W.random_rhs_on_leaves(tree)
    orc.set_cc(M.I_RHS, ids, rhs)
    mg.set_cc(M.I_RHS, ids, rhs)
    hist_o, hist_g = [], []
    orc.fas_fmg(True, have_guess)
    M.mg_fas_fmg(tree, mg, True, have_guess)
    hist_o.append(orc.maxabs(M.I_TMP))
    hist_g.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(n_v):
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
        hist_o.append(orc.maxabs(M.I_TMP))
        hist_g.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    a = orc.get_cc(M.I_PHI, all_ids(tree)).reshape(-1)
    b = mg.get_cc(M.I_PHI, all_ids(tree)).reshape(-1)
    M.mg_destroy(mg)
    return np.array(hist_o), np.array(hist_g), a, b


@pytest.mark.parametrize("name", sorted(TREES))
def test_fmg_then_vcycles(name):
    tree = TREES[name]()
    ho, hg, a, b = run_both(tree, bc_mixed)
    assert ho[-1] < 0.1 * ho[0]  # it converges
    # per-cycle residual history: the residual is rhs - L(phi) with |L| ~ 1/dr^2, so compare with
    # the rounding level of that expression
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))


def test_fmg_with_guess_and_helmholtz():
    tree = T.corner_refined_tree(3, 8, 8, 4)
    ho, hg, a, b = run_both(tree, lambda nb, c: W.bc_table.__defaults__ and (W.AF_BC_DIRICHLET, 0.0) if False else
                            ((W.AF_BC_DIRICHLET, 0.0) if (nb - 1) // 2 == 2 else (W.AF_BC_NEUMANN, 0.0)),
                            have_guess=True, helmholtz_lambda=1.0e3)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho)


def test_sparse_prolongation_and_cycles():
    tree = T.corner_refined_tree(3, 8, 8, 3)
    ho, hg, a, b = run_both(tree, bc_mixed, prolongation_type=M.MG_PROLONG_SPARSE, n_cycle_down=1, n_cycle_up=3)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))


def test_highest_lvl_argument():
    tree = T.uniform_tree(3, 8, 8, 3)
    bc = W.bc_table(tree, bc_mixed)
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    ids, rhs = W.random_rhs_on_leaves(tree)
    orc.set_cc(M.I_RHS, ids, rhs)
    mg.set_cc(M.I_RHS, ids, rhs)
    orc.init_phi_rhs()
    mg.init_phi_rhs()
    orc.fas_vcycle(True, 2, False)
    M.mg_fas_vcycle(tree, mg, True, 2, False)
    a = orc.get_cc(M.I_PHI, all_ids(tree))
    b = mg.get_cc(M.I_PHI, all_ids(tree)).reshape(a.shape)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))
    M.mg_destroy(mg)


def test_errors_are_codes_not_aborts():
    tree = T.uniform_tree(3, 8, 8, 2)
    mg = M.mg_t(sides_bc=M.af_bc_neumann_zero)  # all-Neumann Poisson: singular coarse operator
    M.mg_init(tree, mg)
    with pytest.raises(M.AfmgError) as e:
        M.mg_fas_fmg(tree, mg, True, False)
    assert e.value.code == -5
    with pytest.raises(M.AfmgError):
        mg.gc_lvl(7)
    M.mg_destroy(mg)
    with pytest.raises(M.AfmgError):
        M.mg_fas_vcycle(tree, mg, True)


def test_field_solve_loop_matches_explicit_cycles():
    """afmg_field_solve = the FMG / V-cycle loop of field_compute (src/m_field.f90:491-524)."""
    tree = T.corner_refined_tree(3, 8, 8, 4)
    bc = W.bc_table(tree, bc_mixed)
    ids, rhs = W.random_rhs_on_leaves(tree)
    a = M.mg_t(sides_bc=bc)
    M.mg_init(tree, a)
    a.set_cc(M.I_RHS, ids, rhs)
    res, n_fmg, n_vc = M.field_solve(tree, a, False, 1e-4, num_vcycles=2)
    b = M.mg_t(sides_bc=bc)
    M.mg_init(tree, b)
    b.set_cc(M.I_RHS, ids, rhs)
    ref = []
    for _ in range(n_fmg):
        M.mg_fas_fmg(tree, b, True, True)
        ref.append(M.af_tree_maxabs_cc(tree, b, M.I_TMP))
    for _ in range(n_vc):
        M.mg_fas_vcycle(tree, b, True)
        ref.append(M.af_tree_maxabs_cc(tree, b, M.I_TMP))
    assert n_fmg >= 2 and np.array_equal(res, np.array(ref))
    assert res[n_fmg - 1] < 1e-4 and np.all(res[: n_fmg - 1] >= 1e-4)
    assert np.array_equal(a.get_cc(M.I_PHI, all_ids(tree)), b.get_cc(M.I_PHI, all_ids(tree)))
    with pytest.raises(M.AfmgError):
        M.field_solve(tree, a, False, 1e-300, max_residual=0.0, max_initial_iterations=3)
    M.mg_destroy(a)
    M.mg_destroy(b)


PERIODIC = {
    "periodic_xy_single_coarse_box": lambda: T.uniform_tree(3, 8, 8, 3, periodic=[True, True, False]),
    "periodic_x_multibox_refined": lambda: T.build_tree(3, 8, [16, 8, 8], 3, lambda l, ix, c: c[:, 2] < 0.55,
                                                        periodic=[True, False, False]),
}


@pytest.mark.parametrize("name", sorted(PERIODIC))
def test_periodic_domains(name):
    """tree%periodic (m_af_types.f90:345): neighbours wrap around, the coarse grid couples first and last
    cells (HYPRE_StructGridSetPeriodic, m_coarse_solver.f90:97-104)."""
    tree = PERIODIC[name]()
    ho, hg, a, b = run_both(tree, bc_mixed)
    assert ho[-1] < 0.1 * ho[0]
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))


def test_mixed_bc_types_on_one_domain_face():
    """Level-1 boxes may put different condition types on the same domain face (stencil_handle_boundaries is
    per box, m_coarse_solver.f90:442-491): e.g. a grounded plate covering part of the bottom.  The coarse
    operator is then not separable and the dense coarse solve takes over."""
    tree = T.build_tree(3, 8, [16, 16, 8], 3, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45)

    def bc_plate(nb, coords):
        centre_x = coords[:, :, 0].mean(axis=1)
        if nb == 5:  # low z: Dirichlet under the plate (x < 0.5), Neumann beside it
            ty = np.where(centre_x < 0.5, W.AF_BC_DIRICHLET, W.AF_BC_NEUMANN)
            return ty, np.where(ty[:, None] == W.AF_BC_DIRICHLET, 1.0, 0.0) * np.ones(coords.shape[:2])
        if nb == 6:
            return W.AF_BC_DIRICHLET, 0.0
        return W.AF_BC_NEUMANN, 0.0

    bc = W.bc_table(tree, bc_plate)
    on5 = bc.types[(bc.nbs == 5) & (tree.lvl[bc.ids] == 1)]
    assert len(set(on5.tolist())) == 2, "the case must mix types on one face of the coarse grid"
    ho, hg, a, b = run_both(tree, bc_plate)
    assert ho[-1] < 0.1 * ho[0]
    assert np.all(np.abs(ho - hg) <= 1e-9 * ho[0] + 1e-6 * ho), (ho, hg)
    assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))
