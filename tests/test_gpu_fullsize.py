"""-m gpu: BASELINE.json's full sizes.  S1r (256^3) and S2 (9-level channel tree) directly against the oracle
(seconds on the host cores); S1 = poisson_benchmark 16 16 5 and the S3s shell-refined octree (1.1e8 cells) through
size-independent properties (linearity, fixed point, convergence)."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

pytestmark = pytest.mark.gpu


def leaves_of(tree):
    return np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)


def solve(tree, bc, ids, rhs_interior, n_v=4):
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    mg.set_cc_interior(M.I_RHS, ids, rhs_interior)
    hist = []
    M.mg_fas_fmg(tree, mg, True, False)
    hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(n_v):
        M.mg_fas_vcycle(tree, mg, True)
        hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    return mg, np.array(hist)


def test_interior_upload_equals_full_upload():
    tree = T.corner_refined_tree(3, 8, 8, 4)
    ids, rhs = W.random_rhs_on_leaves(tree)
    a = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, a)
    a.set_cc(M.I_RHS, ids, rhs)
    b = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, b)
    b.set_cc_interior(M.I_RHS, ids, rhs[W.interior(tree)])
    for m in (a, b):
        M.mg_fas_fmg(tree, m, True, False)
    assert np.array_equal(a.get_cc(M.I_PHI, ids), b.get_cc(M.I_PHI, ids))
    assert np.array_equal(a.get_cc(M.I_RHS, ids)[W.interior(tree)], rhs[W.interior(tree)])
    M.mg_destroy(a)
    M.mg_destroy(b)


def test_s1_full_size_convergence_and_linearity():
    tree = T.uniform_tree(3, 16, 16, 5)  # 4681 boxes (afivo/tests/answers/test_refinement_3d), 256^3
    assert tree.n_boxes == 4681
    bc = W.bc_dirichlet_zero(tree)
    ids = leaves_of(tree)
    rng = np.random.default_rng(1)
    r1 = rng.uniform(-1, 1, (len(ids), 16, 16, 16))
    r2 = rng.uniform(-1, 1, (len(ids), 16, 16, 16))
    m1, h1 = solve(tree, bc, ids, r1)
    # residual falls by about an order of magnitude per V-cycle, monotonically (poisson_basic behaviour)
    assert np.all(h1[1:] < 0.25 * h1[:-1]), h1
    sample = ids[:: max(1, len(ids) // 64)]
    p1 = m1.get_cc(M.I_PHI, sample)
    M.mg_destroy(m1)
    m2, _ = solve(tree, bc, ids, r2)
    p2 = m2.get_cc(M.I_PHI, sample)
    M.mg_destroy(m2)
    # the cycle is a linear map of (phi, rhs) for homogeneous boundary conditions
    m3, _ = solve(tree, bc, ids, 2.0 * r1 - 0.5 * r2)
    p3 = m3.get_cc(M.I_PHI, sample)
    M.mg_destroy(m3)
    scale = np.max(np.abs(p1)) + np.max(np.abs(p2))
    assert np.max(np.abs(p3 - (2.0 * p1 - 0.5 * p2))) <= 1e-10 * scale


def test_s1_constant_solution_is_a_fixed_point():
    # rhs = 0 with Dirichlet value 1 on every face: phi == 1 is reproduced exactly by every operation
    tree = T.uniform_tree(3, 16, 16, 5)
    bc = W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, 1.0))
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(6):
        M.mg_fas_vcycle(tree, mg, True)
    ids = leaves_of(tree)[::97]
    phi = mg.get_cc(M.I_PHI, ids)
    assert np.max(np.abs(phi - 1.0)) <= 1e-12
    assert M.af_tree_maxabs_cc(tree, mg, M.I_TMP) <= 1e-7  # residual = O(eps / dr^2)
    M.mg_destroy(mg)


def test_s3s_refined_octree_converges():
    tree = T.shell_tree(16, 16, 5)  # 1.09e8 cells, closed refinement boundary
    bc = W.bc_field_homogeneous(tree, 1.0)
    ids = leaves_of(tree)
    rng = np.random.default_rng(2)
    rhs = rng.uniform(-1, 1, (len(ids), 16, 16, 16))
    mg, h = solve(tree, bc, ids, rhs, n_v=5)
    assert np.all(h[1:] < 0.3 * h[:-1]), h
    # the potential obeys the boundary data: 0 at z = 0, 1 at z = 1 (ghost cell = 2 b - interior)
    top = [b for b in tree.lvl_ids[-2] if tree.neighbors[b, 5] < 0 and not tree.has_children(np.array([b]))[0]][:4]
    phi = mg.get_cc(M.I_PHI, np.array(top, np.int32))
    assert np.allclose(0.5 * (phi[:, 17, 1:-1, 1:-1] + phi[:, 16, 1:-1, 1:-1]), 1.0, atol=1e-12)
    M.mg_destroy(mg)


def parity_with_oracle(tree, rhs, n_v, n_finest):
    """1 FMG + n_v V-cycles on the GPU and on the CPU oracle from the same leaf rhs (interior cells): the residual
    histories agree to the rounding floor of rhs - L(phi), the potentials to 1e-10 relative max-norm."""
    from oracle.oracle import Oracle
    bc = W.bc_field_homogeneous(tree, 1.0)
    ids = leaves_of(tree)
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    for q0 in range(0, len(ids), 4096):  # slabs bound the host memory of the ghost-padded copy
        sub = ids[q0:q0 + 4096]
        full = W.box_array(tree, len(sub))
        full[W.interior(tree)] = rhs[q0:q0 + 4096]
        orc.set_cc(M.I_RHS, sub, full)
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    mg.set_cc_interior(M.I_RHS, ids, rhs)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(n_v):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    ho.append(orc.maxabs(M.I_TMP))
    hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    ho, hg = np.array(ho), np.array(hg)
    assert np.all(ho[1:] < 0.3 * ho[:-1]) or ho[-1] < 1e-8 * ho[0], ho
    floor = 16 * np.finfo(float).eps * 7 * float(n_finest) ** 2
    assert np.all(np.abs(ho - hg) <= floor + 1e-6 * ho), (ho, hg, floor)
    worst = scale = 0.0
    allb = np.concatenate(tree.lvl_ids).astype(np.int32)
    for q0 in range(0, len(allb), 2048):
        sub = allb[q0:q0 + 2048]
        a = orc.get_cc(M.I_PHI, sub)
        b = mg.get_cc(M.I_PHI, sub).reshape(len(sub), -1)
        worst = max(worst, float(np.max(np.abs(a - b))))
        scale = max(scale, float(np.max(np.abs(a))))
    # one-shot interior download of all leaves (many staging chunks, copies overlapping the pack kernels)
    inner = mg.get_cc_interior(M.I_PHI, ids)
    for q0 in range(0, len(ids), 4096):
        a = orc.get_cc(M.I_PHI, ids[q0:q0 + 4096]).reshape((-1,) + (tree.nc + 2,) * 3)[W.interior(tree)]
        assert np.max(np.abs(a.reshape(inner[q0:q0 + 4096].shape) - inner[q0:q0 + 4096])) <= 1e-10 * scale
    del inner
    M.mg_destroy(mg)
    assert worst <= 1e-10 * scale, worst / scale
    return ho, hg, worst / scale


def test_s3s_parity_with_oracle():
    """The bounded CPU sample of the benchmark (S3s: 256^3 uniform + refined shell, 1.1e8 cells, closed refinement
    boundary): 1 FMG + 5 V-cycles, GPU vs oracle on all boxes of all levels."""
    tree = T.shell_tree(16, 16, 5)
    rhs = np.random.default_rng(2).uniform(-1, 1, (len(leaves_of(tree)), 16, 16, 16))
    parity_with_oracle(tree, rhs, 5, 512)


def _host_ram_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


@pytest.mark.skipif(_host_ram_gb() < 110, reason="S3 on the CPU oracle needs ~80 GB of host memory")
def test_s3_benchmark_tree_parity_with_oracle():
    """THE benchmarked configuration (bench.py default, BASELINE.json configs[4]): S3, 1.04e9 cells, 253 449 boxes.
    1 FMG + 2 V-cycles on the GPU and on the oracle (36 GB for its three variables), all boxes compared."""
    tree = T.shell_tree(16, 16, 6)
    n = len(leaves_of(tree))
    rhs = np.empty((n, 16, 16, 16))
    rng = np.random.default_rng(3)
    for q0 in range(0, n, 8192):
        rhs[q0:q0 + 8192] = rng.uniform(-1, 1, rhs[q0:q0 + 8192].shape)
    parity_with_oracle(tree, rhs, 2, 1024)


def test_s1r_full_size_parity_with_oracle():
    """BASELINE.md S1r at its full size (256^3 = poisson_benchmark 16 16 5): random rhs on the leaves,
    field_bc_homogeneous with voltage 1 (Dirichlet 0 / 1 in z, Neumann 0 in x, y), 1 FMG + 10 V-cycles: same
    per-cycle residual history, potential within 1e-10 relative max-norm of the CPU oracle (north_star tolerance).
    The oracle needs a few seconds on the host cores here (1.7e7 finest cells)."""
    from oracle.oracle import Oracle
    tree = T.uniform_tree(3, 16, 16, 5)
    bc = W.bc_field_homogeneous(tree, 1.0)
    ids = leaves_of(tree)
    rhs = np.random.default_rng(12345).uniform(-1.0, 1.0, (len(ids), 16, 16, 16))
    full = W.box_array(tree, len(ids))
    full[W.interior(tree)] = rhs
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    orc.set_cc(M.I_RHS, ids, full)
    del full
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    mg.set_cc_interior(M.I_RHS, ids, rhs)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(10):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    ho.append(orc.maxabs(M.I_TMP))
    hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    ho, hg = np.array(ho), np.array(hg)
    assert ho[-1] < 1e-8 * ho[0], ho
    # the residual is rhs - L(phi): seven terms of size |phi| / dr^2 = 6.6e4 cancel to O(1e-10) at the end, so the
    # two histories can only agree to the rounding of that sum (the potentials themselves agree to 1e-10 below)
    floor = 16 * np.finfo(float).eps * 7 * 256.0 ** 2 * 1.0
    assert np.all(np.abs(ho - hg) <= floor + 1e-6 * ho), (ho, hg, floor)
    worst = 0.0
    scale = 0.0
    for q0 in range(0, len(ids), 512):  # compare in slabs to bound host memory
        sub = ids[q0:q0 + 512]
        a = orc.get_cc(M.I_PHI, sub)
        b = mg.get_cc(M.I_PHI, sub).reshape(len(sub), -1)
        worst = max(worst, float(np.max(np.abs(a - b))))
        scale = max(scale, float(np.max(np.abs(a))))
    assert worst <= 1e-10 * scale, worst / scale
    M.mg_destroy(mg)


def test_s2_field_and_helmholtz_patterns_parity():
    """BASELINE.md S2 at its full size (standard_3d-like channel-refined tree, nc = 8, 9 levels, 8905 boxes): the field
    pattern of field_compute (FMG from scratch, then 2 V-cycles with the residual test, src/m_field.f90:491-524)
    and the Helmholtz pattern of photoi_helmh_compute (3 modes, <= 10 FMG each, src/m_photoi_helmh.f90:162-204),
    GPU against the CPU oracle: same residual histories / FMG counts, results within 1e-10."""
    from oracle.oracle import Oracle
    tree = T.channel_tree(8, 8, 9, 3)
    assert tree.highest_lvl == 9
    all_ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    ids, rhs = W.random_rhs_on_leaves(tree)
    # ---- field pattern
    bc = W.bc_field_homogeneous(tree, 1.0)
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    orc.set_cc(M.I_RHS, ids, rhs)
    mg = M.mg_t(sides_bc=bc)
    M.mg_init(tree, mg)
    mg.set_cc(M.I_RHS, ids, rhs)
    ho, hg = [], []
    orc.fas_fmg(True, False)
    M.mg_fas_fmg(tree, mg, True, False)
    for _ in range(2):
        ho.append(orc.maxabs(M.I_TMP))
        hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        orc.fas_vcycle(True)
        M.mg_fas_vcycle(tree, mg, True)
    ho.append(orc.maxabs(M.I_TMP))
    hg.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    ho, hg = np.array(ho), np.array(hg)
    dr_min = float(np.min(tree.dr[all_ids]))
    floor = 16 * np.finfo(float).eps * 7 / dr_min ** 2
    assert np.all(np.abs(ho - hg) <= floor + 1e-6 * ho), (ho, hg, floor)
    a = orc.get_cc(M.I_PHI, all_ids)
    b = mg.get_cc(M.I_PHI, all_ids).reshape(len(all_ids), -1)
    assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(a))
    M.mg_destroy(mg)
    del orc
    # ---- Helmholtz pattern (Bourdon-3 at 1 bar, 20 % O2, 2 cm domain scaled to the unit cube)
    lambdas = np.array([4147.85, 10950.93, 66755.67]) * 0.2 * 0.02
    coeffs = np.array([1117314.935, 28692377.5, 2748842283.0]) * (0.2 * 0.02) ** 2
    hbc = W.bc_table(tree, M.photoi_helmh_bc)
    leaves = tree.children[all_ids, 0] == 0
    photo = np.zeros((len(all_ids), tree.box_len))
    ncyc_o = []
    max_rhs = None
    for lam, c in zip(lambdas, coeffs):
        o = Oracle(tree, helmholtz_lambda=lam ** 2, prolongation_type=M.MG_PROLONG_LINEAR)
        o.set_bc(hbc)
        o.mg_init()
        o.set_cc(M.I_RHS, ids, rhs)
        if max_rhs is None:
            max_rhs = max(o.maxabs(M.I_RHS), np.sqrt(np.finfo(float).eps))
        n = 0
        for n in range(1, 11):
            o.fas_fmg(True, True)
            if o.maxabs(M.I_TMP) / max_rhs < 1e-2:
                break
        ncyc_o.append(n)
        phi = o.get_cc(M.I_PHI, all_ids)
        photo[leaves] = photo[leaves] - c * phi[leaves]
        del o
    mgs = []
    for lam in lambdas:
        m = M.mg_t(sides_bc=hbc, helmholtz_lambda=float(lam ** 2), prolongation_type=M.MG_PROLONG_LINEAR)
        M.mg_init(tree, m)
        mgs.append(m)
    mgs[0].set_cc(M.I_RHS, ids, rhs)
    ncyc, _ = M.photoi_helmh_compute(tree, mgs, coeffs, 10, 1.0e-2)
    assert list(ncyc) == ncyc_o, (ncyc, ncyc_o)
    got = mgs[0].get_cc(M.I_PHOTO, all_ids).reshape(len(all_ids), -1)
    assert np.max(np.abs(got - photo)) <= 1e-10 * np.max(np.abs(photo))
    for m in mgs:
        M.mg_destroy(m)
