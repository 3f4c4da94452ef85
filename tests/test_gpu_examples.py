"""GPU twins of the reference's self-checking examples (each mirrors an oracle known-answer test that passes on the
CPU, same problem set-up, so a failure points at the device path): poisson_helmholtz, poisson_cyl_analytic,
poisson_cyl_dielectric, poisson_lsf_test, helmholtz_variable_stencil (fully periodic domain), the two-rod electrode
problem, the native 2D electrode example, solve-from-.dat and the field_compute mirror.  Written at the end of round 1
as "staged" tests; first run on a B200 in round 2 (profiles/r02a_staged_tests.log: 14 passed) and since then part of
the regular `-m gpu` suite."""
import os
import subprocess

import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import stencils as S
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def leaves_of(t):
    return np.concatenate([t.leaves(l) for l in range(1, t.highest_lvl + 1)]).astype(np.int32)


def all_ids(t):
    return np.concatenate(t.lvl_ids).astype(np.int32)


@pytest.mark.parametrize("ndim", [2, 3])
def test_poisson_helmholtz_second_order(ndim):
    """twin of test_oracle_known_answers.py::test_poisson_helmholtz_gaussians_second_order"""
    g = W.Gaussians([[0.25] * ndim, [0.75] * ndim], 0.04)
    lam, errs = 1.0e3, []
    for lvls in (4, 5):
        t = T.uniform_tree(ndim, 8, 8, lvls)
        mg = M.mg_t(sides_bc=W.bc_dirichlet_function(t, g.value), helmholtz_lambda=lam)
        M.mg_init(t, mg)
        leaves = t.leaves(lvls).astype(np.int32)
        ctr = W.cell_centres(t, leaves, ghosts=True)
        mg.set_cc(M.I_RHS, leaves, g.laplacian(ctr) - lam * g.value(ctr))
        for it in range(5):
            M.mg_fas_fmg(t, mg, True, it > 0)
        assert M.af_tree_maxabs_cc(t, mg, M.I_TMP) < 1e-6 * np.max(np.abs(g.laplacian(ctr)))
        errs.append(np.max(np.abs(mg.get_cc(M.I_PHI, leaves) - g.value(ctr))[W.interior(t)]))
        M.mg_destroy(mg)
    assert 3.0 < errs[0] / errs[1] < 5.0, errs


def test_poisson_cyl_analytic():
    """twin of test_oracle_known_answers.py::test_poisson_cyl_analytic_gaussian_charge_on_the_axis (8 levels)"""
    from scipy.special import erf
    L = 1.25e-2
    sigma = 4e-4 * np.sqrt(0.5)
    src = np.array([0.0, 0.5]) * L
    eps0 = 8.85e-12
    Q = 3e18 * 1.6022e-19 * sigma ** 3 * np.sqrt(2 * np.pi) ** 3
    nc, max_lvl = 8, 8

    def rhs_f(r):
        return -Q * np.exp(-np.sum((r - src) ** 2, axis=-1) / (2 * sigma ** 2)) / (sigma ** 3 * np.sqrt(2 * np.pi) ** 3 * eps0)

    def sol(r):
        d = np.linalg.norm(r - src, axis=-1)
        small = d < np.sqrt(np.finfo(float).eps)
        safe = np.where(small, 1.0, d)
        return np.where(small, np.sqrt(2 / np.pi) / sigma, erf(safe * np.sqrt(0.5) / sigma) / safe) * Q / (4 * np.pi * eps0)

    def refine(l, ixs, ctr):
        dr = L / (nc * 2 ** (l - 1))
        off = (np.arange(nc) - (nc - 1) / 2) * dr
        gy, gx = np.meshgrid(off, off, indexing="ij")
        pts = ctr[:, None, :] + np.stack([gx, gy], axis=-1).reshape(1, -1, 2)
        return (dr * dr * np.max(np.abs(rhs_f(pts)), axis=1) > 1e-1) & (l < max_lvl)

    t = T.build_tree(2, nc, [nc, nc], max_lvl, refine, r_max=[L, L], coord_t=T.AF_CYL)
    mg = M.mg_t(sides_bc=W.bc_table(t, lambda nb, c: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, sol(c))))
    M.mg_init(t, mg)
    leaves = leaves_of(t)
    ctr = W.cell_centres(t, leaves, ghosts=True)
    mg.set_cc(M.I_RHS, leaves, rhs_f(ctr))
    res = []
    for it in range(10):
        M.mg_fas_fmg(t, mg, True, it > 0)
        res.append(M.af_tree_maxabs_cc(t, mg, M.I_TMP))
    assert res[-1] < 1e-7 * res[0], res
    rel = np.max(np.abs(mg.get_cc(M.I_PHI, leaves) - sol(ctr))[W.interior(t)]) / np.max(sol(ctr))
    assert rel < 3e-4, rel
    M.mg_destroy(mg)


def test_poisson_cyl_dielectric_adaptive():
    """twin of test_oracle_stencils.py::test_poisson_cyl_dielectric_on_the_adaptive_tree, set up by mg_set_operators_tree"""
    r0, s, nc = np.array([0.0, 0.25]), 0.1, 8

    def g(r):
        return np.exp(-np.sum((r - r0) ** 2, axis=-1) / s ** 2)

    def lap_cyl(r):
        x = (r - r0) / s
        return 4 / s ** 2 * (np.sum(x ** 2, axis=-1) - 1 - 0.5 * (r[..., 0] - r0[0]) / r[..., 0]) * g(r)

    def grad(r):
        return -2 * (r - r0) / s ** 2 * g(r)[..., None]

    def eps_f(r):
        return np.where((r[..., 0] < 0.5) & (r[..., 1] < 0.5), 100.0, 1.0)

    def refine(l, ixs, ctr):
        dr = 1.0 / (nc * 2 ** (l - 1))
        off = (np.arange(nc) - (nc - 1) / 2) * dr
        gy, gx = np.meshgrid(off, off, indexing="ij")
        pts = ctr[:, None, :] + np.stack([gx, gy], axis=-1).reshape(1, -1, 2)
        return (dr * dr * np.max(np.abs(lap_cyl(pts)), axis=1) > 1e-3) & (l < 7)

    t = T.build_tree(2, nc, [nc, nc], 7, refine, coord_t=T.AF_CYL)
    ids, leaves = all_ids(t), leaves_of(t)
    c = W.cell_centres(t, leaves, ghosts=True)
    e = eps_f(c)
    rhs = lap_cyl(c) * e
    dr = t.dr[leaves]
    fx = 0.5 * (c[:, 1:-1, :-1] + c[:, 1:-1, 1:])
    q = (e[:, 1:-1, 1:] - e[:, 1:-1, :-1]) * grad(fx)[..., 0] / dr[:, 0][:, None, None]
    w = e[:, 1:-1, 1:] / (e[:, 1:-1, :-1] + e[:, 1:-1, 1:])
    rhs[:, 1:-1, 1:] += w * q
    rhs[:, 1:-1, :-1] += (1 - w) * q
    fy = 0.5 * (c[:, :-1, 1:-1] + c[:, 1:, 1:-1])
    q = (e[:, 1:, 1:-1] - e[:, :-1, 1:-1]) * grad(fy)[..., 1] / dr[:, 1][:, None, None]
    w = e[:, 1:, 1:-1] / (e[:, :-1, 1:-1] + e[:, 1:, 1:-1])
    rhs[:, 1:, 1:-1] += w * q
    rhs[:, :-1, 1:-1] += (1 - w) * q
    mg = M.mg_t(sides_bc=W.bc_table(t, lambda nb, cc: (W.AF_BC_NEUMANN, 0.0) if nb == 1 else (W.AF_BC_DIRICHLET, g(cc))))
    M.mg_init(t, mg)
    eps_cc = np.ones((t.highest_id + 1, nc + 2, nc + 2))
    eps_cc[ids] = eps_f(W.cell_centres(t, ids, ghosts=True))
    entries, _ = M.mg_set_operators_tree(t, mg, eps_cc=eps_cc)
    assert {en["tag"] for en in entries} == {2, 4}
    mg.set_cc(M.I_RHS, leaves, rhs)
    res = []
    for it in range(10):
        M.mg_fas_fmg(t, mg, True, it > 0)
        res.append(M.af_tree_maxabs_cc(t, mg, M.I_TMP))
    assert res[-1] < 1e-8 * res[0], res
    err = np.max(np.abs(mg.get_cc(M.I_PHI, leaves) - g(c))[W.interior(t)])
    assert err < 3e-4, err
    M.mg_destroy(mg)


@pytest.mark.parametrize("nd,coord,lvl", [(2, T.AF_XYZ, 4), (2, T.AF_CYL, 4), (3, T.AF_XYZ, 3)])
def test_poisson_lsf_test_spherical_electrode(nd, coord, lvl):
    """twin of test_stencil_builders.py::test_poisson_lsf_test_spherical_electrode_against_its_analytic_potential"""
    V, R = 1.0, 0.25
    r0 = np.full(nd, 0.5)
    if coord == T.AF_CYL:
        r0[0] = 0.0

    def sol(r):
        d = np.maximum(np.linalg.norm(r - r0, axis=-1) / R, 1e-300)
        return np.where(d < 1, V, V + (np.log(d) if (nd == 2 and coord == T.AF_XYZ) else 1 - 1 / d))

    t = T.build_tree(nd, 8, [8] * nd, lvl, None, coord_t=coord)
    mg = M.mg_t(sides_bc=W.bc_dirichlet_function(t, sol), lsf_boundary_value=V)
    M.mg_init(t, mg)
    M.mg_set_operators_tree(t, mg, lsf=S.electrode("sphere", nd, rod_r0=r0, rod_radius=R),
                            lsf_options=S.lsf_opts(S.LSF_DIST_GSS, length_scale=1e-3))
    res = []
    for it in range(10):
        M.mg_fas_fmg(t, mg, True, it > 0)
        res.append(M.af_tree_maxabs_cc(t, mg, M.I_TMP))
    assert res[-1] < 1e-9 * res[0], res
    leaves = t.leaves(lvl).astype(np.int32)
    c = W.cell_centres(t, leaves, ghosts=True)
    err = np.abs(mg.get_cc(M.I_PHI, leaves) - sol(c))[W.interior(t)]
    assert err.max() < 1.5e-2 and np.sqrt((err ** 2).mean()) < 2e-3
    M.mg_destroy(mg)


@pytest.mark.parametrize("ndim", [2, 3])
def test_implicit_diffusion_in_a_fully_periodic_domain(ndim):
    """twin of test_oracle_known_answers.py::test_implicit_diffusion_in_a_periodic_domain_follows_the_discrete_decay;
    a fully periodic single-box coarse grid with lambda > 0 (helmholtz_variable_stencil.f90)"""
    dlen = 2 * np.arccos(-1.0)
    t = T.build_tree(ndim, 8, [8] * ndim, 3, None, r_max=[dlen] * ndim, periodic=[True] * ndim)
    h = t.dr[t.leaves(3)[0], 0]
    mu = (4 / h ** 2) * 2 * np.sin(h / 2) ** 2
    ids, leaves = all_ids(t), t.leaves(3).astype(np.int32)
    rr = W.cell_centres(t, ids, ghosts=True)
    mode = np.cos(rr[..., 0]) * np.cos(rr[..., 1])
    dt = 0.1
    mg = M.mg_t(sides_bc=W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0)), helmholtz_lambda=1 / dt)
    M.mg_init(t, mg)
    mg.set_cc(M.I_PHI, ids, 1 + mode)
    sel = np.isin(ids, leaves)
    amp = 1.0
    for step in range(3):
        mg.set_cc(M.I_RHS, leaves, -mg.get_cc(M.I_PHI, leaves) / dt)
        for it in range(4):
            M.mg_fas_fmg(t, mg, True, True)
        amp /= 1 + dt * mu
        err = np.max(np.abs(mg.get_cc(M.I_PHI, leaves) - (1 + amp * mode[sel]))[W.interior(t)])
        assert err < 1e-9, (step, err)
    M.mg_destroy(mg)


def test_two_rods_at_different_potentials():
    """twin of test_stencil_builders.py::test_two_rods_at_different_potentials_on_the_oracle (afmg_set_lsf_boundary_values)"""
    V = 2.0
    t = T.build_tree(2, 8, [32, 32], 2, None)
    el = S.electrode("rod_rod", 2, rod_r0=(0.5, 0.0), rod_r1=(0.5, 0.3), rod_radius=0.06, rod2_r0=(0.5, 1.0),
                     rod2_r1=(0.5, 0.72), rod2_radius=0.06, current_voltage=V, electrode2_grounded=1)
    mg = M.mg_t(sides_bc=W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0) if nb <= 2 else (W.AF_BC_NEUMANN, 0.0)),
                lsf_boundary_value=99.0)
    M.mg_init(t, mg)
    _, data = M.mg_set_operators_tree(t, mg, lsf=el)
    centres = W.cell_centres(t, data.ids, ghosts=False).reshape(len(data.ids), -1, 2)
    mg.set_lsf_boundary_values(data.ids, S.electrode_potential(el, centres))
    for it in range(8):
        M.mg_fas_fmg(t, mg, True, it > 0)
    leaves = t.leaves(2).astype(np.int32)
    c = W.cell_centres(t, leaves, ghosts=False)
    phi = mg.get_cc(M.I_PHI, leaves)[:, 1:-1, 1:-1]
    assert phi.min() > -1e-6 and phi.max() < V + 1e-6
    _, _, f = S._callback(el, 2)
    lsf = np.array([f(p) for p in c.reshape(-1, 2)]).reshape(phi.shape)
    deep = lsf < -1.5 * t.dr[leaves[0], 0]
    in1, in2 = deep & (c[..., 1] < 0.5), deep & (c[..., 1] > 0.5)
    assert np.max(np.abs(phi[in1] - V)) < 1e-6 and np.max(np.abs(phi[in2])) < 1e-6
    M.mg_destroy(mg)


@pytest.mark.parametrize("cyl", [0, 1])
def test_native_electrode_example_runs(cyl):
    """tools/electrode_example_2d on the device: residual falls by 1e6 within ten FMG cycles, 0 <= phi <= 1 (the
    Cartesian variant was run by hand with the round's last GPU seconds: profiles/r01g_electrode_example_2d.txt)"""
    exe = os.path.join(ROOT, "tools", "electrode_example_2d")
    out = subprocess.run([exe] + (["cyl"] if cyl else []), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.strip().splitlines()]
    res = [float(r[1]) for r in rows[:10]]
    assert res[5] < 1e-6 * res[0], res
    lo, hi = float(rows[-1][4]), float(rows[-1][5])
    assert lo > -1e-3 and hi < 1 + 1e-3


def test_native_solve_dat_matches_python_solve_dat(tmp_path):
    """tools/solve_dat (C++, include/afmg_dat.hpp) and tools/solve_dat.py print the same residual history for an
    eps + electrode file written by the Python writer."""
    import sys
    import test_gpu_stencils as G3
    from afivo_streamer_b200 import datfile as D
    from dat_util import make_dat
    from oracle.oracle import Oracle
    from util import bc_mixed
    tree = T.corner_refined_tree(3, 8, 8, 3)
    ids = all_ids(tree)
    bc = W.bc_table(tree, bc_mixed)
    orc = Oracle(tree, with_eps=True, lsf_boundary_value=0.9)
    orc.set_bc(bc)
    e = np.ones((tree.highest_id + 1, tree.box_len))
    e[ids] = G3.eps_smooth(W.cell_centres(tree, ids, ghosts=True)).reshape(len(ids), -1)
    orc.set_cc(3, ids, e[ids])
    lsf_dd = G3.lsf_distances(tree, G3.lsf_sphere)
    orc.set_lsf_distances(*lsf_dd)
    orc.mg_init()
    rng = np.random.default_rng(5)
    orc.set_cc(1, ids, rng.uniform(-1, 1, (len(ids), tree.box_len)))
    path = str(tmp_path / "t.dat")
    D.write_tree(path, make_dat(tree, orc, bc, extra_cc={"eps": e}, lsf_dd=lsf_dd))
    py = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "solve_dat.py"), path, "--eps", "eps",
                         "--lsf-boundary-value", "0.9"], capture_output=True, text=True, timeout=300)
    cc = subprocess.run([os.path.join(ROOT, "tools", "solve_dat"), path, "3", "eps", "5", "0.9"], capture_output=True,
                        text=True, timeout=300)
    assert py.returncode == 0 and cc.returncode == 0, py.stderr + cc.stderr
    pick = lambda out: [float(ln.split()[-1]) for ln in out.splitlines() if "residual" in ln]
    a, b = pick(py.stdout), pick(cc.stdout)
    assert len(a) == len(b) == 6
    assert np.allclose(a, b, rtol=1e-6, atol=1e-12), (a, b)


def test_field_compute_mirror():
    """mg.field_compute = field_compute (src/m_field.f90:448-528) composed from afmg_field_solve and
    afmg_field_from_potential: converges below its own threshold and leaves the field norm on the device."""
    t = T.corner_refined_tree(3, 8, 8, 4)
    V = 2.0e3
    mg = M.mg_t(sides_bc=W.bc_field_homogeneous(t, V))
    M.mg_init(t, mg)
    ids, rhs = W.random_rhs_on_leaves(t)
    mg.set_cc(M.I_RHS, ids, 1e6 * rhs)
    res, n_fmg, n_vc = M.field_compute(t, mg, V, False)
    thr = M.field_residual_threshold(t, M.af_tree_maxabs_cc(t, mg, M.I_RHS), V)
    assert n_fmg >= 1 and 1 <= n_vc <= 2 and res[n_fmg - 1] < max(thr, res[0])
    res2, n_fmg2, n_vc2 = M.field_compute(t, mg, V, True)
    assert n_fmg2 == 0 and n_vc2 >= 1 and res2[-1] <= res[-1] * 1.01
    fld = mg.get_cc(M.I_FLD, ids)
    assert np.isfinite(fld).all() and fld.max() > V * 0.5  # ~ V / L = 2e3 in a unit box
    M.mg_destroy(mg)
