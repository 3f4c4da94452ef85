"""CPU: the library's host-side stencil builders (afmg_build_box_*, afivo_streamer_b200/stencils.py)
against the oracle's restatement of the same reference routines, bit for bit, on the trees and coefficient fields of
the GPU stencil suites -- so that what `build_stencils` hands to afmg_set_stencils is exactly what those suites ship --
and against analytic answers for the level-set distance search (mg_lsf_dist_linear / _gss, the gradient search)."""
import ctypes as C

import numpy as np
import pytest

from afivo_streamer_b200 import _lib
from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import stencils as S
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

import test_gpu_2d as G2
import test_gpu_stencils as G3
from util import all_ids, bc_mixed, stencils_from_oracle

CASES = {}
for name, (mk, kw) in G3.CASES.items():
    CASES["3d_" + name] = (mk, kw, G3.lsf_distances, bc_mixed)
for name, (mk, _bc, kw) in G2.CASES.items():
    if "eps" in kw or "lsf" in kw:
        CASES["2d_" + name] = (mk, kw, G2.lsf_distances2, _bc)


def dense(tree, a):
    """Per-box data of all_ids(tree) order -> array indexed by box id."""
    ids = all_ids(tree)
    out = np.zeros((tree.highest_id + 1,) + a.shape[1:])
    out[ids] = a
    return out


def default_prolong(nd):
    return (1, S.STENCIL_P248, np.array([9, 3, 3, 1]) / 16.0 if nd == 2 else np.array([27, 9, 9, 3, 9, 3, 3, 1]) / 64.0)


def same_entries(tree, got, want):
    assert [e["box_id"] for e in got] == [e["box_id"] for e in want]
    for g, w in zip(got, want):
        b = g["box_id"]
        assert g["tag"] == w["tag"], b
        assert g["op"][0] == w["op"][0], (b, "stype")
        assert np.array_equal(np.asarray(g["op"][1]).reshape(-1), np.asarray(w["op"][1]).reshape(-1)), (b, "operator")
        assert (g["f"] is None) == (w["f"] is None), (b, "f")
        if g["f"] is not None:
            assert np.array_equal(g["f"], w["f"]), (b, "f")
        assert bool(g["cyl"]) == bool(w["cyl"]), (b, "cylindrical_gradient")
        if "prolong" in w:
            gp = g.get("prolong") or default_prolong(tree.ndim)
            assert gp[0] == w["prolong"][0] and gp[1] == w["prolong"][1], (b, "prolongation kind", gp[:2], w["prolong"][:2])
            assert np.array_equal(np.asarray(gp[2]).reshape(-1), np.asarray(w["prolong"][2]).reshape(-1)), (b, "prolongation")
        else:
            assert "prolong" not in g


@pytest.mark.parametrize("name", sorted(CASES))
def test_builders_equal_the_oracle_bit_for_bit(name):
    mk, kw, dist_fn, bc_fn = CASES[name]
    kw = dict(kw)
    tree = mk()
    ids = all_ids(tree)
    eps, lsf = kw.pop("eps", None), kw.pop("lsf", None)
    custom = kw.pop("custom_prolong", False)
    kw.pop("lsf_boundary_value", None)
    orc = Oracle(tree, with_eps=eps is not None, **kw)
    orc.set_bc(W.bc_table(tree, bc_fn))
    eps_cc = lsf_data = None
    if eps is not None:
        e = eps(W.cell_centres(tree, ids, ghosts=True))
        orc.set_cc(M.I_EPS, ids, e)
        eps_cc = dense(tree, e)
    if lsf is not None:
        lids, dd = dist_fn(tree, lsf)
        orc.set_lsf_distances(lids, dd)
        pdd = None
        if custom:
            pids, p = G3.lsf_prolong_distances(tree, lsf)
            orc.set_lsf_prolong_distances(pids, p)
            pdd = {int(b): p[n].reshape(-1, tree.ndim + 1) for n, b in enumerate(pids)}
        ncell = tree.nc ** tree.ndim
        lsf_data = S.LsfData(np.asarray(lids, np.int32), dd.reshape(len(lids), ncell, 2 * tree.ndim),
                             np.zeros((len(lids), ncell)), {}, pdd)
    orc.mg_init()
    want = stencils_from_oracle(tree, orc)
    assert want
    got, _ = S.build_stencils(tree, eps_cc=eps_cc, lsf_data=lsf_data, lsf_use_custom_prolongation=custom, **kw)
    same_entries(tree, got, want)


@pytest.mark.parametrize("nd", [2, 3])
def test_linear_distances_equal_the_suites_own(nd):
    """store_lsf_distance_matrix with mg_lsf_dist_linear through the callback interface == the vectorised numpy
    statement the GPU suites use; every boundary cell lies inside the root mask."""
    if nd == 3:
        tree, lsf, ref = T.corner_refined_tree(3, 8, 8, 3), G3.lsf_sphere, G3.lsf_distances
    else:
        tree, lsf, ref = T.uniform_tree(2, 8, 8, 4), G2.lsf_circle, G2.lsf_distances2
    data = S.lsf_distances(tree, lsf)
    lids, dd = ref(tree, lsf)
    assert np.array_equal(data.ids, lids)
    want = dd.reshape(data.dd.shape)
    assert np.max(np.abs(data.dd - want)) < 1e-12
    for n, b in enumerate(data.ids):
        assert np.all(data.root_mask[int(b)][np.any(data.dd[n] < 1, axis=1)] == 1)
    # the built operator agrees with the one made from the suite's distances (up to the rounding of lsf itself)
    got, _ = S.build_stencils(tree, lsf_data=data)
    ref_data = S.LsfData(np.asarray(lids, np.int32), want, data.lsf_cells, {}, None)
    exp, _ = S.build_stencils(tree, lsf_data=ref_data)
    for g, e in zip(got, exp):
        np.testing.assert_allclose(g["op"][1], e["op"][1], rtol=1e-9)
        np.testing.assert_allclose(g["f"], e["f"], rtol=1e-9, atol=1e-6)


def test_custom_prolongation_distances_give_the_suites_stencils():
    tree = T.corner_refined_tree(3, 8, 8, 4)
    data = S.lsf_distances(tree, G3.lsf_sphere, custom_prolongation=True)
    got, _ = S.build_stencils(tree, lsf_data=data, lsf_use_custom_prolongation=True)
    lids, dd = G3.lsf_distances(tree, G3.lsf_sphere)
    pids, p = G3.lsf_prolong_distances(tree, G3.lsf_sphere)
    ref = S.LsfData(np.asarray(lids, np.int32), dd.reshape(len(lids), 8 ** 3, 6), data.lsf_cells, {},
                    {int(b): p[n].reshape(-1, 4) for n, b in enumerate(pids)})
    exp, _ = S.build_stencils(tree, lsf_data=ref, lsf_use_custom_prolongation=True)
    n_var = 0
    for g, e in zip(got, exp):
        gp, ep = g.get("prolong"), e.get("prolong")
        assert (gp is None) == (ep is None), g["box_id"]
        if gp is not None:
            assert gp[:2] == ep[:2]
            np.testing.assert_allclose(gp[2], ep[2], rtol=0, atol=1e-10)
            n_var += gp[0] == 2
    assert n_var > 0


def box_distances(lsf, opts, nd=3, nc=4, r_min=(0.0, 0.0, 0.0), dr=0.25):
    L = _lib.lib()
    ncell = nc ** nd
    cb = _lib.LSF_FN(lambda r, _u: float(lsf(np.array([r[d] for d in range(nd)]))))
    mask = np.zeros(ncell, np.uint8)
    dd = np.zeros((ncell, 2 * nd))
    nb = C.c_int32(0)
    rm = np.array(r_min[:nd], np.float64)
    drv = np.full(nd, dr)
    rc = L.afmg_build_box_lsf_distances(nd, nc, rm.ctypes.data_as(C.POINTER(C.c_double)),
                                        drv.ctypes.data_as(C.POINTER(C.c_double)), cb, None, C.byref(opts), None,
                                        mask.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        dd.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nb))
    assert rc == 0
    return mask, dd, nb.value


def test_gss_distance_finds_the_surface_of_a_sphere():
    """mg_lsf_dist_gss: relative distance from a cell centre to the sphere along each axis = the analytic
    intersection, to lsf_tol; the linear method is only exact for a level set that is linear along the line."""
    centre, R, dr, nc = np.array([0.5, 0.5, 0.5]), 0.3, 0.25, 4
    lsf = lambda r: np.linalg.norm(r - centre) - R
    _, dd, n = box_distances(lsf, S.lsf_opts(S.LSF_DIST_GSS), dr=dr)
    assert n > 0
    checked = 0
    for cell in range(nc ** 3):
        ijk = np.array([cell % nc, (cell // nc) % nc, cell // nc ** 2])
        a = (ijk + 0.5) * dr
        for m in range(6):
            step = np.zeros(3)
            step[m // 2] = dr if m % 2 else -dr
            la, lb = lsf(a), lsf(a + step)
            if la * lb < 0:  # one crossing: |a + t step - c| = R
                u = step / dr
                p = a - centre
                bq, cq = 2 * p @ u, p @ p - R * R
                roots = [(-bq + s * np.sqrt(bq * bq - 4 * cq)) / 2 for s in (-1, 1)]
                t = min(r for r in roots if 0 <= r <= dr) / dr
                assert abs(dd[cell, m] - max(t, 1e-4)) < 1e-7 / dr, (cell, m)
                checked += 1
            elif la > 0 and lb > 0:
                pass  # may or may not graze the sphere: covered by the test below
    assert checked > 10


def test_gss_distance_sees_a_thin_electrode_the_linear_method_misses():
    """Both end points outside (lsf_a * lsf_b > 0) with the electrode in between: golden-section search brackets
    the minimum, bisection finds the first crossing; mg_lsf_dist_linear returns 1 (m_af_multigrid.f90:1651-1684)."""
    dr, nc = 0.25, 4
    x0, half = 0.5, 0.03  # slab |x - 0.5| < 0.03 between the cell centres 0.375 and 0.625
    lsf = lambda r: abs(r[0] - x0) - half
    _, lin, n_lin = box_distances(lsf, S.lsf_opts(S.LSF_DIST_LINEAR), dr=dr)
    assert n_lin == 0 and np.all(lin == 1.0)
    _, dd, n = box_distances(lsf, S.lsf_opts(S.LSF_DIST_GSS), dr=dr)
    assert n == 2 * nc * nc
    want = (x0 - half - 0.375) / dr
    for cell in range(nc ** 3):
        i = cell % nc
        if i == 1:
            assert abs(dd[cell, 1] - want) < 1e-6 and np.all(np.delete(dd[cell], 1) == 1.0)
        elif i == 2:
            assert abs(dd[cell, 0] - want) < 1e-6 and np.all(np.delete(dd[cell], 0) == 1.0)
        else:
            assert np.all(dd[cell] == 1.0)


def test_gradient_search_finds_a_boundary_smaller_than_the_grid():
    """store_lsf_distance_matrix :1044-1074: an electrode tip thinner than the cell spacing, between cell centres, is
    found by walking down the gradient in steps of mg%lsf_length_scale; the distance is rescaled to the grid and
    assigned to the closest direction."""
    dr, nc = 0.25, 4
    a = np.array([0.375, 0.375, 0.375])  # centre of cell (2, 2, 2)
    centre = a + np.array([0.4 * dr, 0.0, 0.0])
    R = 0.3 * dr
    lsf = lambda r: np.linalg.norm(r - centre) - R
    _, dd0, n0 = box_distances(lsf, S.lsf_opts(), dr=dr)
    assert n0 == 0  # no sign change between any pair of neighbouring cell centres
    _, dd, n = box_distances(lsf, S.lsf_opts(length_scale=dr / 8), dr=dr)
    cell = 1 + nc * (1 + nc * 1)
    assert n >= 1
    assert abs(dd[cell, 1] - 0.1) < 1e-6, dd[cell]  # surface at 0.1 dr in +x
    assert np.all(np.delete(dd[cell], 1) == 1.0)


def test_tags_and_argument_checks():
    L = _lib.lib()
    one = np.ones(10 ** 3)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert L.afmg_build_box_tag(3, 8, None, 0) == 0
    assert L.afmg_build_box_tag(3, 8, None, 1) == S.MG_LSF_BOX
    assert L.afmg_build_box_tag(3, 8, dp(one), 0) == 0  # eps == 1: a normal box
    assert L.afmg_build_box_tag(3, 8, dp(one * (1 + 5e-9)), 0) == 0  # within 1e-8 of one
    assert L.afmg_build_box_tag(3, 8, dp(one * 2), 1) == S.MG_CEPS_BOX + S.MG_LSF_BOX
    var = one.copy()
    var[-1] = 1.5  # a ghost cell counts (minval / maxval over the whole array, :1131-1132)
    assert L.afmg_build_box_tag(3, 8, dp(var), 0) == S.MG_VEPS_BOX
    assert L.afmg_build_box_tag(1, 8, None, 0) == -1  # AFMG_ERR_ARG
    st, hf, cy = C.c_int32(), C.c_int32(), C.c_int32()
    v, f = np.zeros(7 * 512), np.zeros(512)
    dr = np.full(3, 0.1)
    assert L.afmg_build_box_operator(3, 8, 1, 0, dp(dr), None, None, None, dp(v), dp(f), C.byref(st), C.byref(hf),
                                     C.byref(cy)) == -1  # a normal box has no explicit stencil
    assert L.afmg_build_box_operator(3, 8, 1, S.MG_VEPS_BOX, dp(dr), None, None, None, dp(v), dp(f), C.byref(st),
                                     C.byref(hf), C.byref(cy)) == -1  # eps missing


# ---- the built-in electrode shapes (src/m_field.f90:686-904) ------------------------------------------------------

def test_rod_and_sphere_level_sets_are_signed_distances():
    rod = S.electrode("rod", 3, rod_r0=(0.5, 0.5, 0.0), rod_r1=(0.5, 0.5, 0.4), rod_radius=0.05)
    _, _, f = S._callback(rod, 3)
    assert abs(f([0.5, 0.7, 0.2]) - (0.2 - 0.05)) < 1e-15          # beside the rod: distance to the axis - radius
    assert abs(f([0.5, 0.5, 0.6]) - (0.2 - 0.05)) < 1e-15          # beyond the end: the semi-spherical cap
    assert abs(f([0.5, 0.5, 0.1]) + 0.05) < 1e-15                  # on the axis
    q = np.array([0.5 + 0.03, 0.5 + 0.04, 0.4 + 0.12])             # off-axis beyond the end: distance to r1
    assert abs(f(q) - (np.sqrt(0.03 ** 2 + 0.04 ** 2 + 0.12 ** 2) - 0.05)) < 1e-15
    sph = S.electrode("sphere", 2, rod_r0=(0.3, 0.4), rod_radius=0.1)
    _, _, g = S._callback(sph, 2)
    assert abs(g([0.6, 0.8]) - 0.4) < 1e-15


def test_conical_rod_surface_is_continuous_and_its_tip_sphere_touches_the_cone():
    """get_conical_rod_properties (:698-719): the tip sphere passes through the circle of radius tip_radius at the end
    of the cone; the surface of conical_rod_lsf_arg (:722-749) is continuous where cylinder, cone and tip meet."""
    r0, r1, R, rt, frac = np.array([0.5, 0.5, 0.0]), np.array([0.5, 0.5, 0.5]), 0.06, 0.02, 0.3
    el = S.electrode("rod_cone_top", 3, rod_r0=r0, rod_r1=r1, rod_radius=R, cone_tip_radius=rt, cone_length_frac=frac)
    _, _, f = S._callback(el, 3)
    angle = np.arctan((R - rt) / (frac * 0.5))
    assert abs(el.cone_tip_r_curvature - rt / np.cos(angle)) < 1e-15
    assert abs(el.cone_tip_center[2] - (0.5 - np.sin(angle) * rt / np.cos(angle))) < 1e-15
    # surface points: on the cylinder, at the cylinder / cone junction, half-way up the cone, on the rim of the tip
    for z, rad in ((0.1, R), (0.5 * (1 - frac), R), (0.5 * (1 - frac / 2), (R + rt) / 2), (0.5 - 1e-12, rt)):
        assert abs(f([0.5 + rad, 0.5, z])) < 1e-10, (z, rad)
    assert abs(f([0.5 + rt, 0.5, 0.5])) < 1e-12                     # the rim, evaluated by the spherical branch
    # away from the surface the reference's piecewise function jumps across frac = 1 (cone: distance to the axis minus
    # the local radius; tip: distance to the tip sphere); only the zero level set matters and that one is continuous
    assert abs(f([0.53, 0.5, 0.5 - 1e-9]) - 0.01) < 1e-8
    assert abs(f([0.53, 0.5, 0.5]) - (np.hypot(0.03, 0.5 - el.cone_tip_center[2]) - el.cone_tip_r_curvature)) < 1e-15
    assert abs(f([0.5, 0.5, 0.7]) - (0.7 - el.cone_tip_center[2] - el.cone_tip_r_curvature)) < 1e-15


def test_two_electrode_shapes_pick_the_nearest_for_the_potential():
    el = S.electrode("rod_rod", 3, rod_r0=(0.5, 0.5, 0.0), rod_r1=(0.5, 0.5, 0.3), rod_radius=0.05,
                     rod2_r0=(0.5, 0.5, 1.0), rod2_r1=(0.5, 0.5, 0.7), rod2_radius=0.04, current_voltage=2.5,
                     electrode2_grounded=1)
    _, _, f = S._callback(el, 3)
    assert abs(f([0.5, 0.5, 0.45]) - 0.10) < 1e-15 and abs(f([0.5, 0.5, 0.6]) - 0.06) < 1e-15
    pot = S.electrode_potential(el, np.array([[0.5, 0.5, 0.35], [0.5, 0.5, 0.65], [0.2, 0.5, 0.1]]))
    assert list(pot) == [2.5, 0.0, 2.5]
    sr = S.electrode("sphere_rod", 3, rod_r0=(0.5, 0.5, 0.2), rod_radius=0.1, rod2_r0=(0.5, 0.5, 1.0),
                     rod2_r1=(0.5, 0.5, 0.8), rod2_radius=0.05, current_voltage=-1.0, electrode_grounded=1)
    assert list(S.electrode_potential(sr, np.array([[0.5, 0.5, 0.35], [0.5, 0.5, 0.7]]))) == [0.0, -1.0]
    two = S.electrode("two_rod_cone_electrodes", 3, rod_r0=(0.5, 0.5, 0.0), rod_r1=(0.5, 0.5, 0.3), rod_radius=0.05,
                      cone_tip_radius=0.02, cone_length_frac=0.5, rod2_r0=(0.5, 0.5, 1.0), rod2_r1=(0.5, 0.5, 0.7),
                      rod2_radius=0.05, cone2_tip_radius=0.02, cone2_length_frac=0.5, current_voltage=1.0,
                      electrode2_grounded=1)
    _, _, g = S._callback(two, 3)
    assert abs(g([0.5, 0.5, 0.4]) - g([0.5, 0.5, 0.6])) < 1e-15     # mirror-symmetric pair
    assert list(S.electrode_potential(two, np.array([[0.5, 0.5, 0.4], [0.5, 0.5, 0.6]]))) == [1.0, 0.0]
    co = S.electrode("coaxial", 3, rod_radius=0.1, rod2_radius=0.45, domain_center=(0.5, 0.5, 0.5), current_voltage=3.0)
    _, _, h = S._callback(co, 3)
    assert abs(h([0.7, 0.5, 0.9]) - 0.1) < 1e-15 and abs(h([0.9, 0.5, 0.1]) - 0.05) < 1e-15
    assert list(S.electrode_potential(co, np.array([[0.62, 0.5, 0.3], [0.93, 0.5, 0.3]]))) == [3.0, 0.0]


def test_electrode_parameter_checks_are_the_references_error_stops():
    with pytest.raises(_lib.AfmgError):
        S.electrode("rod", 3, rod_r0=(0, 0, 0), rod_r1=(0, 0, 1), rod_radius=0.0)
    with pytest.raises(_lib.AfmgError):  # cone tip radius larger than the rod radius (:271-272)
        S.electrode("rod_cone_top", 3, rod_r0=(0, 0, 0), rod_r1=(0, 0, 1), rod_radius=0.1, cone_tip_radius=0.2,
                    cone_length_frac=0.5)
    with pytest.raises(_lib.AfmgError):
        S.electrode("rod_rod", 3, rod_r0=(0, 0, 0), rod_r1=(0, 0, 1), rod_radius=0.1)  # rod2_radius missing
    with pytest.raises(TypeError):
        S.electrode("rod", 3, cone_tip_center=(0, 0, 0))


def test_builtin_electrode_gives_the_same_stencils_as_the_equivalent_python_function():
    """The C-side level-set function goes through the same distance search as a Python callback."""
    tree = T.corner_refined_tree(3, 8, 8, 3)
    r0, r1, R = np.array([0.2, 0.25, 0.0]), np.array([0.2, 0.25, 0.3]), 0.07
    el = S.electrode("rod", 3, rod_r0=r0, rod_r1=r1, rod_radius=R)

    def rod(r):
        h = r1 - r0
        f = np.dot(r - r0, h)
        if f <= 0:
            v = r - r0
        elif f >= np.dot(h, h):
            v = r - r1
        else:
            v = r - (r0 + f / np.dot(h, h) * h)
        return np.sqrt(np.sum(v * v)) - R

    a, da = S.build_stencils(tree, lsf=el, lsf_options=S.lsf_opts(S.LSF_DIST_GSS))
    b, db = S.build_stencils(tree, lsf=rod, lsf_options=S.lsf_opts(S.LSF_DIST_GSS))
    assert len(a) == len(b) > 0 and np.array_equal(da.ids, db.ids)
    np.testing.assert_allclose(da.dd, db.dd, rtol=0, atol=1e-9)
    for x, y in zip(a, b):
        assert x["box_id"] == y["box_id"] and x["tag"] == y["tag"]
        np.testing.assert_allclose(x["op"][1], y["op"][1], rtol=1e-6)


def test_two_rods_at_different_potentials_on_the_oracle():
    """field_electrode_type = rod_rod (src/m_field.f90:280-294): level set min(rod 1, rod 2) and
    mg%lsf_boundary_function = rod_rod_get_potential (rod 1 at the applied voltage, rod 2 grounded), both evaluated
    by the library's C-side functions, distances by the library's search; solved by the oracle.  Physics: the
    potential obeys the maximum principle and each electrode's interior sits at its own potential."""
    V = 2.0
    t = T.build_tree(2, 8, [32, 32], 2, None)  # 4 x 4 coarse boxes (the electrodes are resolved on level 1), 64 x 64
    el = S.electrode("rod_rod", 2, rod_r0=(0.5, 0.0), rod_r1=(0.5, 0.3), rod_radius=0.06, rod2_r0=(0.5, 1.0),
                     rod2_r1=(0.5, 0.72), rod2_radius=0.06, current_voltage=V, electrode2_grounded=1)
    data = S.lsf_distances(t, el)
    nc = t.nc
    o = Oracle(t, lsf_boundary_value=99.0)  # the scalar must not be used where per-cell values are given
    # grounded side walls, zero flux at the ends
    o.set_bc(W.bc_table(t, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0) if nb <= 2 else (W.AF_BC_NEUMANN, 0.0)))
    o.set_lsf_distances(data.ids, data.dd.reshape(len(data.ids), -1))
    centres = W.cell_centres(t, data.ids, ghosts=False).reshape(len(data.ids), -1, 2)
    o.set_lsf_boundary_values(data.ids, S.electrode_potential(el, centres))
    o.mg_init()
    res = []
    for it in range(8):
        o.fas_fmg(True, it > 0)
        res.append(o.maxabs(M.I_TMP))
    assert res[-1] < 1e-8 * max(res[0], 1.0), res
    leaves = t.leaves(2).astype(np.int32)
    c = W.cell_centres(t, leaves, ghosts=False)
    phi = o.get_cc(M.I_PHI, leaves).reshape((len(leaves), nc + 2, nc + 2))[:, 1:-1, 1:-1]
    assert phi.min() > -1e-6 and phi.max() < V + 1e-6
    _, _, f = S._callback(el, 2)
    lsf = np.array([f(p) for p in c.reshape(-1, 2)]).reshape(phi.shape)
    dr = t.dr[leaves[0], 0]
    deep = lsf < -1.5 * dr
    in1, in2 = deep & (c[..., 1] < 0.5), deep & (c[..., 1] > 0.5)
    assert in1.sum() > 10 and in2.sum() > 10
    assert np.max(np.abs(phi[in1] - V)) < 1e-6 and np.max(np.abs(phi[in2])) < 1e-6
    mid = (np.abs(c[..., 0] - 0.5) < dr) & (np.abs(c[..., 1] - 0.5) < dr)  # between the tips: strictly in between
    assert np.all((phi[mid] > 0.2 * V) & (phi[mid] < 0.8 * V))


def test_unresolved_electrode_on_the_coarse_grid_is_refused_like_the_reference():
    """check_coarse_representation_lsf (m_af_multigrid.f90:2142-2161): "level set function not resolved on coarse
    grid" -- an 8 x 8 coarse grid does not see a rod of radius 0.01."""
    t = T.uniform_tree(2, 8, 8, 4)
    el = S.electrode("rod", 2, rod_r0=(0.53, 0.0), rod_r1=(0.53, 0.3), rod_radius=0.01)
    with pytest.raises(_lib.AfmgError, match="not resolved on coarse grid"):
        S.build_stencils(t, lsf=el)


@pytest.mark.parametrize("nd,coord,lvl", [(2, T.AF_XYZ, 4), (2, T.AF_CYL, 4), (3, T.AF_XYZ, 3)])
def test_poisson_lsf_test_spherical_electrode_against_its_analytic_potential(nd, coord, lvl):
    """afivo/examples/poisson_lsf_test.f90 (shape 1; :27-31, 223-240, 262-275): an electrode of radius 0.25 at
    potential 1 in the middle of the unit box (on the axis when cylindrical), Dirichlet = the analytic potential
    1 + log(d) in 2D, 2 - 1/d in 3D and cylindrical (d = r / radius) on the outer boundary; mg%lsf_dist =
    mg_lsf_dist_gss, mg%lsf_length_scale = 1e-3, uniform refinement.  The example prints residual, max error and rmse
    per FMG cycle: here the residual reaches rounding and the errors stay at the cut-cell level."""
    V, R = 1.0, 0.25
    r0 = np.full(nd, 0.5)
    if coord == T.AF_CYL:
        r0[0] = 0.0

    def sol(r):
        d = np.maximum(np.linalg.norm(r - r0, axis=-1) / R, 1e-300)
        out = V + (np.log(d) if (nd == 2 and coord == T.AF_XYZ) else 1 - 1 / d)
        return np.where(d < 1, V, out)

    t = T.build_tree(nd, 8, [8] * nd, lvl, None, coord_t=coord)
    # the example's level set is |r - r0| / R - 1; the built-in sphere |r - r0| - R has the same roots and the same
    # (scale-invariant) root mask, and is evaluated on the C side
    el = S.electrode("sphere", nd, rod_r0=r0, rod_radius=R)
    data = S.lsf_distances(t, el, S.lsf_opts(S.LSF_DIST_GSS, length_scale=1e-3))
    assert any(t.lvl[int(b)] == 1 for b in data.ids)  # resolved on the coarse grid
    o = Oracle(t, lsf_boundary_value=V)
    o.set_bc(W.bc_dirichlet_function(t, sol))
    o.set_lsf_distances(data.ids, data.dd.reshape(len(data.ids), -1))
    o.mg_init()
    res = []
    for it in range(10):
        o.fas_fmg(True, it > 0)
        res.append(o.maxabs(M.I_TMP))
    assert res[-1] < 1e-9 * res[0], res
    leaves = t.leaves(lvl).astype(np.int32)
    c = W.cell_centres(t, leaves, ghosts=True)
    err = np.abs(o.get_cc(M.I_PHI, leaves).reshape(c.shape[:-1]) - sol(c))[W.interior(t)]
    assert err.max() < 1.5e-2 and np.sqrt((err ** 2).mean()) < 2e-3, (err.max(), np.sqrt((err ** 2).mean()))


def test_stored_level_set_values_give_the_same_distances_as_evaluating_the_function():
    """afmg_build_box_lsf_distances reads box%cc(IJK, mg%i_lsf) when the caller has it (the reference's path) and
    evaluates mg%lsf at the cell centres otherwise: same mask, same distances; with a length scale the gradient search
    takes its starting sign from the same values."""
    L = _lib.lib()
    nd, nc, dr = 3, 8, 0.125
    centre = np.array([0.45, 0.55, 0.5])
    lsf = lambda r: np.linalg.norm(r - centre) - 0.2
    cb = _lib.LSF_FN(lambda r, _u: float(lsf(np.array([r[0], r[1], r[2]]))))
    rm, drv = np.zeros(3), np.full(3, dr)
    idx = np.arange(-1, nc + 1) + 0.5  # cell centres of cc(0:nc+1) along one dimension
    zz, yy, xx = np.meshgrid(idx * dr, idx * dr, idx * dr, indexing="ij")
    cc = np.linalg.norm(np.stack([xx, yy, zz], axis=-1) - centre, axis=-1) - 0.2
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for opts in (S.lsf_opts(), S.lsf_opts(S.LSF_DIST_GSS, length_scale=dr / 4)):
        out = []
        for stored in (None, np.ascontiguousarray(cc).reshape(-1)):
            mask, dd, nb = np.zeros(nc ** 3, np.uint8), np.zeros((nc ** 3, 6)), C.c_int32(0)
            rc = L.afmg_build_box_lsf_distances(nd, nc, dp(rm), dp(drv), cb, None, C.byref(opts),
                                                None if stored is None else dp(stored),
                                                mask.ctypes.data_as(C.POINTER(C.c_uint8)), dp(dd), C.byref(nb))
            assert rc == 0 and nb.value > 0
            out.append((mask.copy(), dd.copy(), nb.value))
        assert np.array_equal(out[0][0], out[1][0]) and out[0][2] == out[1][2]
        assert np.max(np.abs(out[0][1] - out[1][1])) < 1e-12
