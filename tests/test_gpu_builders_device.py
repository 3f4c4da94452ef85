"""-m gpu: stencil construction ON THE DEVICE (afmg_build_stencils_device, csrc/builders_dev.cuh)
against the library's host builders (afmg_build_box_*, themselves == the oracle's restatement of the reference's
mg_set_operators_lvl, tests/test_stencil_builders.py): same tags, same stencil kinds and -- for the permittivity
stencils and the level-set stencils with the linear distance rule -- the same coefficients bit for bit; with the
golden-section / bisection rule (transcendental functions differ in the last place between libm and the device) to
1e-12.  Then the two set-ups must give the same solves."""
import numpy as np
import pytest

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import stencils as S
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W

import test_gpu_stencils as G3
from util import all_ids, bc_mixed

pytestmark = pytest.mark.gpu


def host_records(tree, entries, data):
    """the host builders' output in the layout of afmg_built_stencils: v | f | pv | dd per tagged box"""
    nc = tree.nc
    ncell = nc ** 3
    dd_of = {int(b): data.dd[n] for n, b in enumerate(data.ids)} if data is not None else {}
    out = {}
    for e in entries:
        rec = dict(tag=e["tag"], op_stype=0, has_f=0, p_stype=0, v=None, f=None, pv=None, dd=None)
        if e["op"] is not None:
            st, co = e["op"]
            rec["op_stype"] = st
            rec["v"] = np.asarray(co, float).reshape(-1)
        if e["f"] is not None:
            rec["has_f"], rec["f"] = 1, np.asarray(e["f"], float).reshape(-1)
        if "prolong" in e:
            pst, psh, pco = e["prolong"]
            assert psh == S.STENCIL_P234
            rec["p_stype"], rec["pv"] = pst, np.asarray(pco, float).reshape(-1)
        if e["box_id"] in dd_of:
            rec["dd"] = np.asarray(dd_of[e["box_id"]], float).reshape(ncell * 6)
        out[e["box_id"]] = rec
    return out


def compare(tree, host, dev, exact=True):
    ids, tags, meta, blob = dev
    ncell = tree.nc ** 3
    assert sorted(host) == sorted(int(b) for b in ids)
    worst = 0.0
    for q, b in enumerate(ids):
        h = host[int(b)]
        assert h["tag"] == tags[q], (b, h["tag"], tags[q])
        assert (h["op_stype"], h["has_f"], h["p_stype"]) == tuple(meta[q, :3]), (b, h, meta[q])
        rec = blob[q]
        parts = {"v": rec[:7 * ncell], "f": rec[7 * ncell:8 * ncell], "pv": rec[8 * ncell:12 * ncell], "dd": rec[12 * ncell:]}
        for k, a in parts.items():
            if h[k] is None:
                continue
            want = h[k]
            got = a[:len(want)]  # constant stencils: the host keeps the first cell's coefficients only
            if exact:
                assert np.array_equal(want, got), (int(b), k, np.argwhere(want != got)[:3])
            else:
                worst = max(worst, float(np.max(np.abs(want - got) / np.maximum(1.0, np.abs(want)))))
    return worst


def eps_cc_of(tree, fn):
    ids = all_ids(tree)
    e = fn(W.cell_centres(tree, ids, ghosts=True))
    eps_cc = np.zeros((tree.highest_id + 1,) + e.shape[1:])
    eps_cc[ids] = e
    return ids, e, eps_cc


@pytest.mark.parametrize("name", ["eps_smooth_corner_nc8", "eps_jump_uniform_nc8", "eps_const2_corner_nc8", "eps_smooth_uniform_nc16"])
def test_permittivity_stencils_built_on_the_device(name):
    tree, fn = {"eps_smooth_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), G3.eps_smooth),
                "eps_jump_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 3), G3.eps_jump),
                "eps_const2_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 3), G3.eps_const2),
                "eps_smooth_uniform_nc16": (lambda: T.uniform_tree(3, 16, 16, 2), G3.eps_smooth)}[name]
    tree = tree()
    ids, e, eps_cc = eps_cc_of(tree, fn)
    entries, data = S.build_stencils(tree, eps_cc=eps_cc)
    mg = M.mg_t(sides_bc=W.bc_table(tree, bc_mixed))
    M.mg_init(tree, mg)
    mg.set_cc(M.I_EPS, ids, e)
    mg.build_stencils_device()
    compare(tree, host_records(tree, entries, data), mg.built_stencils())
    M.mg_destroy(mg)


ROD = dict(rod_r0=(0.4, 0.4, 0.4), rod_r1=(0.6, 0.6, 0.6), rod_radius=0.05)


@pytest.mark.parametrize("method", ["linear", "gss"])
@pytest.mark.parametrize("shape", ["rod", "sphere", "rod_cone_top"])
def test_electrode_stencils_built_on_the_device(shape, method):
    tree = T.build_tree(3, 8, [16] * 3, 3, lambda l, ixs, ctr: np.linalg.norm(ctr - 0.5, axis=1) < 0.45)
    par = dict(ROD)
    if shape == "sphere":
        par = dict(rod_r0=(0.45, 0.55, 0.5), rod_radius=0.2)
    if shape == "rod_cone_top":
        par.update(cone_tip_radius=0.02, cone_length_frac=0.3)
    el = S.electrode(shape, 3, **par)
    opts = S.lsf_opts(S.LSF_DIST_LINEAR if method == "linear" else S.LSF_DIST_GSS)
    entries, data = S.build_stencils(tree, lsf=el, lsf_options=opts)
    assert data is not None and len(data.ids) > 3
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree), lsf_boundary_value=1.0)
    M.mg_init(tree, mg)
    mg.build_stencils_device(el, opts)
    worst = compare(tree, host_records(tree, entries, data), mg.built_stencils(), exact=(method == "linear"))
    assert worst <= 1e-12, worst
    M.mg_destroy(mg)


def test_device_built_problem_solves_like_the_host_built_one():
    """permittivity + electrode in the same tree (mg_box_lpld_lsf_stencil where they meet): host-built and device-built
    set-ups, same FMG + V-cycles, identical potentials and fields"""
    tree = T.build_tree(3, 8, [16] * 3, 3, lambda l, ixs, ctr: np.linalg.norm(ctr - 0.5, axis=1) < 0.45)
    ids, e, eps_cc = eps_cc_of(tree, G3.eps_smooth)
    el = S.electrode("rod", 3, **ROD)
    bc = W.bc_dirichlet_zero(tree)
    leaves, rhs = W.random_rhs_on_leaves(tree)
    res = []
    for device in (False, True):
        mg = M.mg_t(sides_bc=bc, lsf_boundary_value=1.0)
        M.mg_init(tree, mg)
        if device:
            mg.set_cc(M.I_EPS, ids, e)
            mg.build_stencils_device(el)
        else:
            M.mg_set_operators_tree(tree, mg, eps_cc=eps_cc, lsf=el)
        mg.set_cc(M.I_RHS, leaves, rhs)
        hist = []
        M.mg_fas_fmg(tree, mg, True, False)
        for _ in range(3):
            M.mg_fas_vcycle(tree, mg, True)
            hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
        M.mg_compute_phi_gradient(tree, mg, -1.0, True)
        res.append((np.array(hist), mg.get_cc(M.I_PHI, ids), mg.get_fc(ids), mg.get_cc(M.I_FLD, ids)))
        M.mg_destroy(mg)
    assert res[0][0][-1] < 0.1 * res[0][0][0]  # the cycles converge; what matters here is host-built == device-built
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_unresolved_electrode_is_an_error_like_in_the_reference():
    tree = T.uniform_tree(3, 8, 8, 2)
    el = S.electrode("sphere", 3, rod_r0=(0.51, 0.52, 0.53), rod_radius=1e-3)  # falls between the coarse cell centres
    mg = M.mg_t(sides_bc=W.bc_dirichlet_zero(tree))
    M.mg_init(tree, mg)
    with pytest.raises(M.AfmgError, match="not resolved on coarse grid"):
        mg.build_stencils_device(el)
    M.mg_destroy(mg)
