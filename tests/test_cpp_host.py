"""CPU: the C++ host-side mirror (include/afmg.hpp) compiles with -pedantic -Werror and builds the same tree,
array by array, as the Python builder that follows afivo's conventions (ids, ix, parent, children in af_child_dix
order, neighbors, neighbor_mat, level lists)."""
import os
import subprocess

import numpy as np
import pytest

from afivo_streamer_b200 import tree as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


CASES = {
    "uniform_8_8_3": ("uniform", 8, [8, 8, 8], 3, lambda: T.uniform_tree(3, 8, 8, 3)),
    "uniform_16_16_2": ("uniform", 16, [16, 16, 16], 2, lambda: T.uniform_tree(3, 16, 16, 2)),
    "corner_8_8_4": ("corner", 8, [8, 8, 8], 4, lambda: T.corner_refined_tree(3, 8, 8, 4)),
    "sphere_multibox_8": ("sphere", 8, [16, 8, 24], 3,
                          lambda: T.build_tree(3, 8, [16, 8, 24], 3, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45)),
    "sphere_4_deep": ("sphere", 4, [8, 8, 8], 4,
                      lambda: T.build_tree(3, 4, [8, 8, 8], 4, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45)),
}


@pytest.fixture(scope="module")
def dump_exe(tmp_path_factory):
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    exe = str(tmp_path_factory.mktemp("cpp") / "cpp_tree_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_tree_dump.cpp"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    return exe


@pytest.mark.parametrize("name", sorted(CASES))
def test_cpp_tree_matches_python_builder(dump_exe, name):
    kind, nc, cgs, lvl, mk = CASES[name]
    out = subprocess.run([dump_exe, kind, str(nc)] + [str(c) for c in cgs] + [str(lvl)], capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    t = mk()
    hl, hid = map(int, lines[0].split())
    assert (hl, hid) == (t.highest_lvl, t.highest_id)
    for l in range(hl):
        vals = list(map(int, lines[1 + l].split()[1:]))
        assert vals[0] == len(t.lvl_ids[l]) and vals[1:] == list(map(int, t.lvl_ids[l]))
    boxes = np.array([list(map(int, ln.split()[1:])) for ln in lines[1 + hl:1 + hl + hid]])
    ids = np.arange(1, hid + 1)
    assert np.array_equal(boxes[:, 0], t.lvl[ids])
    assert np.array_equal(boxes[:, 1:4], t.ix[ids])
    assert np.array_equal(boxes[:, 4], t.parent[ids])
    assert np.array_equal(boxes[:, 5:13], t.children[ids])
    assert np.array_equal(boxes[:, 13:19], t.neighbors[ids])
    assert np.array_equal(boxes[:, 19:46], t.neighbor_mat[ids])
    geo = np.array([list(map(float, ln.split()[1:])) for ln in lines[1 + hl + hid:1 + hl + 2 * hid]])
    assert np.array_equal(geo[:, 0:3], t.r_min[ids]) and np.array_equal(geo[:, 3:6], t.dr[ids])
    from afivo_streamer_b200 import workloads as W
    fc = W.face_coords(t, ids, 3)  # af_get_face_coords on the low-y side
    assert np.array_equal(geo[:, 6:9], fc[:, 0, :]) and np.array_equal(geo[:, 9:12], fc[:, -1, :])
    assert lines[-1] in ("device present", "error -2")  # AFMG_ERR_CUDA without a GPU: no CPU fallback


@pytest.fixture(scope="module")
def stencil_exe(tmp_path_factory):
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    exe = str(tmp_path_factory.mktemp("cpp") / "cpp_stencil_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_stencil_dump.cpp"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    return exe


@pytest.mark.parametrize("custom,method", [(0, 0), (1, 0), (0, 1)])
def test_cpp_stencil_builder_matches_python_mirror(stencil_exe, tmp_path, custom, method):
    """afmg::mg_build_stencils (C++ mirror) == stencils.build_stencils (Python mirror): same tags, kinds and
    coefficients for a refined tree with variable permittivity and a spherical electrode; both only walk the tree, the
    arithmetic is the library's (afmg_build_box_*)."""
    from afivo_streamer_b200 import stencils as S
    from afivo_streamer_b200 import workloads as W
    out_bin = str(tmp_path / "st.bin")
    out = subprocess.run([stencil_exe, out_bin, str(custom), str(method)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    n_desc, n_blob, n_lsf = map(int, lines[0].split())
    desc = [list(map(int, ln.split()[1:])) for ln in lines[1:1 + n_desc]]
    lsf_ids = [int(ln.split()[1]) for ln in lines[1 + n_desc:1 + n_desc + n_lsf]]
    raw = np.fromfile(out_bin)
    blob, rest = raw[:n_blob], raw[n_blob:]

    t = T.corner_refined_tree(3, 8, 8, 4)
    ids = np.concatenate(t.lvl_ids).astype(np.int32)
    r = W.cell_centres(t, ids, ghosts=True)
    eps_cc = np.ones((t.highest_id + 1,) + r.shape[1:-1])
    eps_cc[ids] = 1.0 + 0.5 * (r[..., 0] * r[..., 1]) + 0.25 * r[..., 2]

    def lsf(p):
        dx, dy, dz = p[0] - 0.2, p[1] - 0.25, p[2] - 0.15
        return np.sqrt(dx * dx + dy * dy + dz * dz) - 0.11

    entries, data = S.build_stencils(t, eps_cc=eps_cc, lsf=lsf, lsf_options=S.lsf_opts(method),
                                     lsf_use_custom_prolongation=bool(custom))
    assert n_desc == len(entries) and n_lsf == len(data.ids) and lsf_ids == list(map(int, data.ids))
    assert any(e["tag"] & 1 for e in entries) and any(e["tag"] & 2 for e in entries)
    ncell = 8 ** 3
    for d, e in zip(desc, entries):
        box_id, tag, op_stype, cyl, pst, psh, op_off, f_off, p_off = d
        assert (box_id, tag, op_stype, bool(cyl)) == (e["box_id"], e["tag"], e["op"][0], e["cyl"])
        n_op = 7 if op_stype == 1 else 7 * ncell
        assert np.array_equal(blob[op_off:op_off + n_op], np.asarray(e["op"][1]).reshape(-1)), box_id
        assert (f_off >= 0) == (e["f"] is not None)
        if f_off >= 0:
            assert np.array_equal(blob[f_off:f_off + ncell], e["f"])
        if e.get("prolong") is None:
            assert psh == 0
        else:
            assert (pst, psh) == e["prolong"][:2]
            n_p = 4 if pst == 1 else 4 * ncell
            assert np.array_equal(blob[p_off:p_off + n_p], np.asarray(e["prolong"][2]).reshape(-1)), box_id
    per = 6 * ncell + ncell
    for n in range(n_lsf):
        assert np.array_equal(rest[n * per:n * per + 6 * ncell], data.dd[n].reshape(-1))
        assert np.array_equal(rest[n * per + 6 * ncell:(n + 1) * per], data.lsf_cells[n])


CASES2D = {
    "uniform2d_8_4": ("uniform", 8, [8, 8], 4, 0, lambda: T.uniform_tree(2, 8, 8, 4)),
    "corner2d_8_5": ("corner", 8, [8, 8], 5, 0, lambda: T.corner_refined_tree(2, 8, 8, 5)),
    "sphere2d_multibox_cyl": ("sphere", 8, [16, 24], 4, 1,
                              lambda: T.build_tree(2, 8, [16, 24], 4, lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45,
                                                   coord_t=T.AF_CYL)),
}


@pytest.mark.parametrize("name", sorted(CASES2D))
def test_cpp_2d_tree_matches_python_builder(dump_exe, name):
    """afmg::af_build_tree_nd(2, ...): the 2D / cylindrical trees of config C1 from the C++ mirror."""
    kind, nc, cgs, lvl, cyl, mk = CASES2D[name]
    out = subprocess.run([dump_exe, kind, str(nc), str(cgs[0]), str(cgs[1]), "1", str(lvl), "2", str(cyl)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    t = mk()
    hl, hid = map(int, lines[0].split())
    assert (hl, hid) == (t.highest_lvl, t.highest_id)
    for l in range(hl):
        vals = list(map(int, lines[1 + l].split()[1:]))
        assert vals[0] == len(t.lvl_ids[l]) and vals[1:] == list(map(int, t.lvl_ids[l]))
    boxes = np.array([list(map(int, ln.split()[1:])) for ln in lines[1 + hl:1 + hl + hid]])
    ids = np.arange(1, hid + 1)
    assert np.array_equal(boxes[:, 0], t.lvl[ids])
    assert np.array_equal(boxes[:, 1:3], t.ix[ids])
    assert np.array_equal(boxes[:, 3], t.parent[ids])
    assert np.array_equal(boxes[:, 4:8], t.children[ids])
    assert np.array_equal(boxes[:, 8:12], t.neighbors[ids])
    assert np.array_equal(boxes[:, 12:21], t.neighbor_mat[ids])
    geo = np.array([list(map(float, ln.split()[1:])) for ln in lines[1 + hl + hid:1 + hl + 2 * hid]])
    assert np.array_equal(geo[:, 0:2], t.r_min[ids]) and np.array_equal(geo[:, 2:4], t.dr[ids])
    assert lines[-1] == f"coord {t.coord_t}"


@pytest.mark.parametrize("cyl", [0, 1])
def test_native_electrode_example_sets_up_the_same_problem_as_the_python_mirror(cyl):
    """tools/electrode_example.cpp --dry-run (afivo/examples/electrode_example.f90 in 2D on the C++ mirror: 2D tree
    builder, built-in rod level set, mg_build_stencils) against the Python mirror's set-up of the same problem,
    which tests/test_gpu_electrode_example.py solves on the GPU."""
    from afivo_streamer_b200 import stencils as S
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools"), "-s", "electrode_example_2d"])
    out = subprocess.run([os.path.join(ROOT, "tools", "electrode_example_2d")] + (["cyl"] if cyl else []) + ["--dry-run"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    hl, hid, n_desc, n_lsf, n_blob, blob_sum = out.stdout.split()
    nc = 8
    t = T.build_tree(2, nc, [4 * nc] * 2, 5, lambda l, ixs, ctr: (l < 5) & ((ixs[:, 0] - 1) * (0.25 / 2 ** (l - 1)) < 0.5),
                     coord_t=T.AF_CYL if cyl else T.AF_XYZ)
    el = S.electrode("rod", 2, rod_r0=(0.4, 0.4), rod_r1=(0.6, 0.6), rod_radius=0.02)
    entries, data = S.build_stencils(t, lsf=el)
    assert (int(hl), int(hid)) == (t.highest_lvl, t.highest_id)
    assert int(n_desc) == len(entries) and int(n_lsf) == len(data.ids)
    blob = np.concatenate([np.concatenate([np.asarray(e["op"][1]).reshape(-1), e["f"]]) for e in entries])
    assert int(n_blob) == blob.size
    assert abs(float(blob_sum) - blob.sum()) <= 1e-12 * np.abs(blob).sum()


@pytest.fixture(scope="module")
def dat_exe(tmp_path_factory):
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    exe = str(tmp_path_factory.mktemp("cpp") / "cpp_dat_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_dat_dump.cpp"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    return exe


@pytest.mark.parametrize("case", ["plain3d", "eps_lsf_3d", "lsf_cyl_2d"])
def test_cpp_dat_reader_matches_python_reader(dat_exe, tmp_path, case):
    """include/afmg_dat.hpp (af_read_tree for compiled hosts) against afivo_streamer_b200/datfile.py on files written
    by the Python writer: header, variables, topology, stored boundary conditions, and the stencil set handed to
    afmg_set_stencils (descriptors and blob offsets equal, blob checksum equal)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_2d as G2
    import test_gpu_stencils as G3
    from afivo_streamer_b200 import datfile as D
    from afivo_streamer_b200 import workloads as W
    from dat_util import make_dat
    from oracle.oracle import Oracle
    from util import bc_mixed
    if case == "plain3d":
        tree, eps, lsf, dist, bc_fn = T.corner_refined_tree(3, 8, 8, 3), None, None, None, bc_mixed
    elif case == "eps_lsf_3d":
        tree, eps, lsf, dist, bc_fn = T.corner_refined_tree(3, 8, 8, 3), G3.eps_smooth, G3.lsf_sphere, G3.lsf_distances, bc_mixed
    else:
        tree = T.build_tree(2, 8, [8, 8], 4, lambda l, ix, c: np.linalg.norm(c - 0.5, axis=1) < 0.4, coord_t=T.AF_CYL)
        eps, lsf, dist, bc_fn = None, G2.lsf_circle, G2.lsf_distances2, G2.bc_cyl
    ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    bc = W.bc_table(tree, bc_fn)
    orc = Oracle(tree, with_eps=eps is not None)
    orc.set_bc(bc)
    extra, lsf_dd = {}, None
    if eps is not None:
        e = np.ones((tree.highest_id + 1, tree.box_len))
        e[ids] = eps(W.cell_centres(tree, ids, ghosts=True)).reshape(len(ids), -1)
        orc.set_cc(3, ids, e[ids])
        extra["eps"] = e
    if lsf is not None:
        lsf_dd = dist(tree, lsf)
        orc.set_lsf_distances(*lsf_dd)
    orc.mg_init()
    rng = np.random.default_rng(5)
    orc.set_cc(0, ids, rng.uniform(-1, 1, (len(ids), tree.box_len)))
    orc.set_cc(1, ids, rng.uniform(-1, 1, (len(ids), tree.box_len)))
    path = str(tmp_path / "t.dat")
    D.write_tree(path, make_dat(tree, orc, bc, extra_cc=extra, lsf_dd=lsf_dd))
    dat = D.read_tree(path)
    out = subprocess.run([dat_exe, path], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [ln.split() for ln in out.stdout.strip().splitlines()]
    t = dat.tree
    assert list(map(int, rows[0][1:])) == [t.ndim, t.highest_lvl, t.highest_id, t.nc, t.coord_t, int(dat.ready)]
    assert float(rows[1][7]) == t.dr_base[0] and list(map(int, rows[1][1:1 + t.ndim])) == list(map(int, t.coarse_grid_size))
    vrows = [r for r in rows if r[0] == "V"]
    assert [r[1] for r in vrows] == dat.cc_names
    for r, iv in zip(vrows, range(1, len(dat.cc_names) + 1)):
        assert int(r[2]) == int(iv in dat.cc)
        if iv in dat.cc:
            assert abs(float(r[3]) - dat.cc[iv].sum()) <= 1e-11 * np.abs(dat.cc[iv]).sum()
    n = t.highest_id
    idr = np.arange(1, n + 1)
    topo = int((t.lvl[idr].astype(np.int64) * 3 + t.parent[idr] * 5 + dat.tag[idr] * 7).sum()
               + (t.ix[idr].astype(np.int64) * (11 + np.arange(t.ndim))).sum()
               + (t.children[idr].astype(np.int64) * (1 + np.arange(1 << t.ndim))).sum()
               + (t.neighbors[idr].astype(np.int64) * (2 + np.arange(2 * t.ndim))).sum())
    assert int(next(r for r in rows if r[0] == "T")[1]) == topo
    brow = next(r for r in rows if r[0] == "B")
    assert int(brow[1]) == len(dat.bc) and int(brow[2]) == sum(int(b.bc_type.sum()) for b in dat.bc.values())
    assert abs(float(brow[3]) - sum(b.bc_val.sum() for b in dat.bc.values())) < 1e-9
    entries = dat.stencil_entries()
    srow = next(r for r in rows if r[0] == "S")
    drows = [list(map(int, r[1:])) for r in rows if r[0] == "D"]
    assert int(srow[1]) == len(entries) == len(drows)
    off, total = 0, 0.0
    for e, dr in zip(entries, drows):
        want = [e["box_id"], e["tag"], 0, int(bool(e.get("cyl", False))), 0, 0, 0, -1, 0]
        if e.get("op") is not None:
            want[2], want[6] = int(e["op"][0]), off
            off += np.asarray(e["op"][1]).size
            total += float(np.asarray(e["op"][1]).sum())
        if e.get("f") is not None:
            want[7] = off
            off += e["f"].size
            total += float(e["f"].sum())
        if e.get("prolong") is not None:
            want[4], want[5], want[8] = int(e["prolong"][0]), int(e["prolong"][1]), off
            off += np.asarray(e["prolong"][2]).size
            total += float(np.asarray(e["prolong"][2]).sum())
        assert dr == want, (dr, want)
    assert int(srow[2]) == off
    assert abs(float(srow[3]) - total) <= 1e-9 * max(1.0, abs(total))


def test_cpp_caller_side_helpers(tmp_path):
    """photoi_helmh_parameters and field_residual_threshold of the C++ mirror give the reference's numbers."""
    lib_dir = os.path.join(ROOT, "afivo_streamer_b200")
    exe = str(tmp_path / "cpp_params_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp_params_check.cpp"), "-o", exe, "-L", lib_dir, "-lafmg",
                           "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "params ok" in out.stdout, out.stdout + out.stderr
