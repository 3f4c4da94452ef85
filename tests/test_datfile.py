"""CPU: the afivo .dat v3 reader / writer (afivo/src/m_af_output.f90:41-373).  The
reference ships no .dat fixture, so the reader is checked against the writer (which follows af_write_tree
statement by statement), against a byte stream assembled by hand from the write statements, and for its
NDIM detection and error behaviour."""
import struct

import numpy as np
import pytest

from afivo_streamer_b200 import datfile as D
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

from dat_util import assert_same_dat, make_dat
from util import bc_mixed


def oracle_with_data(tree, **kw):
    bc = W.bc_table(tree, bc_mixed)
    orc = Oracle(tree, **kw)
    orc.set_bc(bc)
    orc.mg_init()
    ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    rng = np.random.default_rng(1)
    for var in range(3):
        orc.set_cc(var, ids, rng.uniform(-1, 1, (len(ids), tree.box_len)))
    return orc, bc


@pytest.mark.parametrize("mk", [lambda: T.corner_refined_tree(3, 4, 8, 3), lambda: T.corner_refined_tree(2, 8, 8, 4),
                                lambda: T.build_tree(2, 8, [8, 16], 3, None, coord_t=T.AF_CYL)])
def test_round_trip(tmp_path, mk):
    tree = mk()
    orc, bc = oracle_with_data(tree)
    dat = make_dat(tree, orc, bc)
    dat.other_data = b"\x01\x02\x03\x04 streamer time step data"
    path = str(tmp_path / "sim_000001.dat")
    D.write_tree(path, dat)
    back = D.read_tree(path)  # NDIM detected
    assert back.ndim == tree.ndim
    assert_same_dat(dat, back)
    assert_same_dat(dat, D.read_tree(path, ndim=tree.ndim))
    # derived views
    t2 = back.bc_table("phi")
    order = np.lexsort((bc.nbs, bc.ids))
    order2 = np.lexsort((t2.nbs, t2.ids))
    assert np.array_equal(bc.ids[order], t2.ids[order2]) and np.array_equal(bc.types[order], t2.types[order2])
    assert np.array_equal(bc.vals[order], t2.vals[order2])
    assert back.var_index("rhs") == 2
    with pytest.raises(KeyError):
        back.var_index("nope")


def test_unused_ids_and_variables_not_written(tmp_path):
    tree = T.corner_refined_tree(3, 4, 8, 3).permuted_ids(np.random.default_rng(3))
    orc, bc = oracle_with_data(tree)
    dat = make_dat(tree, orc, bc, removed=(7, 9))
    dat.cc_write_binary[2] = False  # "tmp" has write_binary = F in the streamer (src/m_streamer.f90:304-305)
    del dat.cc[3]
    path = str(tmp_path / "a.dat")
    D.write_tree(path, dat)
    back = D.read_tree(path, ndim=3)
    assert_same_dat(dat, back)
    with pytest.raises(KeyError):
        back.cc_of("tmp", [1])


def test_hand_assembled_header(tmp_path):
    """First bytes of a file written by af_write_tree for a one-box 2D tree, assembled from the write
    statements (m_af_output.f90:54-66): version, ready, box_limit, highest_lvl, highest_id, n_cell,
    n_var_cell, n_var_face, coord_t, coarse_grid_size(2), periodic(2), r_base(2), dr_base(2)."""
    tree = T.uniform_tree(2, 4, 4, 1)
    orc, bc = oracle_with_data(tree)
    dat = make_dat(tree, orc, bc)
    path = str(tmp_path / "b.dat")
    D.write_tree(path, dat)
    raw = open(path, "rb").read()
    want = struct.pack("<9i", 3, 1, dat.box_limit, 1, 1, 4, 3, 1, 1) + struct.pack("<2i", 4, 4) + \
        struct.pack("<2i", 0, 0) + struct.pack("<2d", 0.0, 0.0) + struct.pack("<2d", 0.25, 0.25)
    assert raw[:len(want)] == want
    assert raw[len(want):len(want) + 20] == b"phi".ljust(20)
    # names: 2 x 1024 x 20 characters, then 4 x 1024 four-byte entries (:68-73)
    off = len(want) + 2 * 1024 * 20
    assert struct.unpack("<3i", raw[off:off + 12]) == (1, 1, 1)


def test_wrong_version_and_wrong_ndim(tmp_path):
    tree = T.corner_refined_tree(3, 4, 8, 2)
    orc, bc = oracle_with_data(tree)
    path = str(tmp_path / "c.dat")
    D.write_tree(path, make_dat(tree, orc, bc))
    with pytest.raises(ValueError):
        D.read_tree(path, ndim=2)
    raw = bytearray(open(path, "rb").read())
    raw[0:4] = struct.pack("<i", 2)
    open(path, "wb").write(raw)
    with pytest.raises(ValueError, match="incompatible file versions"):
        D.read_tree(path, ndim=3)


@pytest.mark.parametrize("case", ["eps", "lsf", "eps_lsf_2d_cyl"])
def test_stencils_rebuilt_from_file_data_equal_the_stored_ones(tmp_path, case):
    """A .dat file holds the permittivity variable and the level-set distance stencils next to the operator /
    prolongation stencils built from them: datfile.rebuild_stencil_entries (library builders on the file's data)
    must give back exactly the stored stencils (here the oracle's), so files without stored operators are usable."""
    import test_gpu_2d as G2
    import test_gpu_stencils as G3
    from oracle.oracle import Oracle
    from util import bc_mixed
    if case == "eps":
        tree, eps, lsf, dist, bc_fn = T.corner_refined_tree(3, 8, 8, 3), G3.eps_smooth, None, None, bc_mixed
    elif case == "lsf":
        tree, eps, lsf, dist, bc_fn = T.corner_refined_tree(3, 8, 8, 3), None, G3.lsf_sphere, G3.lsf_distances, bc_mixed
    else:
        tree = T.build_tree(2, 8, [8, 8], 4, None, coord_t=T.AF_CYL)
        eps, lsf, dist, bc_fn = G2.eps2, (lambda r: np.linalg.norm(r - np.array([0.0, 0.5]), axis=-1) - 0.2), \
            G2.lsf_distances2, G2.bc_cyl
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
    path = str(tmp_path / "t.dat")
    D.write_tree(path, make_dat(tree, orc, bc, extra_cc=extra, lsf_dd=lsf_dd))
    dat = D.read_tree(path)
    stored = dat.stencil_entries()
    rebuilt = D.rebuild_stencil_entries(dat, "eps" if eps is not None else None)
    assert stored and [e["box_id"] for e in stored] == [e["box_id"] for e in rebuilt]
    nd = tree.ndim
    default_p = np.array([9, 3, 3, 1]) / 16.0 if nd == 2 else np.array([27, 9, 9, 3, 9, 3, 3, 1]) / 64.0
    for s, r in zip(stored, rebuilt):
        assert s["tag"] == r["tag"] and s["op"][0] == r["op"][0] and bool(s.get("cyl")) == bool(r["cyl"])
        assert np.array_equal(np.asarray(s["op"][1]).reshape(-1), np.asarray(r["op"][1]).reshape(-1)), s["box_id"]
        assert (s.get("f") is None) == (r["f"] is None)
        if r["f"] is not None:
            assert np.array_equal(s["f"], r["f"])
        if "prolong" in s:
            rp = r.get("prolong") or (1, 3, default_p)
            assert (s["prolong"][0], s["prolong"][1]) == (rp[0], rp[1])
            assert np.array_equal(np.asarray(s["prolong"][2]).reshape(-1), np.asarray(rp[2]).reshape(-1))
