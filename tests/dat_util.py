"""Build a DatFile (afivo .dat v3 contents) from a synthetic tree plus the oracle's state, the way a real
simulation would have written it: variables phi / rhs / tmp (/ eps / lsf), the boundary conditions stored in
the boxes, operator (key 1) and prolongation (key 2) stencils of every box, level-set distance stencils."""
import numpy as np

from afivo_streamer_b200 import datfile as D
from afivo_streamer_b200 import workloads as W


def make_dat(tree, orc, bc, *, names=("phi", "rhs", "tmp"), extra_cc=None, lsf_dd=None, seed=0, removed=()):
    """names[k] is stored from oracle variable k; extra_cc: dict name -> (n+1, box_len) array."""
    ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    n, nd, nc = tree.highest_id, tree.ndim, tree.nc
    cc_names = list(names) + sorted(extra_cc or {})
    cc = {}
    for k, _ in enumerate(names):
        a = np.zeros((n + 1, tree.box_len))
        a[ids] = orc.get_cc(k, ids)
        cc[k + 1] = a
    for k, nm in enumerate(sorted(extra_cc or {})):
        cc[len(names) + k + 1] = np.asarray(extra_cc[nm], float).reshape(n + 1, tree.box_len)
    in_use = np.zeros(n + 1, bool)
    in_use[ids] = True
    tag = np.zeros(n + 1, np.int32)
    nface = nc ** (nd - 1)
    bcs, stencils = {}, {}
    rows = {}
    for q in range(len(bc.ids)):
        rows.setdefault(int(bc.ids[q]), []).append(q)
    for bid in ids:
        bid = int(bid)
        tag[bid] = orc.tag(bid)
        r = sorted(rows.get(bid, []), key=lambda q: bc.nbs[q])
        if r:
            n_bc = len(r)
            n2i = np.zeros(2 * nd, np.int32)
            bt = np.zeros((n_bc, len(cc_names)), np.int32)
            bv = np.zeros((n_bc, len(cc_names), nface))
            bco = np.zeros((n_bc, nface, nd))
            for m, q in enumerate(r):
                n2i[bc.nbs[q] - 1] = m + 1
                bt[m, 0] = bc.types[q]
                bv[m, 0] = bc.vals[q]
                bco[m] = W.face_coords(tree, np.array([bid]), int(bc.nbs[q]))[0]
            bcs[bid] = D.DatBC(np.array([bc.nbs[q] for q in r], np.int32), n2i, bt, bv, bco)
        lst = []
        stype, coeff, f, cyl = orc.op_stencil(bid)
        st = D.DatStencil(key=1, shape=D.STENCIL_357, stype=stype, cylindrical_gradient=cyl)
        if stype == D.STENCIL_CONSTANT:
            st.c = np.array(coeff, float)
        else:
            st.v = np.array(coeff, float)
        if f is not None:
            st.f = np.array(f, float)
            st.bc_correction = st.f * orc.opts["lsf_boundary_value"]
        lst.append(st)
        if tree.lvl[bid] > 1:
            pst, pshape, pco = orc.prolong_stencil(bid)
            ps = D.DatStencil(key=2, shape=pshape, stype=pst)
            if pst == D.STENCIL_CONSTANT:
                ps.c = np.array(pco, float)
            else:
                ps.v = np.array(pco, float)
            lst.append(ps)
        stencils[bid] = lst
    if lsf_dd is not None:
        lids, dd = lsf_dd
        for bid, d in zip(lids, np.asarray(dd).reshape(len(lids), nc ** nd, 2 * nd)):
            sel = np.nonzero((d < 1.0).any(axis=1))[0]
            if len(sel) == 0:
                continue
            ixs = np.empty((len(sel), nd), np.int32)
            rr = sel.copy()
            for k in range(nd):
                ixs[:, k] = rr % nc + 1
                rr //= nc
            stencils[int(bid)].append(D.DatStencil(key=D.MG_LSF_DISTANCE_KEY, shape=D.STENCIL_246, stype=D.STENCIL_SPARSE,
                                                    sparse_ix=ixs, sparse_v=d[sel]))
    leaves = [tree.leaves(l).astype(np.int32) for l in range(1, tree.highest_lvl + 1)]
    parents = [tree.parents(l).astype(np.int32) for l in range(1, tree.highest_lvl + 1)]
    return D.DatFile(ndim=nd, tree=tree, ready=True, box_limit=max(1000, 2 * n), cc_names=cc_names, fc_names=["field"],
                     cc_num_copies=np.ones(len(cc_names), np.int32), cc_write_output=np.ones(len(cc_names), bool),
                     cc_write_binary=np.ones(len(cc_names), bool), fc_write_binary=np.zeros(1, bool),
                     removed_ids=np.asarray(removed, np.int32), lvl_leaves=leaves, lvl_parents=parents, in_use=in_use,
                     tag=tag, cc=cc, fc={}, bc=bcs, stencils=stencils, other_data=None)


def assert_same_dat(a, b):
    ta, tb = a.tree, b.tree
    assert a.ndim == b.ndim and a.ready == b.ready and a.box_limit == b.box_limit
    for f in ("nc", "coord_t", "highest_lvl", "highest_id"):
        assert getattr(ta, f) == getattr(tb, f), f
    for f in ("coarse_grid_size", "periodic", "r_base", "dr_base", "lvl", "ix", "parent", "children", "neighbors",
              "neighbor_mat", "r_min", "dr"):
        assert np.array_equal(np.asarray(getattr(ta, f)), np.asarray(getattr(tb, f))), f
    for x, y in zip(ta.lvl_ids, tb.lvl_ids):
        assert np.array_equal(x, y)
    assert a.cc_names == b.cc_names and a.fc_names == b.fc_names
    for f in ("cc_num_copies", "cc_write_output", "cc_write_binary", "fc_write_binary", "removed_ids", "in_use", "tag"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for la, lb in ((a.lvl_leaves, b.lvl_leaves), (a.lvl_parents, b.lvl_parents)):
        assert all(np.array_equal(x, y) for x, y in zip(la, lb))
    assert sorted(a.cc) == sorted(b.cc) and all(np.array_equal(a.cc[k], b.cc[k]) for k in a.cc)
    assert sorted(a.fc) == sorted(b.fc) and all(np.array_equal(a.fc[k], b.fc[k]) for k in a.fc)
    assert sorted(a.bc) == sorted(b.bc)
    for k in a.bc:
        for f in ("bc_index_to_nb", "nb_to_bc_index", "bc_type", "bc_val", "bc_coords"):
            assert np.array_equal(getattr(a.bc[k], f), getattr(b.bc[k], f)), (k, f)
    assert sorted(a.stencils) == sorted(b.stencils)
    for k in a.stencils:
        assert len(a.stencils[k]) == len(b.stencils[k])
        for x, y in zip(a.stencils[k], b.stencils[k]):
            assert (x.key, x.shape, x.stype, x.cylindrical_gradient) == (y.key, y.shape, y.stype, y.cylindrical_gradient)
            for f in ("c", "v", "f", "bc_correction", "sparse_ix", "sparse_v"):
                u, v = getattr(x, f), getattr(y, f)
                assert (u is None) == (v is None), (k, f)
                assert u is None or np.array_equal(u, v), (k, f)
    assert a.other_data == b.other_data
