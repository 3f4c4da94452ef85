"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_check.py

Every rank solves the same problems twice -- with the partitioned multi-GPU handle (boxes split by
Morton ranges, halos through NVLink peer memory) and with a private single-GPU handle on its own
device -- and compares its own boxes BIT FOR BIT, plus the residual histories.  Prints one line per
case and exits non-zero on any mismatch.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from afivo_streamer_b200 import mg as M  # noqa: E402
from afivo_streamer_b200 import tree as T  # noqa: E402
from afivo_streamer_b200 import workloads as W  # noqa: E402


def bc_mixed(nb, coords):
    d = (nb - 1) // 2
    vals = 0.3 * np.sin(3.0 * coords.sum(axis=-1)) + 0.1 * nb
    if d == coords.shape[-1] - 1:
        return W.AF_BC_DIRICHLET, vals
    return W.AF_BC_NEUMANN, vals


CASES = {
    "uniform_nc8_l4": (lambda: T.uniform_tree(3, 8, 8, 4), {}),
    "corner_nc8_l4": (lambda: T.corner_refined_tree(3, 8, 8, 4), {}),
    "shell_nc8_l5": (lambda: T.shell_tree(8, 8, 4, 0.35), {}),
    "uniform_nc16_l3": (lambda: T.uniform_tree(3, 16, 16, 3), {}),
    "multibox_coarse_nc8": (lambda: T.build_tree(3, 8, [16, 8, 24], 3,
                                                  lambda l, ix, c: np.linalg.norm(c - 0.4, axis=1) < 0.45), {}),
    "corner_nc8_l4_corners_mean": (lambda: T.corner_refined_tree(3, 8, 8, 4),
                                   dict(use_corners=True, helmholtz_lambda=50.0)),
    "channel_nc8": (lambda: T.channel_tree(8, 8, 6, 3), {}),
    "eps_smooth_corner_nc8": (lambda: T.corner_refined_tree(3, 8, 8, 4), dict(stencils=("eps", "eps_smooth", 0.0))),
    "eps_jump_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 4), dict(stencils=("eps", "eps_jump", 0.0))),
    "lsf_sphere_uniform_nc8": (lambda: T.uniform_tree(3, 8, 8, 4),
                               dict(stencils=("lsf", "lsf_sphere", 1.5), lsf_boundary_value=1.5)),
}


def explicit_stencils(tree, bc, spec):
    """Stencils for variable-eps / level-set cases, built by the oracle (test infrastructure standing in
    for the reference's host-side builders) exactly as tests/test_gpu_stencils.py does."""
    import test_gpu_stencils as TS
    from oracle.oracle import Oracle
    from util import all_ids, stencils_from_oracle
    kind, fn, val = spec
    orc = Oracle(tree, with_eps=(kind == "eps"), lsf_boundary_value=val)
    orc.set_bc(bc)
    ids = all_ids(tree)
    if kind == "eps":
        orc.set_cc(M.I_EPS, ids, getattr(TS, fn)(W.cell_centres(tree, ids, ghosts=True)))
    else:
        lids, dd = TS.lsf_distances(tree, getattr(TS, fn))
        orc.set_lsf_distances(lids, dd)
    orc.mg_init()
    return stencils_from_oracle(tree, orc)


def solve(tree, bc, ids, rhs, comm, local, opts, n_v=3):
    opts = dict(opts)
    spec = opts.pop("stencils", None)
    mg = M.mg_t(sides_bc=bc, device=local, comm=comm, **opts)
    M.mg_init(tree, mg)
    if spec is not None:
        mg.set_stencils(explicit_stencils(tree, bc, spec))
    mg.set_cc(M.I_RHS, ids, rhs)
    hist = []
    M.mg_fas_fmg(tree, mg, True, False)
    hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(n_v):
        M.mg_fas_vcycle(tree, mg, True)
        hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    M.mg_fas_fmg(tree, mg, False, True)
    all_ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    owner = np.array([mg.owner_of_box(i) for i in all_ids])
    out = {v: mg.get_cc(v, all_ids) for v in (M.I_PHI, M.I_TMP)}
    if spec is None:
        # field from potential: gradient, norm and the norm's ghost cells (af_gc_interp reads peers' halos)
        M.field_from_potential(tree, mg, -1.0)
        out["fc"] = mg.get_fc(all_ids)
        out["fld"] = mg.get_cc(M.I_FLD, all_ids)
    s = M.af_tree_sum_cc(tree, mg, M.I_PHI)
    mx = M.af_tree_maxabs_cc(tree, mg, M.I_PHI)
    M.mg_destroy(mg)
    return np.array(hist), out, owner, s, mx


def helmholtz_case(rank, local, comm):
    """photoi_helmh_compute (three modes, shared rhs) on the partitioned handles vs private single-GPU handles."""
    lambdas = np.array([4147.85, 10950.93, 66755.67]) * 0.2 * 0.02
    coeffs = np.array([1117314.935, 28692377.5, 2748842283.0]) * (0.2 * 0.02) ** 2
    tree = T.uniform_tree(3, 8, 8, 4)
    bc = W.bc_table(tree, M.photoi_helmh_bc)
    ids, rhs = W.random_rhs_on_leaves(tree)
    all_ids = np.concatenate(tree.lvl_ids).astype(np.int32)
    res = []
    for c in (None, comm):
        mgs = []
        for lam in lambdas:
            mg = M.mg_t(sides_bc=bc, device=local, comm=c, helmholtz_lambda=lam ** 2, prolongation_type=M.MG_PROLONG_LINEAR)
            M.mg_init(tree, mg)
            mgs.append(mg)
        mgs[0].set_cc(M.I_RHS, ids, rhs * 1.0e3)
        ncyc, r = M.photoi_helmh_compute(tree, mgs, coeffs, 10, 1.0e-2)
        owner = np.array([mgs[0].owner_of_box(i) for i in all_ids])
        res.append((ncyc, r, mgs[0].get_cc(M.I_PHOTO, all_ids), owner))
        for mg in mgs:
            M.mg_destroy(mg)
        dist.barrier()
    (n1, r1, p1, _), (nN, rN, pN, owner) = res
    mine = owner == rank
    ndiff = int(np.count_nonzero(p1[mine] != pN[mine]))
    ok = list(n1) == list(nN) and np.array_equal(r1, rN) and ndiff == 0 and (not mine.any() or np.abs(pN[mine]).max() > 0)
    print(f"[rank {rank}] helmholtz_modes: FMG cycles {list(nN)} residual_equal={np.array_equal(r1, rN)} "
          f"cells_differing={ndiff} {'OK' if ok else 'MISMATCH'}", flush=True)
    return 0 if ok else 1


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = M.comm_from_torch()
    bad = 0
    names = sys.argv[1:] or sorted(CASES)
    for name in names:
        mk, opts = CASES[name]
        tree = mk()
        bc = W.bc_table(tree, bc_mixed)
        ids, rhs = W.random_rhs_on_leaves(tree)
        h1, o1, _, s1, m1 = solve(tree, bc, ids, rhs, None, local, opts)
        dist.barrier()
        hN, oN, owner, sN, mN = solve(tree, bc, ids, rhs, comm, local, opts)
        mine = owner == rank
        ok = np.array_equal(h1, hN) and s1 == sN and m1 == mN
        ndiff = 0
        for v in o1:
            ndiff += int(np.count_nonzero(o1[v][mine] != oN[v][mine]))
        ok = ok and ndiff == 0
        if "fld" in oN and mine.any():
            ok = ok and float(np.abs(oN["fld"][mine]).max()) > 0.0  # the field was really computed
        print(f"[rank {rank}] {name}: boxes {tree.n_boxes} own {int(mine.sum())} residuals {hN[0]:.3e}->{hN[-1]:.3e} "
              f"hist_equal={np.array_equal(h1, hN)} sum_equal={s1 == sN} cells_differing={ndiff} "
              f"{'OK' if ok else 'MISMATCH'}", flush=True)
        bad += 0 if ok else 1
        dist.barrier()
    bad += helmholtz_case(rank, local, comm)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t.item()) != 0:
        raise SystemExit(f"{int(t.item())} mismatching case(s)")
    if rank == 0:
        print("mgpu_check: all cases bit-identical to the single-GPU solve")


if __name__ == "__main__":
    main()
