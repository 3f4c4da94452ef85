"""A tree that does not fit one GPU, solved by ONE process on N GPUs (afmg_opts.n_gpus, slabs trimmed to the owned boxes):

    python tools/big_tree_demo.py [--gpus 8] [--levels 7] [--vcycles 6]

Default: the 2048^3-equivalent shell-refined octree (1024^3 uniform + level 8 on all but the outermost box layer, 16^3
boxes): 2 027 593 boxes, 8.3e9 cells, 378 GB of cell data (phi, rhs, tmp, field norm incl. ghost cells) -- twice one
B200's memory -- of which every GPU maps its own eighth.  Same boundary conditions and right-hand-side hash as
bench.py's S3.  Prints one JSON line: per-GPU memory, residual history of 1 FMG + the V-cycles, ms per V-cycle (device
time of the slowest GPU) and cell-updates/s.  The right-hand side goes up in chunks (no 57 GB host buffer)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--levels", type=int, default=7, help="uniform levels below the shell level (6 = bench.py's S3)")
    ap.add_argument("--vcycles", type=int, default=6)
    ap.add_argument("--timed", type=int, default=10)
    args = ap.parse_args()
    import torch

    import bench as B
    from afivo_streamer_b200 import mg as M
    from afivo_streamer_b200 import tree as T
    from afivo_streamer_b200 import workloads as W

    t0 = time.perf_counter()
    tree = T.shell_tree(16, 16, args.levels)
    bc = W.bc_field_homogeneous(tree, 1.0)
    t_tree = time.perf_counter() - t0
    mg = M.mg_t(sides_bc=bc, device=0, n_gpus=args.gpus)
    t0 = time.perf_counter()
    M.mg_init(tree, mg)
    t_init = time.perf_counter() - t0
    mapped, full = mg.slab_bytes()
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    ncell, box_len = tree.nc ** 3, tree.box_len
    chunk = 32768
    pinned = torch.empty(chunk * ncell, dtype=torch.float64).pin_memory()
    torch.cuda.set_device(0)
    t0 = time.perf_counter()
    for q0 in range(0, len(leaves), chunk):
        q1 = min(len(leaves), q0 + chunk)
        full_rec = B.synthetic_rhs_device(torch, np.arange(q0, q1), box_len).view(q1 - q0, 18, 18, 18)
        pinned[:(q1 - q0) * ncell].view(q1 - q0, 16, 16, 16).copy_(full_rec[:, 1:-1, 1:-1, 1:-1])
        torch.cuda.synchronize()
        mg.upload_interior_ptr(M.I_RHS, leaves[q0:q1], pinned.data_ptr())
        del full_rec
    t_rhs = time.perf_counter() - t0
    hist = []
    t0 = time.perf_counter()
    M.mg_fas_fmg(tree, mg, True, False)
    t_fmg = time.perf_counter() - t0
    hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    for _ in range(args.vcycles):
        M.mg_fas_vcycle(tree, mg, True)
        hist.append(M.af_tree_maxabs_cc(tree, mg, M.I_TMP))
    mg.fas_vcycle_async(True, 0, args.timed)
    mg.sync()
    ms = mg.last_cycle_ms() / args.timed
    cu = mg.cell_updates(0, False)
    csum = mg.checksum(M.I_PHI)
    out = {"what": "one process, afmg_opts.n_gpus GPUs, slabs trimmed to the owned boxes (tools/big_tree_demo.py)",
           "n_gpus": args.gpus, "n_boxes": int(tree.n_boxes), "levels": int(tree.highest_lvl),
           "cells_all_levels": int(tree.n_boxes) * ncell,
           "slab_GB_per_gpu_mapped": (mapped / 1e9).round(2).tolist(), "slot_space_GB": float(full[0] / 1e9),
           "fits_one_gpu": bool(full[0] < 170e9),
           "ms_per_vcycle": ms, "cell_updates_per_s": cu / (ms * 1e-3), "fmg_from_scratch_s": t_fmg,
           "residual_history": hist, "phi_checksum": f"{csum[0]:016x}",
           "host_seconds": {"tree": t_tree, "mg_init_all_gpus": t_init, "rhs_generate_and_upload": t_rhs}}
    print(json.dumps(out))
    M.mg_destroy(mg)


if __name__ == "__main__":
    main()
