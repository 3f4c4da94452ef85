"""Per-kernel times of field_from_potential (mg_compute_phi_gradient + field norm + af_gc_tree of the norm; SURVEY 8f
rank 2) on a bench.py workload, from the library's profiling mode (CUDA events around every launch):

    python tools/profile_field.py [--workload S3] [--reps 3]

Prints one JSON line: ms per call of every kernel group, the blocking call's wall time and the algorithmic bytes
(read phi incl. face ghost cells, write fc and the norm) over both."""
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
    ap.add_argument("--workload", default="S3")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch

    import bench as B
    from afivo_streamer_b200 import mg as M

    torch.cuda.set_device(0)
    tree, bc, _, _, desc = B.build_workload(args.workload, want_rhs=False)
    mg = M.mg_t(sides_bc=bc, device=0, lsf_boundary_value=1.0)
    M.mg_init(tree, mg)
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    box_len, chunk = tree.box_len, 4096
    h_rhs = torch.empty(len(leaves) * box_len, dtype=torch.float64).pin_memory()
    for q0 in range(0, len(leaves), chunk):
        q1 = min(len(leaves), q0 + chunk)
        h_rhs[q0 * box_len:q1 * box_len].copy_(B.synthetic_rhs_device(torch, np.arange(q0, q1), box_len))
    torch.cuda.synchronize()
    mg.upload_ptr(M.I_RHS, leaves, h_rhs.data_ptr())
    del h_rhs
    M.mg_fas_fmg(tree, mg, True, False)
    M.field_from_potential(tree, mg, -1.0)  # allocates fc / norm
    t0 = time.perf_counter()
    for _ in range(args.reps):
        M.field_from_potential(tree, mg, -1.0)
    wall_ms = 1e3 * (time.perf_counter() - t0) / args.reps
    mg.set_profiling(True)
    for _ in range(args.reps):
        M.field_from_potential(tree, mg, -1.0)
    prof = mg.profile()
    mg.set_profiling(False)
    nc = tree.nc
    per_cell = 8 * ((nc + 2) / nc) ** 3 + 3 * 8 * (nc + 1) / nc + 8
    nbytes = per_cell * tree.n_boxes * nc ** 3
    groups = {}
    for k, (ms, calls) in prof.items():
        g = k.split("_L")[0]
        groups[g] = groups.get(g, 0.0) + ms / args.reps
    sample = leaves[::max(1, len(leaves) // 64)][:64]
    with np.errstate(over="ignore"):
        csum = [int(np.add.reduce(np.ascontiguousarray(a).view(np.uint64).reshape(-1), dtype=np.uint64))
                for a in (mg.get_cc(M.I_FLD, sample), mg.get_fc(sample))]
    print(json.dumps({"workload": args.workload, "n_boxes": int(tree.n_boxes), "blocking_call_ms": wall_ms,
                      "kernel_ms_per_call": {k: round(v, 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1])},
                      "algorithmic_GB": nbytes / 1e9, "algorithmic_GBs_call": nbytes / wall_ms / 1e6,
                      "algorithmic_GBs_grad_kernel": nbytes / max(groups.get("field_grad", 1e-9), 1e-9) / 1e6,
                      "sample_checksums": {"norm": f"{csum[0]:016x}", "fc": f"{csum[1]:016x}",
                                           "what": "wrapping sum of the bit patterns over 64 leaves spread through the tree"}}))
    M.mg_destroy(mg)


if __name__ == "__main__":
    main()
