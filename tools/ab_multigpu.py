"""A/B timing of library settings on a multi-GPU solve without paying the set-up of bench.py once per setting:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ab_multigpu.py \
        [--workload S3] [--steps 20] "AFMG_SURFACE_FIRST=0" "AFMG_SURFACE_FIRST=1" ...

Every positional argument is a comma-separated list of NAME=VALUE environment settings the library reads at
afmg_create.  For each one: new handle on the same tree, same right-hand side (bench.py's hash), 1 FMG + warm-up
V-cycles, then `steps` V-cycles timed on the device (max over ranks), and the checksum of phi.  Prints one JSON line per
setting on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="S3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("settings", nargs="+")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench as B
    from afivo_streamer_b200 import mg as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = M.comm_from_torch()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tree, bc, _, _, desc = B.build_workload(args.workload, want_rhs=False)
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    box_len = tree.box_len
    for setting in args.settings:
        env = dict(kv.split("=", 1) for kv in setting.split(",") if kv)
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        mg = M.mg_t(sides_bc=bc, device=local, comm=comm, lsf_boundary_value=1.0)
        M.mg_init(tree, mg)
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
        sel = np.nonzero(mg.owners(leaves) == rank)[0] if world > 1 else np.arange(len(leaves))
        ids = np.ascontiguousarray(leaves[sel])
        h_rhs = torch.empty(len(ids) * box_len, dtype=torch.float64).pin_memory()
        chunk = 4096
        for q0 in range(0, len(ids), chunk):
            q1 = min(len(ids), q0 + chunk)
            d = B.synthetic_rhs_device(torch, sel[q0:q1], box_len)
            h_rhs[q0 * box_len:q1 * box_len].copy_(d)
            del d
        torch.cuda.synchronize()
        barrier()
        mg.upload_ptr(M.I_RHS, ids, h_rhs.data_ptr())
        del h_rhs
        barrier()
        M.mg_fas_fmg(tree, mg, True, False)
        barrier()
        mg.fas_vcycle_async(True, 0, args.warmup)
        mg.sync()
        barrier()
        l0 = mg.kernel_launches()
        mg.fas_vcycle_async(True, 0, args.steps)
        mg.sync()
        ms = allmax(mg.last_cycle_ms()) / args.steps
        launches = mg.kernel_launches() - l0
        res = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        csum, cxor = mg.checksum(M.I_PHI)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (csum, cxor))
            csum = sum(a for a, _ in parts) & ((1 << 64) - 1)
        n_fmg = max(2, args.steps // 4)
        barrier()
        mg.fas_fmg_async(False, True, n_fmg)
        mg.sync()
        fmg_ms = allmax(mg.last_cycle_ms()) / n_fmg
        if rank == 0:
            print(json.dumps({"setting": setting, "workload": args.workload, "n_gpus": world, "ms_per_vcycle": ms,
                              "fmg_ms": fmg_ms, "launches_per_cycle_rank0": launches / args.steps, "residual": res,
                              "phi_checksum": f"{csum:016x}"}), flush=True)
        barrier()
        M.mg_destroy(mg)
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
