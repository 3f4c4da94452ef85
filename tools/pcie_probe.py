"""Host <-> device copy bandwidth per GPU when 1 .. N GPUs copy at once, with and without binding each process to the
NUMA node of its GPU (what bench.py's e2e leg and a multi-GPU caller of afmg_upload / afmg_download depend on):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--mb 1024]

Prints the GPU/NUMA topology, then per mode ("alone": ranks copy one after the other; "together": all at once) the
GB/s of every rank for H2D and D2H from page-locked memory; first with the process wherever the launcher put it, then
re-allocated after sched_setaffinity to the cores of the GPU's NUMA node."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from afivo_streamer_b200 import numa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    if rank == 0:
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout, flush=True)
        except Exception as e:  # noqa: BLE001
            print("nvidia-smi topo failed:", e, flush=True)
    n = args.mb * (1 << 20) // 8
    dev = torch.empty(n, dtype=torch.float64, device="cuda")

    def measure(label):
        host = torch.empty(n, dtype=torch.float64).pin_memory()
        host.fill_(1.0)  # first touch by this process
        res = {}
        for mode in ("alone", "together"):
            for direction in ("h2d", "d2h"):
                best = 0.0
                for _ in range(args.reps):
                    for turn in range(world if mode == "alone" else 1):
                        barrier()
                        if mode == "together" or turn == rank:
                            t0 = time.perf_counter()
                            if direction == "h2d":
                                dev.copy_(host, non_blocking=True)
                            else:
                                host.copy_(dev, non_blocking=True)
                            torch.cuda.synchronize()
                            best = max(best, n * 8 / (time.perf_counter() - t0) / 1e9)
                        barrier()
                res[f"{mode}_{direction}"] = round(best, 1)
        allres = gather(res)
        if rank == 0:
            keys = sorted(allres[0])
            print(json.dumps({"placement": label, "GBs_per_rank": {k: [r[k] for r in allres] for k in keys}}), flush=True)
        del host

    info = gather({"rank": rank, "gpu": local, "cpus_before": len(os.sched_getaffinity(0)), "node": numa.gpu_numa_node(local)})
    if rank == 0:
        print(json.dumps({"ranks": info}), flush=True)
    measure("as launched")
    bound = numa.bind_to_gpu_node(local)
    info = gather({"rank": rank, "bound": bound, "cpus_after": len(os.sched_getaffinity(0))})
    if rank == 0:
        print(json.dumps({"ranks": info}), flush=True)
    measure("bound to the GPU's NUMA node")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
