#!/usr/bin/env python
"""Benchmark of the afivo FAS multigrid hot path on B200 (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload S1|S1r|S3s]

One step = one standalone ``mg_fas_vcycle(set_residual=.true.)`` (afivo/src/m_af_multigrid.f90:185)
on the 3D Poisson benchmark tree of BASELINE.json configs[1]: uniform 256^3 grid of 16^3 boxes
(``poisson_benchmark 16 16 5``, afivo/examples/poisson_benchmark.f90), rhs = 1, Dirichlet-0.
Metric: cell-updates/s = Gauss-Seidel relaxations per V-cycle x K / device time of the K complete
cycles (ghost fills, transfers, residual, coarse solve included).

  value     V-cycles replayed back to back with all data resident in HBM (CUDA events, max over ranks)
  e2e       the same step through the C ABI with HOST buffers: upload rhs (pinned) -> V-cycle ->
            max-norm of the residual -> download phi, copies inside the timed region
  roofline  dominant kernel = the fused half-sweep k_gsrb on the finest level, timed per launch with
            CUDA events (library profiling mode) right after the timed region
  cpu_baseline  the CPU oracle (OpenMP port of the reference path) on the box's host cores

``--impl reference`` times the oracle port on the host cores for the same workload and metric (the
Fortran reference cannot be built here: no Fortran compiler, Hypre not vendored).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3D FAS multigrid cell-updates/s"
UNIT = "cell-updates/s"
ALGO_BYTES_PER_CELL_HALFSWEEP = 24.0  # R phi, R rhs, W phi


def algo_bytes_gsrb(nc):
    # fused kernel = half-sweep pass + the face ghost fill that follows it: 24 + 96/nc B per cell
    return ALGO_BYTES_PER_CELL_HALFSWEEP + 96.0 / nc


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def build_workload(name):
    from afivo_streamer_b200 import tree as T
    from afivo_streamer_b200 import workloads as W
    if name == "S1":
        tree = T.uniform_tree(3, 16, 16, 5)
        bc = W.bc_dirichlet_zero(tree)
        ids, rhs = W.constant_rhs_on_leaves(tree, 1.0)
        desc = "S1: poisson_benchmark 16 16 5 = uniform 256^3 grid of 16^3 boxes, rhs=1, Dirichlet-0"
    elif name == "S1r":
        tree = T.uniform_tree(3, 16, 16, 5)
        bc = W.bc_field_homogeneous(tree, 1.0)
        ids, rhs = W.random_rhs_on_leaves(tree)
        desc = "S1r: uniform 256^3 of 16^3 boxes, random rhs, field_bc_homogeneous"
    elif name == "S3s":
        tree = T.shell_tree(16, 16, 5)
        bc = W.bc_field_homogeneous(tree, 1.0)
        ids, rhs = W.random_rhs_on_leaves(tree)
        desc = "S3s: 256^3 uniform + one refined level inside (512^3-equivalent shell-refined octree)"
    elif name == "S2":
        tree = T.channel_tree(8, 8, 8, 3)
        bc = W.bc_field_homogeneous(tree, 1.0)
        ids, rhs = W.random_rhs_on_leaves(tree)
        desc = "S2: standard_3d-like channel-refined tree, nc=8, 8 levels"
    else:
        raise SystemExit(f"unknown workload {name}")
    return tree, bc, ids, rhs, desc


def cell_updates_vcycle(tree, n_down=2, n_up=2):
    return float(sum((n_down + n_up) * tree.n_cells_level(l) for l in range(2, tree.highest_lvl + 1)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_oracle(tree, bc, ids, rhs, steps, warmup):
    """CPU oracle: `steps` V-cycles (set_residual + max-norm each), all host threads."""
    from oracle.oracle import I_RHS, I_TMP, Oracle
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.set_cc(I_RHS, ids, rhs)
    orc.mg_init()
    orc.fas_fmg(True, False)
    for _ in range(warmup):
        orc.fas_vcycle(True)
    t0 = time.perf_counter()
    res = None
    for _ in range(steps):
        orc.fas_vcycle(True)
        res = orc.maxabs(I_TMP)
    dt = time.perf_counter() - t0
    return dt, orc.num_threads(), res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tree, bc, ids, rhs, desc = build_workload(args.workload)
    steps = max(1, min(args.steps, 10))
    dt, cores, _ = time_oracle(tree, bc, ids, rhs, steps, min(args.warmup, 1))
    cu = cell_updates_vcycle(tree)
    val = cu * steps / dt
    sample = f"{steps} full V-cycles (set_residual + max-norm) of {args.workload} after 1 FMG"
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_boxes": tree.n_boxes, "n_cell": tree.nc, "levels": tree.highest_lvl},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle (C++/OpenMP port of the reference path); the Fortran reference cannot be built here",
    }
    print(json.dumps(out))


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from afivo_streamer_b200 import mg as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    tree, bc, ids, rhs, desc = build_workload(args.workload)
    mg = M.mg_t(sides_bc=bc, device=local)
    M.mg_init(tree, mg)
    nbytes = rhs.size * 8
    # pinned host buffers for the e2e leg
    h_rhs = torch.empty(rhs.size, dtype=torch.float64).pin_memory()
    h_rhs.numpy()[:] = rhs.reshape(-1)
    h_phi = torch.empty(rhs.size, dtype=torch.float64).pin_memory()

    mg.upload_ptr(M.I_RHS, ids, h_rhs.data_ptr())
    M.mg_fas_fmg(tree, mg, True, False)  # start-up solve as in field_compute (src/m_field.f90:491-517)
    res0 = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident V-cycles -----------------------------------------------------------
    mg.fas_vcycle_async(True, 0, args.warmup)
    mg.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = mg.kernel_launches()
    mg.fas_vcycle_async(True, 0, args.steps)
    mg.sync()
    ms = mg.last_cycle_ms()
    l1 = mg.kernel_launches()
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    res1 = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
    cu = mg.cell_updates(0, False)
    value = world * cu * args.steps / (ms_max * 1e-3)

    # ---- FMG timing (secondary figure) ------------------------------------------------------
    mg.fas_fmg_async(False, True, 1)
    mg.sync()
    mg.fas_fmg_async(False, True, max(2, args.steps // 4))
    mg.sync()
    fmg_ms = mg.last_cycle_ms() / max(2, args.steps // 4)
    cu_fmg = mg.cell_updates(0, True)

    # ---- e2e: host buffers through the C ABI -----------------------------------------------
    def e2e_step():
        mg.upload_ptr(M.I_RHS, ids, h_rhs.data_ptr())
        M.mg_fas_vcycle(tree, mg, True)
        r = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        mg.download_ptr(M.I_PHI, ids, h_phi.data_ptr())
        return r

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0  # the C ABI calls are blocking: host wall time == end-to-end time
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * cu * e2e_steps / float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing with CUDA events (library profiling mode, no graph) -----------------
    mg.set_profiling(True)
    nprof = 3
    mg.fas_vcycle_async(True, 0, nprof)
    mg.sync()
    prof = mg.profile()
    mg.set_profiling(False)
    total_prof = sum(v[0] for v in prof.values())
    top = f"gsrb_L{tree.highest_lvl}"
    g_ms, g_calls = prof.get(top, (0.0, 0))
    peak, peak_src = peaks()
    roof = None
    if g_calls:
        cells = tree.n_cells_level(tree.highest_lvl)
        per_launch_bytes = algo_bytes_gsrb(tree.nc) * cells
        dur = g_ms / g_calls * 1e-3
        achieved = per_launch_bytes / dur / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gsrb_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "k_gsrb (finest level)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_launch_bytes, "launch_us": dur * 1e6,
                "share_of_step": g_ms / total_prof if total_prof else None,
                "whole_cycle": {"algorithmic_bytes_per_cell_update": 78.0,
                                "achieved_GBs": value / world * 78.0 / 1e9, "frac": value / world * 78.0 / 1e9 / peak}}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = 5
        dt, cores, _ = time_oracle(tree, bc, ids, rhs, n_cpu, 1)
        cpu = {"value": cu * n_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full V-cycles of {args.workload} with the OpenMP oracle after 1 FMG + 1 warm-up",
               "vcycles_per_s": n_cpu / dt}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_boxes": tree.n_boxes, "n_cell": tree.nc, "levels": tree.highest_lvl,
                       "cells_finest": tree.n_cells_level(tree.highest_lvl), "step": "mg_fas_vcycle(set_residual=T)",
                       "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (weak)",
                       "l2_policy": "working set 655 MB (3 variables) exceeds the 126 MB L2; no flush needed"},
            "vcycles_per_s": world * args.steps / (ms_max * 1e-3),
            "fmg": {"ms": fmg_ms, "cell_updates_per_s": cu_fmg / (fmg_ms * 1e-3)},
            "residual": {"after_fmg": res0, "after_timed_cycles": res1},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 8,
                    "steps": e2e_steps, "ms_per_step": 1e3 * float(t.item()) / e2e_steps},
            "gpu_launches": l1 - l0,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "kernel_profile_ms_per_cycle": {k: v[0] / nprof for k, v in sorted(prof.items())},
        }
        print(json.dumps(out))
    M.mg_destroy(mg)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="S1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
