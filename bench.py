#!/usr/bin/env python
"""Benchmark of the afivo FAS multigrid hot path on B200 (contract: see the task prompt / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload S3|S1|S1r|S3s|S2]

One step = one standalone ``mg_fas_vcycle(set_residual=.true.)`` (afivo/src/m_af_multigrid.f90:185)
on the default workload S3 = BASELINE.json configs[4]: the 1024^3-equivalent refined octree (512^3 uniform +
level 7 on all but the outermost box layer, 16^3 boxes, 1.04e9 cells), the >= 1e8-cell tree the north-star
target is quoted on; it fits one B200 and is used at every N (strong scaling).  ``--workload S1`` is
BASELINE.json configs[1] (``poisson_benchmark 16 16 5``, afivo/examples/poisson_benchmark.f90: uniform 256^3,
rhs = 1, Dirichlet-0), S2 the standard_3d-like channel tree (configs[2], [3]), S0 the 2D cylindrical one.
Metric: cell-updates/s = Gauss-Seidel relaxations per V-cycle x K / device time of the K complete
cycles (ghost fills, transfers, residual, coarse solve included).

  value     V-cycles replayed back to back with all data resident in HBM (CUDA events, max over ranks)
  e2e       the same step through the C ABI with HOST buffers: upload rhs (pinned) -> V-cycle ->
            max-norm of the residual -> download phi, copies inside the timed region
  roofline  dominant kernel = the fused half-sweep k_gsrb2 on the finest level, timed per launch with
            CUDA events (library profiling mode) right after the timed region; achieved = the bytes
            the colour-split layout makes one launch move (COL + 2 NI + 6 NF doubles per box
            = 15 B per cell for nc = 16) / that time; the reference-layout figure of SURVEY 8(d)
            (30 B per cell) is reported next to it as reference_layout_equivalent
  cpu_baseline  the CPU oracle (OpenMP port of the reference path) on the box's host cores

``--impl reference`` times the oracle port on all host cores for the same metric: on the SAME tree when
the host has the memory for it (S3 needs ~50 GB), else on the bounded sample S3s (the Fortran reference
cannot be built: no Fortran compiler here or on the GPU box, profiles/r02a_fortran_probe.txt; Hypre not
vendored).
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


def ref_layout_bytes_gsrb(nc):
    # reference layout: half-sweep pass + the face ghost fill that follows it: 24 + 96/nc B per cell
    return ALGO_BYTES_PER_CELL_HALFSWEEP + 96.0 / nc


def layout_bytes_gsrb_per_box(nc):
    """Algorithmic bytes one half-sweep moves per box in the colour-split record layout (csrc/layout.cuh):
    read the other colour's block (interior + its 6 ghost faces: COL), read the own colour's rhs (NI), write the own
    colour's interior (NI), push six boundary layers into the neighbours' ghost faces (6 NF)."""
    h = nc // 2
    ni, nf = nc * nc * h, nc * h
    col = ni + 6 * nf
    return 8.0 * (col + 2 * ni + 6 * nf)


def algo_bytes_vcycle(tree):
    """SURVEY 8(d) byte model of one V-cycle (set_residual=T): leaf-level cells 305 B (nc=16) + 62 B more on
    cells of boxes that have children; the face-ghost term scales with 1/nc."""
    nc = tree.nc
    ghost = 9 * 96.0 / nc
    leaf_cell = 192 + ghost + 18 + 17 + 24
    parent_extra = 32 + 24 + 96.0 / nc
    total = 0.0
    for l in range(2, tree.highest_lvl + 1):
        ids = tree.lvl_ids[l - 1]
        npar = int(np.count_nonzero(tree.has_children(ids)))
        total += len(ids) * nc ** 3 * leaf_cell + npar * nc ** 3 * parent_extra
    return total


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def host_ram_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


def cpu_baseline(workload):
    name = CPU_SAMPLE.get(workload, workload)
    tree, bc, ids, rhs, desc = build_workload(name)
    n_cpu = 3
    dt, cores, _, _ = time_oracle(tree, bc, ids, rhs, n_cpu, 1)
    cu = cell_updates_vcycle(tree)
    what = f"{n_cpu} full V-cycles (set_residual + max-norm) with the OpenMP oracle after 1 FMG + 1 warm-up on {name}"
    if name != workload:
        what += f" = the same shell-refined octree one level coarser ({tree.n_boxes * tree.nc ** 3 / 1e6:.0f} M cells), " \
                f"a bounded sample of {workload} (cell-updates/s is size-independent at this scale)"
    return {"value": cu * n_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": what,
            "vcycles_per_s_on_sample": n_cpu / dt, "cpu_model": cpu_model()}


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


CPU_SAMPLE = {"S3": "S3s", "S2e": "S2"}  # bounded CPU sample of a workload too large to time on the host in seconds


def build_workload(name, want_rhs=True):
    """(tree, bc, leaf ids, rhs, description) of a named workload; want_rhs=False skips the host-side right-hand
    side (ids, rhs = None, None) for callers that generate it on the device."""
    from afivo_streamer_b200 import tree as T
    from afivo_streamer_b200 import workloads as W
    rhs_kind = "random"
    if name == "S1":
        tree = T.uniform_tree(3, 16, 16, 5)
        bc = W.bc_dirichlet_zero(tree)
        rhs_kind = "one"
        desc = "S1: poisson_benchmark 16 16 5 = uniform 256^3 grid of 16^3 boxes, rhs=1, Dirichlet-0"
    elif name == "S1r":
        tree = T.uniform_tree(3, 16, 16, 5)
        bc = W.bc_field_homogeneous(tree, 1.0)
        desc = "S1r: uniform 256^3 of 16^3 boxes, random rhs, field_bc_homogeneous"
    elif name == "S3s":
        tree = T.shell_tree(16, 16, 5)
        bc = W.bc_field_homogeneous(tree, 1.0)
        desc = "S3s: 256^3 uniform + one refined level inside (512^3-equivalent shell-refined octree)"
    elif name == "S3":
        tree = T.shell_tree(16, 16, 6)
        bc = W.bc_field_homogeneous(tree, 1.0)
        desc = ("S3: 1024^3-equivalent refined octree (BASELINE.json configs[4]): 512^3 uniform (levels 1-6) + level 7 "
                "on all but the outermost box layer, 16^3 boxes, 1.04e9 cells, field_bc_homogeneous; it fits one B200 "
                "and is the >=1e8-cell tree the north-star target is quoted on")
    elif name == "S0":
        tree = T.build_tree(2, 8, [8, 8], 7, None, coord_t=T.AF_CYL)
        bc = W.bc_table(tree, lambda nb, c: (W.AF_BC_DIRICHLET, 0.0 if nb == 3 else 1.0) if (nb - 1) // 2 == 1
                        else (W.AF_BC_NEUMANN, 0.0))
        desc = "S0: 2D cylindrical (BASELINE.json configs[0] stand-in), nc=8, coarse 8^2, 7 uniform levels (512^2 cells)"
    elif name in ("S2", "S2e"):
        tree = T.channel_tree(8, 8, 9, 3)
        bc = W.bc_field_homogeneous(tree, 1.0)
        desc = "S2: standard_3d-like channel-refined tree, nc=8, 9 levels"
        if name == "S2e":
            desc += ("; rod electrode (radius 0.1) from the top plate down to the channel + dielectric slab (eps = 3) "
                     "below z = 0.2: boxes with explicit stencils, built on the device (afmg_build_stencils_device)")
    else:
        raise SystemExit(f"unknown workload {name}")
    ids = rhs = None
    if want_rhs:
        ids, rhs = W.constant_rhs_on_leaves(tree, 1.0) if rhs_kind == "one" else W.random_rhs_on_leaves(tree)
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


_M64 = (1 << 64) - 1
_H1, _H2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB  # splitmix64 finaliser


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def synthetic_rhs_host(leaf_index, box_len):
    """The benchmark's right-hand side on the host: an integer hash (splitmix64 finaliser) of the global index
    leaf * box_len + cell, mapped to (-1, 1).  Pure 64-bit integer arithmetic, so it is bit-identical to
    synthetic_rhs_device(): the CPU arm and the GPU arm (at every N) solve the same problem."""
    idx = (np.asarray(leaf_index, dtype=np.uint64)[:, None] * np.uint64(box_len)
           + np.arange(box_len, dtype=np.uint64)[None, :]) + np.uint64(1)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_H1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_H2)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


def time_oracle(tree, bc, ids, rhs, steps, warmup, budget_s=None, rhs_from_hash=False):
    """CPU oracle: `steps` V-cycles (set_residual + max-norm each), all host threads.  rhs_from_hash: generate the
    benchmark rhs (synthetic_rhs_host) slab by slab instead of taking `rhs`.  budget_s bounds the timed region:
    the loop stops after the step that crosses it.  Returns (seconds, threads, residual history, steps done)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.oracle import I_RHS, I_TMP, Oracle
    orc = Oracle(tree)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its ranks when the variable
    # is unset, which would time a 1-thread run; AFMG_CPU_THREADS overrides
    nthr = int(os.environ.get("AFMG_CPU_THREADS", 0)) or len(os.sched_getaffinity(0))
    orc.set_num_threads(nthr)
    orc.set_bc(bc)
    if rhs_from_hash:
        slab = 2048
        starts = list(range(0, len(ids), slab))

        def put(q0):
            sel = np.arange(q0, min(len(ids), q0 + slab))
            return q0, synthetic_rhs_host(sel, tree.box_len)
        with ThreadPoolExecutor(max_workers=min(nthr, 16)) as ex:
            for q0, data in ex.map(put, starts):
                orc.set_cc(I_RHS, ids[q0:q0 + slab], data)
    else:
        orc.set_cc(I_RHS, ids, rhs)
    orc.mg_init()
    orc.fas_fmg(True, False)
    hist = [orc.maxabs(I_TMP)]
    for _ in range(warmup):
        orc.fas_vcycle(True)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        orc.fas_vcycle(True)
        hist.append(orc.maxabs(I_TMP))
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dt, orc.num_threads(), hist, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same tree as the GPU arm when the host can hold it (S3: 36 GB for the oracle's three variables), else
    # the bounded sample (S3s, the same octree one level coarser)
    need_gb = {"S3": 60.0}.get(args.workload, 0.0)
    same = host_ram_gb() >= need_gb or os.environ.get("AFMG_REF_FULL") == "1"
    if os.environ.get("AFMG_REF_FULL") == "0":
        same = False
    sample_name = args.workload if same else CPU_SAMPLE.get(args.workload, args.workload)
    tree, bc, _, _, desc = build_workload(sample_name, want_rhs=False)
    gpu_desc = build_workload(args.workload, want_rhs=False)[4] if sample_name != args.workload else desc
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    if args.workload == "S1":
        from afivo_streamer_b200 import workloads as W
        ids, rhs = W.constant_rhs_on_leaves(tree, 1.0)
        dt, cores, hist, steps = time_oracle(tree, bc, ids, rhs, args.steps, args.warmup, budget_s=150.0)
    else:
        dt, cores, hist, steps = time_oracle(tree, bc, leaves, None, args.steps, args.warmup, budget_s=150.0,
                                             rhs_from_hash=True)
    cu = cell_updates_vcycle(tree)
    val = cu * steps / dt
    sample = (f"{steps} full V-cycles (set_residual + max-norm) of {sample_name} after 1 FMG + {args.warmup} warm-up "
              f"cycles; same rhs hash as the GPU arm" + ("" if same else
              f"; bounded sample of {args.workload} (host memory {host_ram_gb():.0f} GB < {need_gb:.0f} GB)"))
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": gpu_desc, "n_boxes": tree.n_boxes, "n_cell": tree.nc, "levels": tree.highest_lvl,
                   "cells_all_levels": tree.n_boxes * tree.nc ** 3, "cpu_tree": desc, "same_tree_as_gpu_arm": bool(same)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(), "host_ram_gb": round(host_ram_gb(), 1)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "residual": {"after_fmg": hist[0], "after_timed_cycles": hist[-1]},
        "note": "CPU oracle (C++/OpenMP port of the reference path, oracle/afmg_oracle.cpp) on all host threads; the "
                "Fortran reference cannot be built: no Fortran compiler in the image or on the GPU box "
                "(profiles/r02a_fortran_probe.txt)",
    }
    print(json.dumps(out))


def synthetic_rhs_device(torch, leaf_index, box_len):
    """Deterministic pseudo-random rhs in (-1, 1), a function of the global (leaf index, cell) only, so every GPU
    count -- and the CPU arm, synthetic_rhs_host() -- solves the same problem; generated on the device in packed box
    order.  int64 arithmetic wraps like uint64; logical right shifts are arithmetic shifts + mask."""
    base = torch.as_tensor(np.asarray(leaf_index, dtype=np.int64) * box_len, device="cuda")
    z = (base[:, None] + torch.arange(box_len, dtype=torch.int64, device="cuda")[None, :]) + 1

    def lsr(x, n):
        return torch.bitwise_and(torch.bitwise_right_shift(x, n), (1 << (64 - n)) - 1)
    z = z * _s64(0x9E3779B97F4A7C15)
    z = torch.bitwise_xor(z, lsr(z, 30)) * _s64(_H1)
    z = torch.bitwise_xor(z, lsr(z, 27)) * _s64(_H2)
    z = torch.bitwise_xor(z, lsr(z, 31))
    return (lsr(z, 11).to(torch.float64) * (2.0 ** -52) - 1.0).reshape(-1)


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
    # page-locked buffers next to the GPU: bind this process to the GPU's NUMA node before anything is allocated
    # (no-op on a single-node host such as the B200 boxes of this pool, profiles/r02r_pcie_probe_n8.txt)
    from afivo_streamer_b200 import numa
    numa_node = numa.bind_to_gpu_node(local)
    comm = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner out of the one-JSON-line stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = M.comm_from_torch()  # plumbing only: exchanges the CUDA IPC handles of the ranks' arrays

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    tree, bc, _, _, desc = build_workload(args.workload, want_rhs=False)
    # --single-process: ONE process drives all GPUs (afmg_opts.n_gpus), the mode a single-process caller like the
    # reference uses; launched without torchrun
    sp_gpus = args.gpus if (args.single_process and world == 1) else 0
    mg = M.mg_t(sides_bc=bc, device=local, comm=comm, lsf_boundary_value=1.0, n_gpus=sp_gpus)
    M.mg_init(tree, mg)
    explicit = None
    if args.workload == "S2e":
        # variable-coefficient / level-set boxes.  Permittivity up (one variable, once per refinement),
        # then tags + operator + prolongation stencils are built where the data lives.
        from afivo_streamer_b200 import stencils as St
        from afivo_streamer_b200 import workloads as Wk
        boxes = np.concatenate(tree.lvl_ids).astype(np.int32)
        ctr = Wk.cell_centres(tree, boxes, ghosts=True)
        mg.set_cc(M.I_EPS, boxes, np.where(ctr[..., 2] < 0.2, 3.0, 1.0))
        el = St.electrode("rod", 3, rod_r0=(0.5, 0.5, 0.7), rod_r1=(0.5, 0.5, 1.2), rod_radius=0.1)
        t0 = time.perf_counter()
        mg.build_stencils_device(el)
        t_build = time.perf_counter() - t0
        b_ids, b_tags, b_meta, _ = mg.built_stencils()
        explicit = {"boxes_with_tags": int(len(b_ids)), "electrode_boxes": int(np.count_nonzero(b_tags & 1)),
                    "variable_eps_boxes": int(np.count_nonzero(b_tags & 2)), "constant_eps_boxes": int(np.count_nonzero(b_tags & 4)),
                    "variable_operators": int(np.count_nonzero(b_meta[:, 0] == 2)), "of_boxes": int(tree.n_boxes),
                    "device_build_ms": 1e3 * t_build,
                    "what": "afmg_build_stencils_device: tags, operator and prolongation stencils from the resident "
                            "permittivity and the built-in rod electrode (blocking call, includes the ingestion)"}
    # this rank's leaves, in the tree's (level, list) order
    leaves = np.concatenate([tree.leaves(l) for l in range(1, tree.highest_lvl + 1)]).astype(np.int32)
    if world > 1:
        owner = mg.owners(leaves)
        sel = np.nonzero(owner == rank)[0]
    else:
        sel = np.arange(len(leaves))
    ids = np.ascontiguousarray(leaves[sel])
    box_len = tree.box_len
    # rhs: constant 1 for S1 (poisson_benchmark), deterministic pseudo-random otherwise; made on the device
    chunk = 4096
    h_rhs = torch.empty(len(ids) * box_len, dtype=torch.float64).pin_memory()
    for q0 in range(0, len(ids), chunk):
        q1 = min(len(ids), q0 + chunk)
        if args.workload == "S1":
            d = torch.ones((q1 - q0) * box_len, dtype=torch.float64, device="cuda")
        else:
            d = synthetic_rhs_device(torch, sel[q0:q1], box_len)
        h_rhs[q0 * box_len:q1 * box_len].copy_(d)
        del d
    torch.cuda.synchronize()
    nbytes = h_rhs.numel() * 8

    barrier()
    mg.upload_ptr(M.I_RHS, ids, h_rhs.data_ptr())
    barrier()
    M.mg_fas_fmg(tree, mg, True, False)  # start-up solve as in field_compute (src/m_field.f90:491-517)
    res0 = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)

    # ---- device-resident V-cycles -----------------------------------------------------------
    barrier()
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
    ms_max = allmax(ms)
    launches = int(allsum(l1 - l0))
    res1 = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
    # bitwise checksum of phi (whole records, ghost cells included) over all boxes of all ranks after the same
    # W + K cycles: equal at every N when the partitioned solve is bit-identical to the single-GPU one
    csum, cxor = mg.checksum(M.I_PHI)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (csum, cxor))
        csum, cxor = 0, 0
        for a, b in parts:
            csum = (csum + a) & _M64
            cxor ^= b
    cu = mg.cell_updates(0, False)  # whole tree, all ranks together
    value = cu * args.steps / (ms_max * 1e-3)

    # ---- FMG timing (secondary figure) ------------------------------------------------------
    n_fmg = max(2, args.steps // 4)
    barrier()
    mg.fas_fmg_async(False, True, 1)
    mg.sync()
    barrier()
    mg.fas_fmg_async(False, True, n_fmg)
    mg.sync()
    fmg_ms = allmax(mg.last_cycle_ms()) / n_fmg
    cu_fmg = mg.cell_updates(0, True)

    # ---- the callers' next step, field_from_potential (src/m_field.f90:531-548), on the device (1 GPU) ------
    field = None
    if world == 1 and tree.ndim == 3 and sp_gpus <= 1:
        M.field_from_potential(tree, mg, -1.0)  # allocates fc / norm
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            M.field_from_potential(tree, mg, -1.0)  # blocking C-ABI call
        f_ms = 1e3 * (time.perf_counter() - t0) / reps
        nc = tree.nc
        per_cell = 8 * ((nc + 2) / nc) ** 3 + 3 * 8 * (nc + 1) / nc + 8  # read phi, write fc and the norm
        field = {"ms": f_ms, "algorithmic_GBs": per_cell * tree.n_boxes * nc ** 3 / (f_ms * 1e-3) / 1e9,
                 "what": "gradient + norm on all boxes, ghost cells of the norm on all levels"}

    # ---- BASELINE.json configs[3]: the three Helmholtz photoionization solves on the standard_3d-like tree ----
    helm = None
    if world == 1 and args.workload == "S2":
        lambdas, coeffs = M.photoi_helmh_parameters("Bourdon-3", frac_O2=0.2, gas_pressure=1.0)  # 1/m, 1/m^2
        lambdas, coeffs = lambdas * 0.02, coeffs * 0.02 ** 2                                      # 2 cm domain
        from afivo_streamer_b200 import workloads as Wk
        hbc = Wk.bc_table(tree, M.photoi_helmh_bc)
        modes = []
        for lam in lambdas:
            m = M.mg_t(sides_bc=hbc, device=local, helmholtz_lambda=float(lam ** 2), prolongation_type=M.MG_PROLONG_LINEAR)
            M.mg_init(tree, m)
            modes.append(m)
        modes[0].upload_ptr(M.I_RHS, ids, h_rhs.data_ptr())
        t0 = time.perf_counter()
        n_cold, _ = M.photoi_helmh_compute(tree, modes, coeffs, 10, 1.0e-2)   # from zero modes (first time step)
        t_cold = time.perf_counter() - t0
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            n_warm, _ = M.photoi_helmh_compute(tree, modes, coeffs, 10, 1.0e-2)  # from the previous modes (every later step)
        t_warm = (time.perf_counter() - t0) / reps
        helm = {"what": "photoi_helmh_compute, 3 modes (src/m_photoi_helmh.f90:162-204), blocking C-ABI call",
                "first_call_ms": 1e3 * t_cold, "fmg_cycles_first": [int(x) for x in n_cold],
                "steady_ms": 1e3 * t_warm, "fmg_cycles_steady": [int(x) for x in n_warm]}
        for m in modes:
            M.mg_destroy(m)

    # ---- e2e: host buffers through the C ABI -----------------------------------------------
    # one step = upload rhs (interior cells of this rank's leaves: ghost cells of rhs are never read) -> V-cycle ->
    # max-norm of the residual -> download phi (interior cells of the leaves; its ghost cells stay valid on the
    # device).  Pinned host buffers; the library overlaps the PCIe copy of one chunk with the (un)pack kernel of the
    # previous one.  All four calls are blocking, so host wall time is the end-to-end time.
    ncell = tree.nc ** tree.ndim
    h_rhs_int = torch.empty(len(ids) * ncell, dtype=torch.float64).pin_memory()
    h_phi_int = torch.empty(len(ids) * ncell, dtype=torch.float64).pin_memory()
    nd = tree.ndim
    full = h_rhs.view((len(ids),) + (tree.nc + 2,) * nd)
    h_rhs_int.view((len(ids),) + (tree.nc,) * nd).copy_(full[(slice(None),) + (slice(1, -1),) * nd])
    nbytes_up = h_rhs_int.numel() * 8
    nbytes_dn = h_phi_int.numel() * 8
    parts = np.zeros(4)

    def e2e_step():
        t0 = time.perf_counter()
        mg.upload_interior_ptr(M.I_RHS, ids, h_rhs_int.data_ptr())
        t1 = time.perf_counter()
        M.mg_fas_vcycle(tree, mg, True)
        t2 = time.perf_counter()
        r = M.af_tree_maxabs_cc(tree, mg, M.I_TMP)
        t3 = time.perf_counter()
        mg.download_interior_ptr(M.I_PHI, ids, h_phi_int.data_ptr())
        t4 = time.perf_counter()
        parts[:] += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
        return r

    big = nbytes > (2 << 30)
    for _ in range(1 if big else 3):
        barrier()
        e2e_step()
    e2e_steps = 3 if big else max(3, min(args.steps, 10))
    parts[:] = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0  # the C ABI calls are blocking: host wall time == end-to-end time
    wall_max = allmax(wall)
    e2e_val = cu * e2e_steps / wall_max
    h2d_total, d2h_total = allsum(nbytes_up), allsum(nbytes_dn + 8)
    e2e_parts = {k: allmax(1e3 * v / e2e_steps) for k, v in zip(("upload_ms", "vcycle_ms", "maxnorm_ms", "download_ms"), parts)}
    e2e_parts["pcie_GBs_per_gpu"] = (nbytes_up + nbytes_dn) / 1e9 / max(1e-9, (parts[0] + parts[3]) / e2e_steps)
    # ---- the Fortran shim's call sequence (fortran/m_af_multigrid_gpu.f90), emulated on one GPU: per field solve it
    # (1) packs box%cc(i_rhs) of the leaves into its page-locked buffer, (2) afmg_upload (whole records),
    # (3) the cycle + max-norm, (4) afmg_download of phi on ALL boxes with ghost cells (the reference's post-condition)
    # and unpacks it into the boxes, (5) the same for i_tmp unless switched off (mg_gpu_set_download_tmp).  phi is not
    # uploaded: the device copy is current between solves.  Packing / unpacking = one host memcpy pass each.
    shim = None
    if world == 1 and tree.ndim == 3 and sp_gpus <= 1:
        import ctypes as C
        from afivo_streamer_b200 import _lib
        Lb = _lib.lib()
        allb = np.concatenate(tree.lvl_ids).astype(np.int32)
        n_up, n_dn = len(ids) * box_len, len(allb) * box_len
        p_buf = Lb.afmg_host_alloc(max(n_up, n_dn) * 8)
        pinned = np.ctypeslib.as_array(C.cast(p_buf, C.POINTER(C.c_double)), shape=(max(n_up, n_dn),))
        host_rhs = h_rhs.numpy()                      # stands for the boxes' own (pageable) cc arrays
        host_phi = np.empty(n_dn)
        tt = np.zeros(6)

        from concurrent.futures import ThreadPoolExecutor
        nthr = len(os.sched_getaffinity(0))
        pool = ThreadPoolExecutor(max_workers=nthr)

        def pcopy(dst, src):  # the shim packs with `!$omp parallel do` over the boxes: all host threads
            n, step = len(src), -(-len(src) // nthr)
            list(pool.map(lambda q: np.copyto(dst[q:q + step], src[q:q + step]), range(0, n, step)))

        def shim_step(with_tmp):
            t = [time.perf_counter()]
            pcopy(pinned[:n_up], host_rhs); t.append(time.perf_counter())
            mg.upload_ptr(M.I_RHS, ids, p_buf); t.append(time.perf_counter())
            M.mg_fas_vcycle(tree, mg, True)
            M.af_tree_maxabs_cc(tree, mg, M.I_TMP); t.append(time.perf_counter())
            mg.download_ptr(M.I_PHI, allb, p_buf); t.append(time.perf_counter())
            pcopy(host_phi, pinned[:n_dn]); t.append(time.perf_counter())
            if with_tmp:
                mg.download_ptr(M.I_TMP, allb, p_buf)
                pcopy(host_phi, pinned[:n_dn])
            t.append(time.perf_counter())
            tt[:] += np.diff(t)

        shim_step(True)
        reps = 2 if big else 5
        out_s = {}
        for with_tmp in (True, False):
            tt[:] = 0
            t0 = time.perf_counter()
            for _ in range(reps):
                shim_step(with_tmp)
            out_s["ms_per_solve_with_tmp" if with_tmp else "ms_per_solve_phi_only"] = 1e3 * (time.perf_counter() - t0) / reps
            if not with_tmp:
                out_s["phases_ms_phi_only"] = dict(zip(("pack_rhs", "upload_rhs", "cycle_and_norm", "download_phi", "unpack_phi", "tmp"),
                                                       (1e3 * tt / reps).round(3).tolist()))
        out_s["bytes"] = {"h2d": n_up * 8, "d2h_phi": n_dn * 8}
        out_s["host_threads"] = nthr
        pool.shutdown()
        out_s["what"] = ("fortran/m_af_multigrid_gpu.f90 call sequence for one mg_fas_vcycle: rhs of the leaves up and phi of "
                         "all boxes down as whole records through a page-locked packing buffer, host pack / unpack passes "
                         "included; with_tmp also brings i_tmp back (the reference's set_residual post-condition)")
        shim = out_s
        del pinned
        Lb.afmg_host_free(p_buf)

    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing with CUDA events (library profiling mode, no graph) -----------------
    barrier()
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
    if g_calls and tree.ndim == 3:
        own = mg.own_boxes(tree.highest_lvl)   # finest-level boxes this rank sweeps per launch
        if sp_gpus > 1:
            own = int(np.count_nonzero(mg.owners(tree.lvl_ids[-1]) == 0))
        cells = own * tree.nc ** 3
        per_launch_bytes = layout_bytes_gsrb_per_box(tree.nc) * own
        dur = g_ms / g_calls * 1e-3
        achieved = per_launch_bytes / dur / 1e9
        # DRAM traffic of one launch from the ncu --set full capture of this kernel (dram__bytes_read + write, N = 1),
        # per box, times the boxes THIS rank sweeps: profiles/gsrb_traffic.json says which capture it came from
        traffic = traffic_src = None
        tpath = os.path.join(ROOT, "profiles", "gsrb_traffic.json")
        if os.path.exists(tpath):
            try:
                t = json.load(open(tpath)).get(args.workload)
                if t:
                    traffic = t["dram_bytes_per_box"] * own
                    traffic_src = t["source"]
            except Exception:
                traffic = None
        # whole V-cycle: DRAM bytes of all its launches from the ncu launch list (N = 1), against the cycle time
        whole = {"reference_layout_bytes_per_vcycle": algo_bytes_vcycle(tree)}
        wpath = os.path.join(ROOT, "profiles", "vcycle_dram.json")
        if os.path.exists(wpath):
            try:
                w = json.load(open(wpath)).get(args.workload)
                if w:
                    per_gpu = w["dram_bytes_per_vcycle"] / max(world, sp_gpus)
                    cyc_s = ms_max / args.steps * 1e-3
                    whole.update({"dram_bytes_per_gpu": per_gpu, "dram_GBs_per_gpu": per_gpu / cyc_s / 1e9,
                                  "frac_per_gpu": per_gpu / cyc_s / 1e9 / peak, "source": w["source"]})
            except Exception:
                pass
        roof = {"bound": "hbm", "kernel": "k_gsrb2 (finest level, rank 0)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "algorithmic_bytes_per_cell": layout_bytes_gsrb_per_box(tree.nc) / tree.nc ** 3,
                "launch_us": dur * 1e6, "boxes_per_launch": own,
                "share_of_step": g_ms / total_prof if total_prof else None,
                "reference_layout_equivalent": {
                    "bytes_per_cell": ref_layout_bytes_gsrb(tree.nc),
                    "GBs": ref_layout_bytes_gsrb(tree.nc) * cells / dur / 1e9,
                    "note": "SURVEY 8(d) counts 24 + 96/nc B per cell for the reference's record layout; the "
                            "colour-split layout moves half of it, so this figure is not a roofline fraction"},
                "whole_cycle": whole}
    barrier_stat = None
    if world > 1:
        b_ms, b_calls = prof.get("barrier", (0.0, 0))
        barrier_stat = {"ms_per_cycle_rank0": b_ms / nprof, "count_per_cycle": b_calls / nprof}

    # ---- CPU baseline on the host cores (rank 0, N=1 only): bounded sample ---------------------
    cpu = None
    if rank == 0 and world == 1 and sp_gpus <= 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload)


    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, sp_gpus), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_boxes": tree.n_boxes, "n_cell": tree.nc, "levels": tree.highest_lvl,
                       "cells_all_levels": tree.n_boxes * tree.nc ** 3,
                       "cells_finest": tree.n_cells_level(tree.highest_lvl), "step": "mg_fas_vcycle(set_residual=T)",
                       "parallelism": (f"ONE process, {sp_gpus} GPUs (afmg_opts.n_gpus): one host thread per GPU inside the library, "
                                       "Morton partition, halos via NVLink peer memory") if sp_gpus > 1 else
                       "single GPU" if world == 1 else
                       f"boxes partitioned over {world} GPUs by Morton ranges per level; halos via NVLink peer memory",
                       "l2_policy": f"working set {3 * tree.n_boxes * box_len * 8 / world / 1e6:.0f} MB per GPU "
                                    "(3 variables) exceeds the 126 MB L2; no flush needed"},
            "slab_GB_per_gpu": {"mapped": (mg.slab_bytes()[0] / 1e9).round(2).tolist(), "full_slot_space": float(mg.slab_bytes()[1][0] / 1e9)},
            "vcycles_per_s": args.steps / (ms_max * 1e-3),
            "fmg": {"ms": fmg_ms, "cell_updates_per_s": cu_fmg / (fmg_ms * 1e-3)},
            "field_from_potential": field,
            "helmholtz_photoionization": helm,
            "explicit_stencils": explicit,
            "residual": {"after_fmg": res0, "after_timed_cycles": res1},
            "phi_checksum": {"sum_u64": f"{csum:016x}", "xor_u64": f"{cxor:016x}",
                             "what": "wrapping sum / xor of the bit patterns of phi over the complete records of all "
                                     "boxes, all ranks combined, after the FMG + W + K cycles"},
            "barrier": barrier_stat,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                    "steps": e2e_steps, "ms_per_step": 1e3 * wall_max / e2e_steps, "phases_max_over_ranks": e2e_parts,
                    "host_numa_node_bound": numa_node,
                    "what": "upload rhs (interior, leaves) -> V-cycle -> residual max-norm -> download phi (interior, "
                            "leaves); pinned host buffers, blocking C-ABI calls"},
            "shim_sequence": shim,
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "kernel_profile_ms_per_cycle_rank0": {k: v[0] / nprof for k, v in sorted(prof.items())},
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
    ap.add_argument("--workload", default="S3")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N and no torchrun: one process drives N GPUs through afmg_opts.n_gpus")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
