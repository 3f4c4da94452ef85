"""Summarise an .ncu-rep (raw page) and a launch-list csv into profiles/*.md / *.csv."""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def to_bytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * m.get(unit, 1)


def rep_summary(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        d = {"kernel": name}
        for w in WANT:
            if w in col:
                d[w] = (r[col[w]], units[col[w]])
        out.append(d)
    return out


def launches_summary(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        grid = r[col["Grid Size"]] if "Grid Size" in col else ""
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        agg[(name, grid)][0] += 1
        agg[(name, grid)][1] += v
    return agg


def vcycle_windows(path, finest_grid, all_grid):
    """Split a launch list of bench.py into V-cycles (each ends with the final residual over all boxes) and give,
    per window, the summed duration and the share of the finest-level half-sweeps: the figure bench.py reports
    as roofline.share_of_step from CUDA events."""
    rows = list(csv.reader(open(path, errors="ignore")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[start])}
    seq = []
    for r in rows[start + 1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[col["Metric Unit"]], 1.0)
        seq.append((r[col["Kernel Name"]], r[col["Grid Size"]], v))
    ends = [i for i, (n, g, _) in enumerate(seq) if "k_resid3" in n and ", 0, " in n.split("(")[0] and g.startswith(f"({all_grid},")]
    out = []
    for a, b in zip(ends[:-1], ends[1:]):
        w = seq[a + 1:b + 1]
        tot = sum(v for _, _, v in w)
        top = sum(v for n, g, v in w if "k_gsrb2" in n and g.startswith(f"({finest_grid},"))
        out.append((len(w), tot, top / tot))
    return out


def vcycle_dram(path, all_grid, workload=None, out_json=None):
    """Launch list taken with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum: per complete
    V-cycle (the launches between two final residuals over all boxes) the summed duration and DRAM bytes, and a
    per-kernel table of the last complete cycle.  Optionally records the figure in a json bench.py reads."""
    import json
    rows = list(csv.reader(open(path, errors="ignore")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[start])}
    by_id = {}
    order = []
    for r in rows[start + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        i = r[col["ID"]]
        if i not in by_id:
            by_id[i] = {"name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "us": 0.0, "rd": 0.0, "wr": 0.0}
            order.append(i)
        m, u = r[col["Metric Name"]], r[col["Metric Unit"]]
        v = float(r[col["Metric Value"]].replace(",", ""))
        if m == "gpu__time_duration.sum":
            by_id[i]["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            by_id[i]["rd" if "read" in m else "wr"] = b
    seq = [by_id[i] for i in order]
    ends = [i for i, d in enumerate(seq)
            if "k_resid3" in d["name"] and ", 0, " in d["name"].split("(")[0] and d["grid"].startswith(f"({all_grid},")]
    wins = []
    for a, b in zip(ends[:-1], ends[1:]):
        w = seq[a + 1:b + 1]
        wins.append((len(w), sum(d["us"] for d in w), sum(d["rd"] + d["wr"] for d in w), w))
    # complete V-cycles all have the same number of launches; other windows (FMG, set-up) differ
    if not wins:
        print("no complete V-cycle in the list")
        return
    from collections import Counter
    nl = Counter(w[0] for w in wins).most_common(1)[0][0]
    cyc = [w for w in wins if w[0] == nl]
    for n, us, by, _ in cyc:
        print(f"{n} launches, {us / 1e3:.3f} ms (serialised, cold cache), DRAM {by / 1e9:.3f} GB "
              f"-> {by / us / 1e3:.0f} GB/s over the serialised time")
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in cyc[-1][3]:
        k = d["name"].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += d["us"]
        agg[k][2] += d["rd"] + d["wr"]
    print("| kernel | launches | us | DRAM MB | GB/s |")
    print("|---|---|---|---|---|")
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us:.1f} | {by / 1e6:.1f} | {by / us / 1e3 if us else 0:.0f} |")
    if out_json and workload:
        try:
            cur = json.load(open(out_json))
        except Exception:
            cur = {}
        med = sorted(w[2] for w in cyc)[len(cyc) // 2]
        cur[workload] = {"dram_bytes_per_vcycle": med, "launches_per_vcycle": nl,
                         "serialised_ms": sorted(w[1] for w in cyc)[len(cyc) // 2] / 1e3,
                         "source": f"ncu launch list {path} (dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                                   f"{nl} launches of one V-cycle, median of {len(cyc)} cycles, N = 1)"}
        json.dump(cur, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    if kind == "vcycle_dram":  # vcycle_dram <csv> <all_boxes grid> [workload out.json]
        vcycle_dram(path, sys.argv[3], *(sys.argv[4:6]))
        sys.exit(0)
    if kind == "vcycle":
        for n, tot, share in vcycle_windows(path, sys.argv[3], sys.argv[4]):
            print(f"{n} launches, {tot / 1e3:.2f} ms, finest-level k_gsrb2 share {100 * share:.1f} %")
        sys.exit(0)
    if kind == "rep":
        for d in rep_summary(path):
            t, tu = d["gpu__time_duration.sum"]
            rd = to_bytes(*d["dram__bytes_read.sum"])
            wr = to_bytes(*d["dram__bytes_write.sum"])
            us = float(t) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(tu, 1)
            print(f"| {d['kernel']} | grid {d['launch__grid_size'][0]} x {d['launch__block_size'][0]} | {us:.1f} us | "
                  f"dram R {rd / 1e6:.1f} MB W {wr / 1e6:.1f} MB ({(rd + wr) / us / 1e3:.0f} GB/s) | "
                  f"regs {d['launch__registers_per_thread'][0]} | warps {float(d['sm__warps_active.avg.pct_of_peak_sustained_active'][0]):.0f}% | "
                  f"dram {float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]):.0f}% | sm {float(d['sm__throughput.avg.pct_of_peak_sustained_elapsed'][0]):.0f}% |")
    else:
        agg = launches_summary(path)
        total = sum(v[1] for v in agg.values())
        print("| kernel | grid | launches | total us | share |")
        print("|---|---|---|---|---|")
        for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| {name} | {grid} | {n} | {us:.1f} | {100 * us / total:.1f}% |")
        print(f"| total | | {sum(v[0] for v in agg.values())} | {total:.1f} | |")
