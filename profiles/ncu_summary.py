#!/usr/bin/env python
"""Summarise ncu output brought back from the GPU box (gpurun_out/) into text files under profiles/.

    python profiles/ncu_summary.py launches gpurun_out/X_launches.csv          > profiles/X_launches.txt
    python profiles/ncu_summary.py kernel   gpurun_out/X_kernel.ncu-rep [...]  > profiles/X_ncu.txt
"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache launches): {path}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:66s} launches={v[0]:5d} total_ms={v[1]:10.3f} share={v[1] / tot:6.1%} avg_ms={v[1] / v[0]:.3f}")
    print(f"{'TOTAL':66s} launches={sum(v[0] for v in agg.values()):5d} total_ms={tot:10.3f}")


def kernel(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            print(f"## {p}: {name[:100]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"{w:86s} {vals[i]:>18s} {units[i]}")
            print()


def dominant_json(paths, out_path, what):
    """profiles/ncu_dominant_launch.json for bench.py's roofline.traffic: dram bytes, duration and pipe utilisation of the given
    --set full captures (one launch each: forward + backtrace of the dominant K1 launch), summed."""
    import json
    total = {"what": what, "captures": [], "dram_bytes_per_launch": 0.0, "time_ms": 0.0}
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        def get(name):
            return float(vals[hdr.index(name)].replace(",", "")) if name in hdr else None
        def scaled(name):
            v = get(name)
            u = units[hdr.index(name)] if name in hdr else ""
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
            return None if v is None else v * mult
        rec = {"file": p.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")][:80],
               "time_ms": scaled("gpu__time_duration.sum"), "dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum"),
               "alu_pipe_pct": get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), "warp_inst_per_cycle_per_sm": get("sm__inst_executed.avg.per_cycle_elapsed"),
               "active_lanes_per_warp_inst": get("smsp__thread_inst_executed_per_inst_executed.ratio"), "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "registers": get("launch__registers_per_thread"), "grid": get("launch__grid_size"), "block": get("launch__block_size")}
        total["captures"].append(rec)
        total["dram_bytes_per_launch"] += (rec["dram_read_bytes"] or 0) + (rec["dram_write_bytes"] or 0)
        total["time_ms"] += rec["time_ms"] or 0
    with open(out_path, "w") as f:
        json.dump(total, f, indent=1)
    print(json.dumps(total))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "dominant":
        dominant_json(sys.argv[4:], sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2:])
