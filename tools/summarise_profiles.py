#!/usr/bin/env python3
"""Collect the round's ncu output from gpurun_out/ into profiles/: copies the per-kernel details pages and the launch
list, writes r2_ncu_summary.csv (selected metrics of every --set full capture), r2_traffic.json (what bench.py reads
for roofline.traffic) and prints the per-kernel shares of the launch list.   python tools/summarise_profiles.py"""
import csv
import json
import os
import shutil
import statistics
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KERNELS = ["k_accumulate", "k_prepare", "k_scalars", "k_scatter", "k_ell2_maps", "k_scalar_mul_plan", "k_scalar_mul_proj", "k_dec_finish"]
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw(kernel):
    f = os.path.join(G, f"r2_ncu_{kernel}_raw.csv")
    if not os.path.exists(f):
        return None
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    table = {}
    for k in KERNELS:
        src = os.path.join(G, f"r2_ncu_{k}_details.csv")
        if os.path.exists(src):
            shutil.copy(src, os.path.join(P, f"r2_ncu_{k}_details.csv"))
        r = raw(k)
        if r:
            table[k] = r
    with open(os.path.join(P, "r2_ncu_summary.csv"), "w") as f:
        ks = [k for k in KERNELS if k in table]
        f.write("metric," + ",".join(ks) + "\n")
        for m in METRICS:
            f.write(m + "," + ",".join((table[k].get(m, ("", ""))[0] + " " + table[k].get(m, ("", ""))[1]).strip() for k in ks) + "\n")
    if "k_accumulate" in table:
        t = table["k_accumulate"]

        def num(m):
            v, u = t[m]
            x = float(v.replace(",", ""))
            return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        tr = {"k_accumulate": {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                               "additions_per_launch": 59243748,
                               "source": "ncu --set full --clock-control none, tools/dev_verify_once.py 20 2, second launch; "
                                         "profiles/r2_ncu_k_accumulate_details.csv"}}
        json.dump(tr, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
        print("traffic", tr)
    ll = os.path.join(G, "r2_launches_bench_py.csv")
    if os.path.exists(ll):
        shutil.copy(ll, os.path.join(P, "r2_launches_bench_py.csv"))
        per = {}
        for row in csv.reader(open(ll)):
            if len(row) > 10 and row[0].isdigit():
                name = row[4].split("(")[0].replace("void ", "").split("<")[0].replace("avrf::", "")
                try:
                    per.setdefault(name, []).append(float(row[-1].replace(",", "")))
                except ValueError:
                    pass
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            print(f"{name:28s} launches {len(v):5d}  median {statistics.median(v):10.1f}  sum {sum(v):12.1f}")


if __name__ == "__main__":
    main()
