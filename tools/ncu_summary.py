#!/usr/bin/env python
"""Summarises an ncu report (.ncu-rep, read here on the CPU box) or a launch-list CSV into the
text files kept under profiles/.  Usage:
   python tools/ncu_summary.py full gpurun_out/x.ncu-rep  > profiles/rNN_x_full.txt
   python tools/ncu_summary.py launches gpurun_out/x_launches.csv > profiles/rNN_x_launches.txt"""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
           "sm__cycles_elapsed.avg.per_second",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of {path} (per launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n== {d['Kernel Name'][:110]}  (launch id {d.get('ID','?')})")
        for m in METRICS:
            if m in d:
                print(f"   {m:78s} {d[m]:>16s} {units[hdr.index(m)]}")
        try:   # the two columns may carry different units (Gbyte / Mbyte / Kbyte)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            tr = sum(float(d[m]) * scale[units[hdr.index(m)]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            print(f"   {'traffic = dram read + write':78s} {tr / 1e9:16.4f} Gbyte")
        except Exception:
            pass


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:90]
        if "<" in r[4]:
            name = r[4][:r[4].index("(")][:90]
        ns = float(r[-1])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"# launch list {path}: {len(rows)} launches, {tot/1e6:.3f} ms total device time (gpu__time_duration.sum)")
    print(f"{'kernel':92s} {'launches':>8s} {'total ms':>10s} {'avg ms':>10s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {n:8d} {ns/1e6:10.3f} {ns/1e6/n:10.4f} {100*ns/tot:6.1f}%")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
