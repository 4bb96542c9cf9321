#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (one row per kernel launch x metric) into one line per launch.

    python tools/ncu_list.py gpurun_out/launches.csv [first_id [last_id]]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
hdr, data, order = None, {}, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        key = (int(d["ID"]), d["Kernel Name"])
        if key not in data:
            data[key] = {}
            order.append(key)
        data[key][d["Metric Name"]] = d["Metric Value"]
short = {"gpu__time_duration.sum": "ns", "sm__inst_executed_pipe_fp64.sum": "fp64_warp_inst",
         "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_%", "launch__registers_per_thread": "regs",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_%", "launch__occupancy_limit_registers": "occ_regs"}
for k in order:
    if lo <= k[0] <= hi:
        print(k[0], k[1][:56], " ".join(f"{short.get(m, m)}={v}" for m, v in data[k].items()))
