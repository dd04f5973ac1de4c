#!/usr/bin/env python3
"""Condenses `ncu --set full` captures of the step kernel into profiles/r01_step_kernel_ncu_summary.json
(read by bench.py for roofline.traffic).  usage: ncu_summary.py out.json name=capture.ncu-rep|capture_raw.csv[:note] ..."""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "sm__icc_request_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}

out = {}
for arg in sys.argv[2:]:
    name, rest = arg.split("=", 1)
    path, _, note = rest.partition(":")
    if path.endswith(".csv"):  # `ncu -i X.ncu-rep --page raw --csv` already exported on the GPU box
        txt = open(path, errors="replace").read()
    else:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h):
            try:
                m[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    dram = sum(m[k]["value"] * UNIT.get(m[k]["unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    out[name] = {"kernel": kname, "capture": note or path, "dram_bytes_per_launch": dram, "metrics": m}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print("wrote", sys.argv[1], {k: (v["kernel"], v["dram_bytes_per_launch"]) for k, v in out.items()})
