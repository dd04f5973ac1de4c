#!/usr/bin/env python3
"""Condenses `ncu --metrics ... --csv` passes over steady-state step-kernel launches (tools/prof_steady.py) into
profiles/r02_step_kernel_ncu_summary.json: per-launch warp instructions, FP32 flops (fadd + fmul + 2 ffma thread
instructions), DRAM bytes, duration -- the measured numerators of bench.py's issue_slot_frac / fp32_frac / traffic.
usage: ncu_counts.py out.json task_batch=log.csv ..."""
import csv, json, sys
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}
out = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=", 1)
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    iid, ikern, imet, iunit, ival = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = {}
    for r in rows[1:]:
        try:
            launches.setdefault(r[iid], {"kernel": r[ikern]})[r[imet]] = (float(r[ival].replace(",", "")), r[iunit])
        except ValueError:
            pass  # 'n/a': the metric is not collectable in this pass
    L = list(launches.values())
    n = len(L)
    def mean(metric, scale=None):
        vals = []
        for l in L:
            if metric in l:
                v, u = l[metric]
                vals.append(v * (scale.get(u, 1.0) if scale else 1.0))
        return sum(vals) / len(vals) if vals else None
    fadd, fmul, ffma = (mean("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k) for k in ("fadd", "fmul", "ffma"))
    task, _, batch = name.rpartition("_")
    out[name] = {
        "kernel": L[0]["kernel"], "task": task, "batch": int(batch), "launches_averaged": n,
        "capture": "ncu --metrics (counts) --clock-control none over %d steady-state launches of tools/prof_steady.py %s %s" % (n, task, batch),
        "duration_ms_under_ncu": mean("gpu__time_duration.sum", TIME),
        "warp_inst_per_launch": mean("smsp__inst_executed.sum"),
        "fp32_flop_per_launch": (fadd + fmul + 2 * ffma) if None not in (fadd, fmul, ffma) else None,
        "fp32_thread_inst": {"fadd": fadd, "fmul": fmul, "ffma": ffma},
        "dram_bytes_per_launch": (mean("dram__bytes_read.sum", UNIT) or 0) + (mean("dram__bytes_write.sum", UNIT) or 0),
        "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "local_load_inst": mean("sass__inst_executed_local_loads"), "local_store_inst": mean("sass__inst_executed_local_stores"),
        "icc_hit_pct": mean("sm__icc_request_hit_rate.pct"), "warps_active_per_cycle": mean("sm__warps_active.avg.per_cycle_active"),
    }
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ("capture", "fp32_thread_inst")})
