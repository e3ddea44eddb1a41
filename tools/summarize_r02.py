#!/usr/bin/env python
"""Summarise the round-2 ncu artefacts (gpurun_out/, scratch) into profiles/ (tracked):
  launches_r02.csv      ncu --metrics gpu__time_duration.sum of tools/profile_step.py (one step)
  prof_step_r02.ncu-rep ncu --set full of the same command (every kernel of the step once)
-> profiles/launches_r02.csv (this library's kernels), profiles/ncu_summary_r02.json:
   per-kernel share of the step + the key raw metrics of each full capture."""
import collections
import csv
import io
import json
import os
import subprocess

OUT = {}
KEYS = {"ipm_quad": "ipm_quad_kernel", "knn_search2": "knn_search2_kernel", "cloud_index": "cloud_index_kernel",
        "group_boxes": "group_boxes_kernel", "cloud_compact": "cloud_compact_kernel", "pack_prefix": "pack_prefix_kernel",
        "replan": "replan_kernel", "solve_order_class": "solve_order_class_kernel", "solve_order_scatter": "solve_order_scatter_kernel"}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

lp = "gpurun_out/launches_r02.csv"
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr, agg, keep = None, collections.defaultdict(lambda: [0.0, 0]), []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            keep.append(r)
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"]
        if "ampc::" not in name and not any(k in name for k in ("pack_prefix", "replan_kernel", "fill_i32")):
            continue
        keep.append(r)
        short = name.split("(")[0].replace("void ", "").split("<")[0]
        t = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
        agg[short][0] += t_us
        agg[short][1] += 1
    tot = sum(v[0] for v in agg.values()) or 1.0
    OUT["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv "
                                     "python tools/profile_step.py   (one step of 32768 instances: index + round)",
                          "note": "per-launch times under ncu are cold-cache and serialised: compare SHARES",
                          "kernels": {k: {"us": v[0], "launches": v[1], "share": v[0] / tot} for k, v in
                                      sorted(agg.items(), key=lambda kv: -kv[1][0])}}
    csv.writer(open("profiles/launches_r02.csv", "w")).writerows(keep)

rp = "gpurun_out/prof_step_r02.ncu-rep"
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ci = {n: i for i, n in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ci["Kernel Name"]]
        for key, pat in KEYS.items():
            if pat in name and key not in OUT:
                OUT[key] = {"kernel": name.split("(")[0], "command": "ncu --set full --clock-control none --import-source on "
                            "--profile-from-start off python tools/profile_step.py",
                            "metrics": {m: {"value": r[ci[m]].replace(",", ""), "unit": units[ci[m]]} for m in WANT if m in ci}}
json.dump(OUT, open("profiles/ncu_summary_r02.json", "w"), indent=1)
print(json.dumps({k: (v.get("metrics", {}).get("gpu__time_duration.sum") if isinstance(v, dict) else None) for k, v in OUT.items()}, indent=1)[:1500])
if "launch_list" in OUT:
    for k, v in OUT["launch_list"]["kernels"].items():
        print(f"{k:40s} {v['us']:10.1f} us {v['launches']:3d} launches  share {v['share']:.3f}")
