#!/usr/bin/env python
"""Summarise the ncu artefacts of a round into profiles/ (tracked).

usage: summarize_profiles.py <tag, e.g. r01b>

inputs (gpurun_out/, scratch):
  launches_<tag>.csv     ncu --metrics gpu__time_duration.sum --clock-control none --csv of
                         `python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline`
  prof_all_<tag>.ncu-rep ncu --set full --clock-control none --import-source on
                         --profile-from-start off -k regex:<our hot kernels> python tools/profile_all.py
outputs (profiles/):
  launches_<tag>.csv     the launch list restricted to this library's kernels
  ncu_summary_<tag>.json per-kernel totals/shares of the bench step + key raw metrics of each capture
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1]
out = {}
OURS = ("ampc::", "pack_prefix_kernel", "replan_kernel", "best_of_kernel", "repack_kernel")

# ---- launch list of the bench command -----------------------------------------------------
lpath = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
    hdr, seq = None, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            seq.append((d["Kernel Name"], d.get("Grid Size", ""), float(d["Metric Value"].replace(",", ""))))
    # bench.py --streams 1: set-up (1 index), 3 warm-up steps, 3 single-stream steps, 3 timed steps,
    # then the e2e / depth / latency legs.  A "step" here = the launches between two consecutive
    # full-batch index builds; steps 4..9 (after set-up + warm-up) are the measured ones.
    full = lambda g: g.startswith("(1024,")
    step_id, steps = -1, collections.defaultdict(lambda: collections.defaultdict(float))
    for name, grid, t in seq:
        if not any(o in name for o in OURS):
            continue
        short = name.split("(")[0].replace("void ", "")
        if "cloud_index_kernel" in short and full(grid):
            step_id += 1
        if step_id >= 0:
            steps[step_id][short] += t
    meas = [steps[i] for i in range(4, 10) if i in steps]
    agg = collections.defaultdict(float)
    for s in meas:
        for k, v in s.items():
            agg[k] += v
    tot = sum(agg.values()) or 1.0
    out["launch_list"] = {
        "command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python bench.py --steps 3 --warmup 3 --streams 1 --no-cpu-baseline",
        "note": "per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes; "
                "averages over the 6 measured steps (single-stream pass + timed pass)",
        "steps_averaged": len(meas),
        "kernels": [{"kernel": k, "avg_us_per_step": v / max(len(meas), 1) / 1e3, "share_of_our_kernels": v / tot}
                    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])],
        "launches_of_other_kernels": sum(1 for n, _, _ in seq if not any(o in n for o in OURS))}
    os.makedirs("profiles", exist_ok=True)
    with open(f"profiles/launches_{tag}.csv", "w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow(hdr)
        for r in rows:
            if r[0] != "ID" and any(o in r[hdr.index("Kernel Name")] for o in OURS):
                wr.writerow(r)

# ---- full captures ------------------------------------------------------------------------
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
# order of the captured launches in tools/profile_all.py (second pass of each part)
LABELS = ["cloud_index (C1: 1024 organised 50k-point clouds)", "knn_search (C1)", "ipm_solve (C1)",
          "depth_obstacle (1024 frames 200x250)", "edge_grad", "edge_cloud",
          "cloud_index (clouds built from depth, Obstacle)", "cloud_index (clouds built from depth, Edge)",
          "cloud_index (shuffled clouds, original order: NaN filter + scene box)", "cloud_sort (Morton bucketing)",
          "cloud_index (over the bucketed copy)", "knn_search (bucketed copy)"]
rep = f"gpurun_out/prof_all_{tag}.ncu-rep"
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u = r[0], r[1]
    caps = []
    for i, v in enumerate(r[2:]):
        if len(v) != len(h):
            continue
        m = {n: {"value": v[j], "unit": u[j]} for j, n in enumerate(h) if n in want}
        caps.append({"label": LABELS[len(caps)] if len(caps) < len(LABELS) else "extra",
                     "kernel": v[h.index("Kernel Name")], "metrics": m})
    out["captures"] = {"command": "ncu --set full --clock-control none --import-source on --profile-from-start off "
                                  "-k regex:'cloud_index_kernel|knn_search_kernel|ipm_solve_kernel|cloud_sort_kernel|"
                                  "depth_obstacle_kernel|edge_grad_kernel|edge_cloud_kernel' python tools/profile_all.py",
                       "results": caps}
    # bench.py reads this key for roofline.traffic
    for c in caps:
        if c["label"].startswith("cloud_index (C1"):
            out["cloud_index"] = {"capture": out["captures"]["command"], "metrics": c["metrics"]}
        if c["label"].startswith("knn_search (C1"):
            out["knn_search"] = {"capture": out["captures"]["command"], "metrics": c["metrics"]}
        if c["label"].startswith("ipm_solve"):
            out["ipm_solve"] = {"capture": out["captures"]["command"], "metrics": c["metrics"]}
json.dump(out, open(f"profiles/ncu_summary_{tag}.json", "w"), indent=1)
if "launch_list" in out:
    print(json.dumps(out["launch_list"]["kernels"], indent=1))
for c in out.get("captures", {}).get("results", []):
    mm = c["metrics"]
    g = lambda k: (mm.get(k, {}).get("value", "?") + " " + mm.get(k, {}).get("unit", ""))
    print(f'{c["label"]:70s} {c["kernel"][:28]:28s} t={g("gpu__time_duration.sum")} rd={g("dram__bytes_read.sum")} '
          f'wr={g("dram__bytes_write.sum")} dram={g("dram__bytes.sum.per_second")} regs={g("launch__registers_per_thread")}')
