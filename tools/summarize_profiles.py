#!/usr/bin/env python
"""Summarise ncu artefacts of a round into profiles/ (tracked): the launch list of the bench
command (per-kernel totals and shares) and the key raw metrics of each full capture.
usage: summarize_profiles.py <round tag, e.g. r01>"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1]
out = {}
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{tag}.csv")) if len(r) > 5]
hdr, agg, order = None, collections.defaultdict(list), []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") == "gpu__time_duration.sum":
        name = d["Kernel Name"]
        grid = d.get("Grid Size", "")
        agg[name].append((float(d["Metric Value"].replace(",", "")), grid))
OURS = ("ampc::", "pack_prefix_kernel", "replan_kernel", "best_of_kernel", "repack_kernel")
mine = {k: v for k, v in agg.items() if any(o in k for o in OURS)}
# bench.py order: [1 index launch at set-up] 3 warm-up steps, 3 TIMED steps, then e2e steps and the
# single-instance latency loop.  Timed region = launches W..W+2 among the full-batch launches.
steps, warm = 3, 3
sel = {}
for k, v in mine.items():
    full = [t for t, g in v if "1024" in g or g.startswith("(512,") or g.startswith("(8,")]  # B=1024 launches
    skip = warm + (1 if "cloud_index" in k or "cloud_compact" in k else 0)
    if len(full) >= skip + steps:
        sel[k] = full[skip:skip + steps]
tot = sum(sum(v) for v in sel.values())
lst = []
for k, t in sorted(sel.items(), key=lambda kv: -sum(kv[1])):
    lst.append({"kernel": k.split("(")[0], "launches_in_timed_region": len(t), "avg_us": sum(t) / len(t) / 1e3,
                "share_of_our_kernels": sum(t) / tot})
out["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 python bench.py --steps 3 --warmup 3 --no-cpu-baseline",
                      "note": "per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes",
                      "kernels": lst, "other_kernels_total_launches": sum(len(v) for k, v in agg.items() if "ampc::" not in k)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct"]
for k in ("cloud_index", "knn_search", "ipm_solve"):
    rep = f"gpurun_out/prof_{k}_{tag}.ncu-rep"
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, u, v = r[0], r[1], r[2]
    m = {}
    for i, n in enumerate(h):
        if n in want:
            m[n] = {"value": v[i], "unit": u[i]}
    out[k] = {"capture": f"ncu --set full --clock-control none --import-source on -k regex:{k} -s 4 -c 1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline", "metrics": m}
os.makedirs("profiles", exist_ok=True)
# the launch list itself, restricted to this library's kernels (torch data-generation kernels dropped)
with open(f"profiles/launches_{tag}.csv", "w", newline="") as fh:
    wr = csv.writer(fh)
    wr.writerow(hdr)
    for r in rows:
        if r[0] != "ID" and any(o in r[hdr.index("Kernel Name")] for o in OURS):
            wr.writerow(r)
json.dump(out, open(f"profiles/ncu_summary_{tag}.json", "w"), indent=1)
print(json.dumps(out["launch_list"]["kernels"], indent=1))
for k in ("cloud_index", "knn_search", "ipm_solve"):
    if k in out:
        print(k, {a: b["value"] + " " + b["unit"] for a, b in out[k]["metrics"].items()})
