#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump per CUDA source line.

usage: [DIS_KERN=<mangled substring>] ncu_lines.py <prof.ncu-rep> <libampc.so> <kernel-substring> [top]
Joins ncu's per-SASS-instruction counters with `nvdisasm --print-line-info` of the same
kernel (instruction order is identical) and prints executed instructions and stall
samples per source line."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
# split per function
lines = []
cur_fn, cur_line, active = None, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        active = os.environ.get("DIS_KERN", kern) in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur_line)
if len(lines) != len(body):
    print(f"warning: {len(lines)} disassembled vs {len(body)} profiled instructions", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for ln, r in zip(lines, body):
    ex = int(r[ci["Instructions Executed"]] or 0)
    st = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    a = agg[ln]
    a[0] += ex
    a[1] += st
    a[2] += 1
    tot[0] += ex
    tot[1] += st
print(f"total executed warp-instructions {tot[0]:,}  stall samples {tot[1]:,}")
src_cache = {}
keyi = 0 if os.environ.get("BY_EXEC") else 1
for ln, (ex, st, n) in sorted(agg.items(), key=lambda kv: -kv[1][keyi])[:top]:
    text = ""
    if ln:
        for base in (os.path.dirname(os.path.abspath(lib)) + "/../csrc", "."):
            p = os.path.join(base, ln[0])
            if os.path.exists(p):
                src_cache.setdefault(p, open(p).read().splitlines())
                if ln[1] - 1 < len(src_cache[p]):
                    text = src_cache[p][ln[1] - 1].strip()[:90]
                break
    print(f"{str(ln):28s} sass={n:4d} exec={ex:13,} ({100*ex/max(tot[0],1):5.1f}%) stall={st:7,} ({100*st/max(tot[1],1):5.1f}%)  {text}")
