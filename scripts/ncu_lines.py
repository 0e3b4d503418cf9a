#!/usr/bin/env python
"""Aggregates the stall samples of an ncu report by CUDA source line (needs -lineinfo and --import-source on).
  python scripts/ncu_lines.py report.ncu-rep [top_n [kernel_regex]]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []   # optional: only launches of this kernel
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
agg = {}
cur = None
for r in rows:
    if not r:
        continue
    if "# Samples" in r:
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) // 2:
        continue
    d = dict(zip(hdr, r))
    # in the cuda,sass view a CUDA line row is followed by its SASS rows; CUDA rows carry "Line No"/"File" style first column
    first = r[0]
    src = d.get("Source", "")
    try:
        s = int(d.get("# Samples", "0") or 0)
        ie = int(d.get("Instructions Executed", "0") or 0)
    except ValueError:
        continue
    if first.startswith("0x"):
        continue  # SASS row (already included in its CUDA line's totals)
    key = (first, src.strip()[:120])
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += s
    a[1] += ie
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                a[2][k] = a[2].get(k, 0) + int(v or 0)
            except ValueError:
                pass
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot)
for (line, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%%  inst %8d  L%-5s %-46s | %s" % (a[0], 100.0 * a[0] / tot, a[1], line, ",".join("%s=%d" % (k[6:], v) for k, v in st if v), src))
