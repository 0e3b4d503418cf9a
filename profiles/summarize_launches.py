#!/usr/bin/env python
"""Turns ncu launch lists (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`)
of `bench.py --only cN --only-step --no-graph` runs into profiles/r02_kernels.json, the table bench.py reads for
`roofline.traffic`, `frac_dram` and the choice of the dominant kernel.

  python profiles/summarize_launches.py c2=profiles/r02_launches_c2.csv:7:64 c3=...:CALLS:BATCH  [-o profiles/r02_kernels.json]

CALLS = how many times the step ran in that capture (warm-up included), BATCH = its per-GPU batch.  Only this
repository's kernels are kept (torch's input-generation kernels are dropped).  ncu times are cold-cache and serialised:
the per-kernel SHARES are what is compared with the live CUDA-event timing, not the absolutes."""
import csv
import json
import os
import sys


def parse(path):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = {}
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]].lower()
        name = r[col["Metric Name"]]
        if name.startswith("gpu__time_duration"):
            v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1e-3)
        elif "bytes" in name:
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        d = launches.setdefault(int(r[col["ID"]]), {"name": r[col["Kernel Name"]]})
        d[name] = v
    return [launches[k] for k in sorted(launches)]


def ours(name):
    return not (name.startswith("void at::") or name.startswith("at::") or "at::native" in name or "cub::" in name or "nccl" in name.lower())


def summarize(path, calls, batch):
    per = {}
    for l in parse(path):
        if not ours(l["name"]):
            continue
        k = per.setdefault(l["name"], {"name": l["name"], "launches": 0, "us_sum": 0.0, "bytes_sum": 0.0})
        k["launches"] += 1
        k["us_sum"] += l.get("gpu__time_duration.sum", 0.0)
        k["bytes_sum"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    kernels = []
    for k in per.values():
        kernels.append({"name": k["name"], "launches_per_step": k["launches"] / calls, "us": k["us_sum"] / k["launches"],
                        "dram_bytes": k["bytes_sum"] / k["launches"], "us_per_step": k["us_sum"] / calls,
                        "dram_bytes_per_step": k["bytes_sum"] / calls})
    # kernels that ran less than once per step are input preparation (target assignment, anchor tables): listed, not summed
    setup = [k for k in kernels if k["launches_per_step"] < 0.9]
    kernels = [k for k in kernels if k["launches_per_step"] >= 0.9]
    tot_us = sum(k["us_per_step"] for k in kernels) or 1.0
    for k in kernels:
        k["share"] = k["us_per_step"] / tot_us
    for k in setup:
        k["share"] = 0.0
    kernels.sort(key=lambda k: -k["us_per_step"])
    return {"source": os.path.basename(path), "calls": calls, "batch": batch, "us_per_step_sum": tot_us,
            "dram_bytes_per_step": sum(k["dram_bytes_per_step"] for k in kernels), "dominant": kernels[0] if kernels else None,
            "kernels": kernels, "setup_kernels_not_in_step": setup}


def main():
    out = "profiles/r02_kernels.json"
    args = sys.argv[1:]
    if "-o" in args:
        i = args.index("-o")
        out = args[i + 1]
        del args[i:i + 2]
    table = {}
    if os.path.exists(out):
        table = json.load(open(out))
    for a in args:
        cfg, rest = a.split("=", 1)
        path, calls, batch = rest.rsplit(":", 2)
        table[cfg] = summarize(path, int(calls), int(batch))
    json.dump(table, open(out, "w"), indent=1)
    for cfg, t in sorted(table.items()):
        print("%s  %s  sum %.1f us/step  dram %.1f MB/step" % (cfg, t["source"], t["us_per_step_sum"], t["dram_bytes_per_step"] / 1e6))
        for k in t["kernels"]:
            print("   %6.1f us x %.2f  %8.1f MB  share %.2f  %s" % (k["us"], k["launches_per_step"], k["dram_bytes"] / 1e6, k["share"], k["name"][:70]))


if __name__ == "__main__":
    main()
