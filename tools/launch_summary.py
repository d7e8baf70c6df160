#!/usr/bin/env python
"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into a markdown table of kernel shares:

    python tools/launch_summary.py gpurun_out/r2k_launches_step707.csv "title" "command" > profiles/r2_launches_step707.md"""
import csv
import re
import sys
from collections import defaultdict

path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
n = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[unit]
    name = re.sub(r"^(void )?(tsl::)?", "", r["Kernel Name"])
    name = re.sub(r"\(.*$", "", name)
    tot[name][0] += 1
    tot[name][1] += us
    n += 1
total = sum(v[1] for v in tot.values())
print(f"# {title}\n\nCommand: `{cmd}`\n(cold-cache, serialised under the profiler: compare SHARES, not absolute times).\nWindow: {n} launches, {total / 1e3:.2f} ms of kernel time.\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if t / total < 0.001:
        continue
    print(f"| `{k}` | {c} | {t:.0f} | {t / c:.1f} | {100 * t / total:.1f}% |")
