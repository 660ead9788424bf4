#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel for the LAST training step."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))


def short(n):
    n = re.sub(r"ddl::\(anonymous namespace\)::", "", n)
    n = re.sub(r"void ", "", n)
    n = re.sub(r"\(.*", "", n)
    return n[:90]


# last step = launches after the last philox / since the 2nd-to-last softmax_xent
idx = [i for i, (n, _) in enumerate(rows) if "softmax_xent" in n]
if len(idx) >= 2:
    # a step spans from just after the previous step's last fused_sgd to this step's last fused_sgd
    sgd = [i for i, (n, _) in enumerate(rows) if "fused_sgd" in n or "fused_allreduce" in n]
    last_end = sgd[-1]
    prev_end = max(i for i in sgd if i < idx[-1])
    # previous step's last sgd:
    prev_sgds = [i for i in sgd if i < idx[-1]]
    start = prev_sgds[-1] + 1 if prev_sgds else 0
    step = rows[start:last_end + 1]
else:
    step = rows
tot = sum(ns for _, ns in step)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in step:
    k = short(n)
    agg[k][0] += 1
    agg[k][1] += ns
print(f"launches in step: {len(step)}, sum of kernel durations: {tot / 1e6:.3f} ms")
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% |")
