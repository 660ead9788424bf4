#!/usr/bin/env python
"""cuobjdump -sass of the built module -> profiles/sass_summary.md (which kernels contain tcgen05/TMA/multimem SASS)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "distributeddeeplearning_b200", "_C.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
pats = [("UTCHMMA", r"UTCHMMA"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"UTMALDG"), ("UTCBAR", r"UTCBAR"), ("SYNCS", r"SYNCS"),
        ("LDGSTS", r"LDGSTS"), ("LDGMC (multimem.ld_reduce)", r"LDGMC"), ("STG.E.{64,128}.STRONG.SYS (multimem.st)", r"STG\.E\.(64|128)\.STRONG\.SYS"),
        ("REDG/ATOMG", r"\bREDG|\bATOMG")]
lines = ["# SASS evidence (`cuobjdump -sass distributeddeeplearning_b200/_C.so`, sm_100a)", "",
         "UTCHMMA = tcgen05.mma · LDTM = tcgen05.ld · UTMALDG = cp.async.bulk.tensor (TMA) · UTCBAR = tcgen05.commit · SYNCS = mbarrier ·",
         "LDGSTS = cp.async · LDGMC = multimem.ld_reduce (NVLS in-switch reduction) · multimem.st lowers to a system-scope vector store on the multicast address (STG.E.128.STRONG.SYS)", "",
         "| kernel | " + " | ".join(n for n, _ in pats) + " |", "|---|" + "---|" * len(pats)]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    m = re.search(r"\d+([a-z_0-9]+kernel)(I[A-Za-z0-9_]*E)?", name)
    disp = (m.group(1) + (m.group(2) or "")) if m else name[:60]
    counts = [len(re.findall(p, f)) for _, p in pats]
    if any(counts):
        lines.append("| " + disp + " | " + " | ".join(str(c) for c in counts) + " |")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "sass_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
print(f"... {len(lines) - 7} kernels listed")
