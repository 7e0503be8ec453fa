#!/usr/bin/env python3
"""Stall-reason and opcode summary of one kernel from an `ncu --set full --import-source on` report
(`ncu -i <rep> --page source --csv`, SASS view).   python tools/ncu_source_summary.py <capture.ncu-rep>"""
import csv
import io
import subprocess
import sys
from collections import Counter

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
lines = out.splitlines()
print(lines[0].strip().strip(",").replace('"', ""))
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
ops, execd, stalls, samples = Counter(), 0, Counter(), 0
top = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    n = int(r[col["Instructions Executed"]] or 0)
    ops[op.split(".")[0]] += n
    execd += n
    s = int(r[col["# Samples"]] or 0)
    samples += s
    for h in stall_cols:
        stalls[h] += int(r[col[h]] or 0)
    top.append((s, r[col["Address"]][-5:], src[:70]))
print("static instructions %d, executed warp instructions %d, samples %d" % (len(rows) - 1, execd, samples))
print("opcodes (share of executed warp instructions):")
for op, n in ops.most_common(12):
    print("  %-10s %5.1f %%" % (op, 100.0 * n / max(execd, 1)))
tot = sum(stalls.values())
print("stall samples by reason:")
for h, n in stalls.most_common(10):
    print("  %-28s %5.1f %%" % (h, 100.0 * n / max(tot, 1)))
print("instructions with the most samples:")
for s, a, src in sorted(top, reverse=True)[:12]:
    print("  %6d (%4.1f %%)  ...%s  %s" % (s, 100.0 * s / max(samples, 1), a, src))
