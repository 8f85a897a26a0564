#!/usr/bin/env python
"""Dynamic opcode histogram of one kernel from an ncu report captured with --import-source on.
usage: python tools/ncu_ophist.py report.ncu-rep <launch id> [pixels]   (pixels: print warp instructions per 32 pixels)"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
pixels = float(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(kid) + 1}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]
si, ei, st = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
c, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    src = re.sub(r"^@!?U?P\w+\s+", "", r[si].strip())
    op = src.split()[0] if src else ""
    opb = op.split(".")[0]
    if op.startswith("IMAD.MOV") or op.startswith("IMAD.IADD") or op.startswith("IMAD.SHL"):
        opb = ".".join(op.split(".")[:2])
    n = int(r[ei])
    c[opb] += n
    tot += n
    samp[opb] += int(r[st] or 0)
print("total warp instructions", tot, ("= %.1f per pixel" % (tot / (pixels / 32))) if pixels else "")
for k, v in c.most_common(28):
    print(f"{k:14s} {v:10d} {100 * v / tot:5.1f}%  stall samples {samp[k]}")
