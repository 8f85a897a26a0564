#!/usr/bin/env python
"""Warp instructions executed per CUDA source line of one kernel (ncu report captured with --import-source on, code built with -lineinfo).
usage: python tools/ncu_lines.py report.ncu-rep <launch id> [top n]"""
import csv
import io
import subprocess
import sys

rep, kid = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid + 1}"],
                     capture_output=True, text=True).stdout
cur_file, hdr, lines, total = None, None, [], 0
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ei, st = r.index("Instructions Executed"), r.index("Warp Stall Sampling (All Samples)")
    elif hdr and len(r) > 8 and r[0].isdigit():
        # the source text may contain commas the exporter does not quote: count the metric columns from the right
        off = len(r) - len(hdr)
        try:
            n, stl = int(r[ei + off] or 0), int(r[st + off] or 0)
        except ValueError:
            continue
        lines.append((n, stl, cur_file, int(r[0]), ",".join(r[1:2 + off]).strip()))
        total += n
print("total warp instructions", total)
for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{n:10d} {100 * n / total:5.1f}%  stalls {s:6d}  {f}:{ln}  {src[:110]}")
