#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of the shipped library, the instruction count and the counts of the mnemonics
that show which hardware paths it uses (tcgen05 = UTCIMMA/UTCHMMA/LDTM, TMA = UTMALDG, bulk copies = UBLKCP, mbarrier =
SYNCS, packed fp32 = FFMA2/FADD2/FMUL2, PDL = ACQBULK/...).  usage: python tools/sass_report.py [library] > profiles/sass_r2.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vulkansift_b200/lib/libvulkansift.so"
KEYS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "FADD", "FMUL", "IMAD", "IDP", "LDS", "STS", "LDG", "STG", "ATOMG", "RED", "MUFU", "BAR", "ACQBULK", "PREEXIT"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name)
        name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "")
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        hist[name][m.group(1)] += 1
print(f"# SASS mnemonic counts per kernel of {lib} (cuobjdump -sass, static counts; tools/sass_report.py)")
for k, h in hist.items():
    total = sum(h.values())
    marks = " ".join(f"{m}={h[m]}" for m in KEYS if h.get(m))
    print(f"{k}\n    {total} instructions: {marks}")
