"""Executed-instruction view of one kernel of an .ncu-rep captured with --import-source on: totals per opcode and the
plateaus of the per-instruction execution counts (loops show up as runs of equal counts).
usage: python tools/sass_hot.py report.ncu-rep <kernel-name substring> [dump.txt]"""
import collections
import csv
import subprocess
import sys


def sections(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    return secs


def main():
    secs = [s for s in sections(sys.argv[1]) if sys.argv[2] in s["name"]]
    s = secs[-1]
    ix = {h: i for i, h in enumerate(s["hdr"])}
    data = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), r[ix["Avg. Predicated-On Threads Executed"]]) for r in s["rows"]]
    tot = sum(c for _, c, _ in data)
    print(s["name"][:100], "\nwarp instructions executed:", tot, " SASS lines:", len(data))
    h = collections.Counter()
    for src, c, _ in data:
        t = src.split()
        op = t[1] if t[0].startswith("@") else t[0]
        h[op.split(".")[0]] += c
    print("  ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in h.most_common(16)))
    start, acc, n, prev = 0, 0, 0, None
    for i, (src, c, _) in enumerate(data + [("", -10**9, "")]):
        if prev is not None and abs(c - prev) > 0.15 * max(c, prev, 1):
            if acc > 0.01 * tot:
                print("  lines %4d-%4d  count ~%9d  n=%4d  %5.1f%%   %s" % (start, i - 1, prev, n, 100 * acc / tot, data[start][0][:60]))
            start, acc, n = i, 0, 0
        acc += max(c, 0)
        n += 1
        prev = c
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join("%4d %9d %5s  %s" % (i, c, p, src) for i, (src, c, p) in enumerate(data)))


if __name__ == "__main__":
    main()
