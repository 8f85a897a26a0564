"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the launch sequence."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        out.append((row["Kernel Name"].split("(")[0], v, row.get("Grid Size", ""), row.get("Block Size", "")))
    return out


if __name__ == "__main__":
    rows = load(sys.argv[1])
    agg = collections.defaultdict(list)
    for n, v, g, b in rows:
        agg[n].append(v)
    tot = sum(v for _, v, _, _ in rows)
    print("%-36s %5s %11s %9s %6s" % ("kernel", "n", "total_us", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-36s %5d %11.1f %9.2f %5.1f%%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    if len(sys.argv) > 2:
        for n, v, g, b in rows[int(sys.argv[2]):int(sys.argv[3])]:
            print("%-36s %10.2f %s %s" % (n, v, g, b))
