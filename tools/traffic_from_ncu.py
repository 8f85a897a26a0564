"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the octave-0 layer launches of the blur kernel, from an
`ncu --set full` report -> profiles/traffic_r2.json (the `roofline.traffic` of the bench line).
usage: python tools/traffic_from_ncu.py <report.ncu-rep> <out.json>"""
import csv
import json
import subprocess
import sys


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def mb(r, k):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    per = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        grid = r[idx["Grid Size"]]
        targs = name.split("<")[1].split(">")[0].split(",") if "<" in name else []
        is_layer = len(targs) >= 2 and targs[1].strip().lstrip("(int)") != "0"  # template arguments <R, KIND, H16>: KIND 0 = seed pass
        if "blur_pass_fast_kernel" in name and is_layer and ("2040" in grid or "(60, 34" in grid):  # octave 0 = 60 x 34 tiles
            per.append({"kernel": name.split("(")[0], "read": mb(r, "dram__bytes_read.sum"), "write": mb(r, "dram__bytes_write.sum"),
                        "us": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))})
    res = {"source": rep, "launches": per}
    if per:
        res["fast_octave0_layer_launch_dram_bytes"] = sum(p["read"] + p["write"] for p in per) / len(per)
        res["note"] = ("mean over the octave-0 layer launches captured; ncu runs a kernel with cold caches and reads the write counter before L2 has "
                       "written everything back, so the write part under-counts (66.4 MB are written per launch)")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "launches"}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
