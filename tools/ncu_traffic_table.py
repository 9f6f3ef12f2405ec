#!/usr/bin/env python
"""Per-kernel launch time and DRAM traffic from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv` log of bench.py (tools/gpu_ncu_r1d.sh).  Writes a small JSON keyed by kernel name
(template arguments kept, parameter lists dropped): launches seen, mean ms / read bytes / written bytes per launch.
bench.py reads that JSON for the `traffic` field of its roofline object.

  python tools/ncu_traffic_table.py gpurun_out/r1d_traffic_256.csv profiles/r1d_traffic_256.json
"""
import collections
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src) if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    ki, mi, vi, ui, ii = (head.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    launches = collections.OrderedDict()
    for row in rd:
        d = launches.setdefault(row[ii], {"kernel": re.sub(r"\(.*", "", row[ki]).replace("void ", "")})
        v, u = float(row[vi].replace(",", "")), row[ui]
        if row[mi].startswith("gpu__time_duration"):
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
        else:
            d[row[mi]] = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    agg = collections.OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["kernel"], {"launches": 0, "ms": 0.0, "read": 0.0, "written": 0.0})
        a["launches"] += 1
        a["ms"] += d.get("ms", 0.0)
        a["read"] += d.get("dram__bytes_read.sum", 0.0)
        a["written"] += d.get("dram__bytes_write.sum", 0.0)
    out = {"source": src, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none, python bench.py --steps 1 --warmup 1 (256^3, 130 M particles); per-launch means; "
           "times under ncu are serialised and cold-cache", "kernels": {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        n = a["launches"]
        out["kernels"][k] = {"launches": n, "ms_per_launch": a["ms"] / n, "dram_read_bytes_per_launch": a["read"] / n,
                             "dram_written_bytes_per_launch": a["written"] / n,
                             "dram_bytes_per_launch": (a["read"] + a["written"]) / n}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in list(out["kernels"].items())[:14]:
        print("%-28s n=%5d  %9.4f ms  %9.1f MB read  %9.1f MB written" % (k, v["launches"], v["ms_per_launch"],
              v["dram_read_bytes_per_launch"] / 1e6, v["dram_written_bytes_per_launch"] / 1e6))


if __name__ == "__main__":
    main()
