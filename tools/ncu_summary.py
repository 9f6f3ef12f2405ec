#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` capture exported with tools/ncu_capture.sh:
key counters per kernel from <tag>_raw.csv and the stall-reason split + hottest SASS lines from <tag>_source.csv.

  python tools/ncu_summary.py gpurun_out/r1d_particles128 > profiles/r1d_particles128_ncu_summary.txt
"""
import csv
import sys

csv.field_size_limit(10 ** 9)
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram written"), ("launch__registers_per_thread", "registers / thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("lts__t_sector_hit_rate.pct", "L2 hit rate")]


def num(s):
    try:
        return int(s.replace(",", ""))
    except ValueError:
        return 0


def main():
    base = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    raw = list(csv.reader(open(base + "_raw.csv")))
    H, U = raw[0], raw[1]
    names = [r[H.index("Kernel Name")].split("(")[0].replace("void ", "") for r in raw[2:]]
    print("kernels:", ", ".join(names))
    for key, label in KEYS:
        if key in H:
            i = H.index(key)
            print("%-30s %-16s %s" % (label, U[i], "  ".join("%14.6g" % float(r[i].replace(",", "")) for r in raw[2:])))
    rows = list(csv.reader(open(base + "_source.csv")))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    import re
    seen = set()
    for k, h0 in enumerate(heads):
        hi = heads[k + 1] - 1 if k + 1 < len(heads) else len(rows)
        Hs = rows[h0]
        body = [r for r in rows[h0 + 1:hi] if len(r) > 5]
        # newer ncu exports name every section (and may repeat a kernel: one section per captured launch)
        label = None
        if h0 > 0 and rows[h0 - 1] and rows[h0 - 1][0] == "Kernel Name":
            label = re.sub(r"\(int\)|\(bool\)", "", re.sub(r"\(GridDesc.*|\(LevelDev.*|\(TailLevels.*", "", rows[h0 - 1][1])).replace("void ", "")
            if (label, len(body)) in seen:
                continue
            seen.add((label, len(body)))
        si, ii, src = Hs.index("# Samples"), Hs.index("Instructions Executed"), Hs.index("Source")
        tot = sum(num(r[si]) for r in body) or 1
        toti = sum(num(r[ii]) for r in body) or 1
        stall = {}
        for i, h in enumerate(Hs):
            if h.startswith("stall_") and "Not Issued" not in h and h not in stall:
                stall[h] = sum(num(r[i]) for r in body if len(r) > i)
        ssum = sum(stall.values()) or 1
        print("\n== %s: %d SASS lines, %d stall samples, %.1f M warp instructions" %
              (label or (names[k] if k < len(names) else k), len(body), tot, toti / 1e6))
        print("stall reasons (%% of samples): " + ", ".join("%s %.1f" % (h[6:], 100.0 * v / ssum) for h, v in
              sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
        for n, r in sorted(enumerate(body), key=lambda nr: -num(nr[1][si]))[:top]:
            print("  line %5d  %5.2f%% of samples  %9.2f M executions  %s" %
                  (n, 100.0 * num(r[si]) / tot, num(r[ii]) / 1e6, r[src].strip()[:80]))


if __name__ == "__main__":
    main()
