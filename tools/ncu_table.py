"""Condenses an `ncu --page raw --csv` export into one line per launch (duration, DRAM bytes, throughputs).
Usage: python tools/ncu_table.py gpurun_out/<tag>_raw.csv [--agg]"""
import collections
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("l1tex__t_bytes.sum", "l1_MB"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "inst_M"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64_M"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf_M"),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    idx = [(H.index(c), n, U[H.index(c)]) for c, n in COLS if c in H]
    out = []
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "")
        vals = {}
        for i, n, unit in idx:
            v = to_float(r[i])
            if n == "dur_us":
                v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
            if n.endswith("_MB"):
                v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)
            if n.endswith("_M"):
                v = v / 1e6
            vals[n] = v
        out.append((name, vals))
    names = [n for _, n, _ in idx]
    if "--agg" in sys.argv:
        agg = collections.OrderedDict()
        for name, v in out:
            a = agg.setdefault(name, [0, collections.defaultdict(float)])
            a[0] += 1
            for k, x in v.items():
                a[1][k] += x
        print("%-28s %5s " % ("kernel", "n") + " ".join("%11s" % n for n in names))
        for name, (cnt, v) in agg.items():
            print("%-28s %5d " % (name[:28], cnt) + " ".join(
                "%11.2f" % (v[n] / cnt if n.endswith("%") or n == "regs" else v[n]) for n in names))
    else:
        print("%-28s " % "kernel" + " ".join("%11s" % n for n in names))
        for name, v in out:
            print("%-28s " % name[:28] + " ".join("%11.2f" % v[n] for n in names))


if __name__ == "__main__":
    main()
