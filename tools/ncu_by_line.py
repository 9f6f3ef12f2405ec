#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an `ncu --page source --csv` export to CUDA source lines, using the
line table of the object the kernel was built from (nvdisasm --print-line-info; inlined code is attributed to the
innermost line).  The n-th SASS instruction of the kernel in the export is the n-th instruction of the function in
the disassembly (same build), so no address arithmetic is needed.

  python tools/ncu_by_line.py <source.csv[.gz]> <section> <object.o> <mangled-function-substring> [top]
"""
import collections
import csv
import gzip
import os
import re
import subprocess
import sys
import tempfile

csv.field_size_limit(10 ** 9)


def num(s):
    try:
        return int(s.replace(",", ""))
    except ValueError:
        return 0


def main():
    src, sec, obj, fun = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
    op = gzip.open if src.endswith(".gz") else open
    rows = list(csv.reader(op(src, "rt")))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    h0 = heads[sec]
    hi = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
    H = rows[h0]
    body = [r for r in rows[h0 + 1:hi] if len(r) > 5]
    si, ii = H.index("# Samples"), H.index("Instructions Executed")
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, infun = [], ("?", 0), False
    for l in dis.splitlines():
        if l.startswith(".text."):
            infun = fun in l
            continue
        if not infun:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    if len(lines) != len(body):
        print("warning: %d instructions in the disassembly, %d in the export" % (len(lines), len(body)))
    stall_cols = {}
    for i, h in enumerate(H):
        if h.startswith("stall_") and "Not Issued" not in h and h not in stall_cols:
            stall_cols[h] = i
    agg = collections.defaultdict(lambda: [0, 0])
    why = collections.defaultdict(lambda: collections.Counter())
    for k, r in enumerate(body[:len(lines)]):
        a = agg[lines[k]]
        a[0] += num(r[si])
        a[1] += num(r[ii])
        for h, i in stall_cols.items():
            if i < len(r):
                why[lines[k]][h[6:]] += num(r[i])
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    cache = {}
    print("%-28s %8s %8s  %-34s" % ("source line", "samples", "instrs", "top stall reasons of the line"))
    for (f, n), (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        text = ""
        for root in ("libfluid_b200/csrc", "."):
            path = os.path.join(root, f)
            if os.path.exists(path):
                if path not in cache:
                    cache[path] = open(path).read().splitlines()
                if 0 < n <= len(cache[path]):
                    text = cache[path][n - 1].strip()[:70]
                break
        w = why[(f, n)]
        tot = sum(w.values()) or 1
        reasons = " ".join("%s:%d%%" % (k, round(100.0 * v / tot)) for k, v in w.most_common(2))
        print("%-20s:%-5d %7.2f%% %7.2f%%  %-34s %s" % (f, n, 100.0 * s_ / ts, 100.0 * i_ / ti, reasons, text[:60]))
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, n), (s_, i_) in agg.items():
        byfile[f][0] += s_
        byfile[f][1] += i_
    print("by file:", ", ".join("%s %.1f%% / %.1f%%" % (f, 100.0 * v[0] / ts, 100.0 * v[1] / ti) for f, v in byfile.items()))


if __name__ == "__main__":
    main()
