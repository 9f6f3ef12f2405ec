"""Top SASS instructions by stall samples from an `ncu --page source --csv` export.
Usage: python tools/ncu_hot.py <source.csv> <section-index> [top]"""
import collections
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sec = int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lo = heads[sec] + 1
hi = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
H = rows[heads[sec]]
si, ii, src = H.index("# Samples"), H.index("Instructions Executed"), H.index("Source")
body = [r for r in rows[lo:hi] if len(r) > si and r[si].replace(",", "").isdigit()]
tot = sum(int(r[si].replace(",", "")) for r in body)
toti = sum(int(r[ii].replace(",", "")) for r in body)
print("section %d: %d SASS lines, %d samples, %.1f M warp-instructions" % (sec, len(body), tot, toti / 1e6))
byop = collections.defaultdict(lambda: [0, 0])
for r in body:
    op = r[src].split()[0] if not r[src].strip().startswith("@") else r[src].split()[1]
    op = op.split(".")[0]
    byop[op][0] += int(r[si].replace(",", ""))
    byop[op][1] += int(r[ii].replace(",", ""))
print("by opcode (samples%, inst%):", ", ".join("%s %.1f/%.1f" % (k, 100 * v[0] / tot, 100 * v[1] / toti)
      for k, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]))
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][si].replace(",", "")))[:top]:
    print("%5d %6.2f%% inst %9.1fM  %s" % (n, 100 * int(r[si].replace(",", "")) / tot, int(r[ii].replace(",", "")) / 1e6,
                                          r[src].strip()[:90]))
