"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last N launches (one batched step) per kernel."""
import collections
import csv
import sys

f, n = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
body = rows[1:][-n:]
agg = collections.OrderedDict()
for r in body:
    k = r[ki].split("(")[0][:60]
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f, "last step: %.1f us over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("  %-62s n=%4d  %9.1f us  (%.1f us each)" % (k, a[0], a[1] / 1e3, a[1] / 1e3 / a[0]))
for r in body[2:11] + body[-3:]:
    print("    %-50s %8.1f us" % (r[ki].split("(")[0][-48:], float(r[vi].replace(",", "")) / 1e3))
