"""usage: tools/launch_agg.py <ncu launch csv> <iterations in the file> : per-kernel totals of the last iteration + every launch above a threshold"""
import collections
import csv
import re
import sys

path, iters = sys.argv[1], int(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 40.0
lines = [l for l in open(path) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    seq.append((k, v, row["Grid Size"], row["Block Size"]))
n = len(seq) // iters
it = seq[-n:]
agg = collections.OrderedDict()
for k, v, g, b in it:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print(n, "launches per iteration (torch kernels included), total", round(sum(v for _, v, _, _ in it)), "us")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:22]:
    print(f"{k:70s} x{c:3d} {v:9.1f} us")
print()
for s in it:
    if s[1] > thr:
        print(f"{s[0]:60s} {s[1]:9.1f} us  {s[2]} {s[3]}")
