"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel over the LAST step (the launches
between the last two occurrences of the marker kernel).  usage: launch_summary.py <csv> [marker substring]"""
import csv, collections, sys
path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "k_adam"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = []
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(",", ""))
    v = v / 1000 if row[ui] == "ns" else v * 1000 if row[ui] == "ms" else v
    rows.append((row[ki].split("(")[0], v))
idx = [i for i, (k, _) in enumerate(rows) if marker in k]
step = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in step)
print("launches %d  total %.1f us (serialised, cold caches: shares, not absolute times)" % (len(step), tot))
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-56s %4d %9.1f us %5.1f%%  %7.1f us each" % (k[:56], n, v, 100 * v / tot, v / n))
