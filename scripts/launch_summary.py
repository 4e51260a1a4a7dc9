"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares of the step)."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    k = r[ix["Kernel Name"]][:80]
    agg[k][0] += 1
    agg[k][1] += float(r[ix["Metric Value"]].replace(",", ""))
tot = sum(v for _, v in agg.values())
print(f"# {sys.argv[1]}: {sum(n for n, _ in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:82s} n={n:3d} {v / 1e3:12.1f} us {100 * v / tot:6.2f}%")
