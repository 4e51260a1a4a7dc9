"""Executed warp-instructions and stall samples of one kernel per SOURCE region, from an ncu report's SASS page joined with
the line table of the library that was profiled (nvdisasm -g; needs -lineinfo).  Dev tool.
Usage: python scripts/ncu_source_regions.py <report.ncu-rep> <lib.so> <kernel-substring> [lines-per-bucket]"""
import collections
import csv
import io
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import sass_cost as SC

rep, lib, kern = sys.argv[1:4]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
ins = [i for i in SC.parse(SC.disasm(lib, kern)) if "label" not in i]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if r and r[0] == "Address")
data = rows[rows.index(hdr) + 1:]
assert len(data) == len(ins), (len(data), len(ins), "the library is not the build that was profiled")
ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, 0.0])
cls = collections.defaultdict(lambda: [0, 0])
for r, i in zip(data, ins):
    n, th, s = int(r[ia]), float(r[it]), int(r[isamp])
    f, ln = i["line"] or ("?", 0)
    k = f"{f}:{ln // bucket * bucket}"
    agg[k][0] += n; agg[k][1] += s; agg[k][2] += n * th
    c = "1 lane" if th <= 1.5 else "2-8 lanes" if th <= 8.5 else "9-32 lanes"
    cls[c][0] += n; cls[c][1] += s
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{kern}: {ti:.4g} warp-instructions, {ts} stall samples")
for c, (n, s) in sorted(cls.items()):
    print(f"  {c:12s} instructions {100 * n / ti:5.1f} %   samples {100 * s / ts:5.1f} %")
print(f"{'region':32s} {'inst %':>7s} {'samples %':>9s} {'avg lanes':>9s}")
for k, (n, s, nt) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:32s} {100 * n / ti:7.1f} {100 * s / ts:9.1f} {nt / max(n, 1):9.1f}")
