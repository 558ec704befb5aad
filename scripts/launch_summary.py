"""Aggregate an ncu launch list (gpu__time_duration per launch) by kernel name -> profiles/*.txt"""
import collections
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
total = 0.0
for row in r:
    if len(row) <= iv:
        continue
    try:
        v = float(row[iv].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", row[ik]).replace("void ", "").replace("mmtg::", "").replace("<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += v
    total += v
steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out_lines = [f"# ncu launch list summary of {src} ({steps:g} train steps captured; cold-cache, serialised: compare SHARES)",
             f"# total {total / 1e3 / steps:.1f} us per step",
             f"{'share':>7} {'us/step':>10} {'n/step':>7} {'avg us':>8}  kernel"]
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out_lines.append(f"{100 * t / total:6.1f}% {t / 1e3 / steps:10.1f} {n / steps:7.1f} {t / n / 1e3:8.1f}  {k}")
open(out, "w").write("\n".join(out_lines) + "\n")
print("\n".join(out_lines[:32]))
