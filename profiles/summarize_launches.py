"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
usage: python profiles/summarize_launches.py profiles/launches_r01.csv"""
import collections
import csv
import io
import sys

lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.3f} ms {100 * v[1] / tot:5.1f}% {v[0]:5d}x  {k[:100]}")
