"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (second half = the timed step)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
n = len(rows)
half = rows[n // 2:] if "--all" not in sys.argv else rows
if "--last" in sys.argv:      # the last N launches (N = launches/step printed by quick_time.py: the first step also packs and initialises)
    half = rows[n - int(sys.argv[sys.argv.index("--last") + 1]):]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for row in half:
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v
    a = agg[name]
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(v[1] for v in agg.values())
print(f"launches in file {n}; timed step = last {len(half)}; sum of kernel durations {tot:.2f} ms")
print("| kernel | launches | ms | share | max ms |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[2]:.3f} |")
