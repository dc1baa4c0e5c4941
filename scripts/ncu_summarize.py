"""Summarise an ncu --csv launch list (gpu__time_duration.sum) per kernel name: launches, total, average, share."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # drop the first `skip` launches (warm-up)
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6)
    rows.append((r["Kernel Name"], us))
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    name = k.split("(")[0]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total")
print("| kernel | launches | avg us | total ms | share |\n|---|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| {k[:90]} | {n} | {us / n:.1f} | {us / 1e3:.3f} | {us / tot:.3f} |")
