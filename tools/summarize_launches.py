"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time
and share of the captured window.  Usage: python tools/summarize_launches.py launches.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if len(r) <= idx["Metric Value"] or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[idx["Kernel Name"]]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    v = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    agg[name][0] += 1
    agg[name][1] += us
    total += us
print(f"captured launches: {sum(a[0] for a in agg.values())}, total device time {total / 1e3:.2f} ms "
      f"(serialised, cold-cache under ncu: compare SHARES, not absolutes)")
print(f"{'share':>7s} {'ms':>10s} {'launches':>9s}  kernel")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * us / total:6.2f}% {us / 1e3:10.3f} {n:9d}  {name[:110]}")
mine = sum(us for n_, (c, us) in agg.items() if n_.startswith("clipa::"))
print(f"clipa_b200 kernels: {100 * mine / total:.1f}% of captured device time")
