"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
Usage: python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    unit = r[ix["Metric Unit"]]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((name, us))
rows = rows[skip:]
agg = OrderedDict()
for name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short).replace("iefvad::<unnamed>::", "")
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache per-launch times)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {us / 1e3:10.3f} {100 * us / tot:6.1f}%")
