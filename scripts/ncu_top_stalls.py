"""Top stalled SASS instructions of an ncu report (source page): python scripts/ncu_top_stalls.py rep.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
recs = []
for n, r in enumerate(rows[1:]):
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    recs.append((s, n, r))
total = sum(s for s, _, _ in recs)
print(f"total samples {total}")
for s, n, r in sorted(recs, reverse=True)[:top]:
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    st = " ".join(f"{c}={v}" for v, c in st if v)
    print(f"{s:7d} {100.0 * s / max(total, 1):5.1f}%  #{n:5d} {r[ix['Source']][:90]:90s} | {st}")

# samples per 50-instruction bucket (roles of a warp-specialised kernel occupy disjoint address ranges)
if len(sys.argv) > 3:
    step = int(sys.argv[3])
    buckets = {}
    for s, n, r in recs:
        buckets[n // step] = buckets.get(n // step, 0) + s
    for b in sorted(buckets):
        if buckets[b]:
            print(f"  #{b * step:5d}-{b * step + step - 1:5d}: {buckets[b]:6d} {100.0 * buckets[b] / total:5.1f}%")
