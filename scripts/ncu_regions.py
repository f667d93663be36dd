"""Stall samples of an ncu source page per code region: python scripts/ncu_regions.py src.csv kernel_index [bucket]"""
import csv, io, sys
lines = open(sys.argv[1]).read().splitlines()
k = int(sys.argv[2]); step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
st = starts[k]; en = starts[k + 1] if k + 1 < len(starts) else len(lines)
rows = list(csv.reader(io.StringIO("\n".join(lines[st:en]))))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
recs = []
for n, r in enumerate(rows[1:]):
    if len(r) < len(hdr): continue
    try: s = int(r[ix['# Samples']] or 0)
    except ValueError: continue
    recs.append((s, n, r))
tot = sum(s for s, _, _ in recs)
b = {}
for s, n, r in recs:
    d = b.setdefault(n // step, {})
    d['_'] = d.get('_', 0) + s
    for c in stall:
        v = int(r[ix[c]] or 0)
        if v: d[c[6:]] = d.get(c[6:], 0) + v
print('total', tot)
for kk in sorted(b):
    d = b[kk]
    if d['_'] >= tot * 0.004:
        top = sorted(((v, c) for c, v in d.items() if c != '_'), reverse=True)[:4]
        print(f"#{kk*step:5d}: {d['_']:6d} {100*d['_']/tot:5.1f}%  " + " ".join(f"{c}={v}" for v, c in top))
marks = ('LDTM', 'STTM', 'UTMASTG', 'ATOMG', 'BAR.SYNC', 'UCGABAR_WAIT', 'UTCHMMA', 'LDG.E.STRONG', 'MEMBAR', 'STG.E.128')
for s, n, r in recs:
    src = r[ix['Source']]
    if any(t in src for t in marks): print(n, src[:70], s)
