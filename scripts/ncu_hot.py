"""Reads `ncu --page source --csv` output and prints the hottest SASS instructions with
their dominant stall reasons.  usage: ncu_hot.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print('kernel', rows[0][1][:80] if rows[0] else '', 'total samples', tot, 'ninstr', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for n, r in enumerate(data):
    s = int(r[idx['# Samples']] or 0)
    ex = int(r[idx['Instructions Executed']] or 0)
    if s > tot * minpct / 100:
        top = sorted([(int(r[idx[h]] or 0), h[6:]) for h in stalls], reverse=True)[:2]
        print(f"{n:5d} {s:6d} {100*s/tot:5.1f}% ex={ex:7d} {r[idx['Source']].strip()[:64]:64s} {top}")
