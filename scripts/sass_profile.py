"""Compact view of `ncu --page source --csv`: runs of SASS instructions with the same executed
count, their sample share and opcode mix.  usage: sass_profile.py file.csv [full]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
full = len(sys.argv) > 2
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) >= len(hdr) and r[0].startswith('0x')]
tot_s = sum(int(r[idx['# Samples']] or 0) for r in data)
tot_e = sum(int(r[idx['Instructions Executed']] or 0) for r in data)
print('instructions', len(data), 'executed', tot_e, 'samples', tot_s)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
runs = []
for n, r in enumerate(data):
    ex = int(r[idx['Instructions Executed']] or 0)
    s = int(r[idx['# Samples']] or 0)
    op = r[idx['Source']].split()[0] if not r[idx['Source']].strip().startswith('@') else r[idx['Source']].split()[1]
    if runs and runs[-1]['ex'] == ex:
        u = runs[-1]
    else:
        u = dict(ex=ex, n0=n, cnt=0, s=0, ops=collections.Counter(), st=collections.Counter()); runs.append(u)
    u['cnt'] += 1; u['s'] += s; u['ops'][op.split('.')[0]] += 1
    for h in stalls:
        u['st'][h[6:]] += int(r[idx[h]] or 0)
    if full:
        top = sorted([(int(r[idx[h]] or 0), h[6:]) for h in stalls], reverse=True)[:2]
        print(f"{n:5d} ex={ex:8d} s={s:5d} {r[idx['Source']].strip()[:90]:90s} {[t for t in top if t[0]]}")
if not full:
    for u in runs:
        if u['ex'] * u['cnt'] < tot_e * 0.003 and u['s'] < tot_s * 0.003:
            continue
        print(f"@{u['n0']:5d} n={u['cnt']:4d} ex={u['ex']:8d} instr%={100*u['ex']*u['cnt']/tot_e:5.1f} samp%={100*u['s']/max(tot_s,1):5.1f} "
              f"{dict(u['ops'].most_common(6))} {dict(u['st'].most_common(3))}")
