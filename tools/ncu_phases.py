"""Debug helper: split the ncu source page of a kernel into barrier-separated phases and print, per phase, its share of
the stall samples, its FP64 instruction count and the top stall reasons.
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_phases.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, x in enumerate(rows) if 'Source' in x and 'Address' in x][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [x for x in rows[hi + 1:] if len(x) >= len(hdr) - 3]
base = int(data[0][ix['Address']], 16)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
ph = []; cur = []
for x in data:
    cur.append(x)
    if 'BAR.' in x[ix['Source']]:
        ph.append(cur); cur = []
ph.append(cur)
tot = sum(int(x[ix['# Samples']]) for x in data)
print(rows[0][1][:140] if hi > 0 else '', 'total samples', tot)
for i, p in enumerate(ph):
    s = sum(int(x[ix['# Samples']]) for x in p)
    if s < tot * 0.008:
        continue
    agg = {k: sum(int(x[ix[k]]) for x in p) for k in stalls}
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    nd = sum(int(x[ix['Instructions Executed']]) for x in p if x[ix['Source']].split()[0] in ('DFMA', 'DMUL', 'DADD') or (len(x[ix['Source']].split()) > 1 and x[ix['Source']].split()[1] in ('DFMA', 'DMUL', 'DADD')))
    bar = p[-1][ix['Source']].strip()[:28]
    print(f"{i:3d} @{int(p[0][ix['Address']], 16) - base:6x} {s / tot:6.1%}  FP64 {nd:9d}  {[(k[6:], v) for k, v in top]}  ends: {bar}")
