"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:60]
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1.0)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':62s} {'n':>4s} {'total ms':>10s} {'mean ms':>9s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:62s} {len(v):4d} {sum(v):10.3f} {sum(v)/len(v):9.3f} {100*sum(v)/tot:5.1f}%")
