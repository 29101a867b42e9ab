"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
    a = agg.setdefault(r[ki][:64], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':66s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:66s} {a[0]:4d} {a[1]:10.1f} {a[1]/a[0]:9.2f} {a[1]/tot*100:5.1f}%")
