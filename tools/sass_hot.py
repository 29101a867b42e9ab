"""Hot SASS regions of an `ncu --page source --csv` dump: python tools/sass_hot.py src.csv [lo hi]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; isrc = hdr.index('Source'); ie = hdr.index('Instructions Executed'); iss = hdr.index('# Samples')
data = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
tot = sum(d[1] for d in data)
if len(sys.argv) > 3:
    for i in range(int(sys.argv[2]), int(sys.argv[3])):
        s, e, sm = data[i]
        print(f"{i:5d} {e/1e6:8.2f}M {sm:6d}  {s[:110]}")
    sys.exit()
print("total inst", tot, "sass lines", len(data), "samples", sum(d[2] for d in data))
segs = []; start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][1] - data[start][1]) > 0.02 * max(data[start][1], 1):
        segs.append((start, i, data[start][1])); start = i
for s, e, c in segs:
    w = (e - s) * c
    if w > 8e6:
        fp = sum(1 for i in range(s, e) if any(x in data[i][0] for x in ('DFMA', 'DADD', 'DMUL', 'DSETP')))
        samples = sum(data[i][2] for i in range(s, e))
        print(f"[{s:4d},{e:4d}) n={e-s:3d} exec={c/1e6:7.2f}M total={w/1e6:7.1f}M ({w/tot*100:4.1f}%) fp64={fp:2d} samples={samples:6d}  first: {data[s][0][:50]}")
