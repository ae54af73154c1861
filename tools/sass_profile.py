"""Per-SASS-instruction execution profile from an ncu report: ncu -i X.ncu-rep --page source --csv | this."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
ith = hdr.index('Avg. Threads Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) <= ie or r[0] in ('Address', 'Kernel Name'): break
    try: data.append((int(r[ie]), int(r[isamp]), r[isrc].strip(), r[ith]))
    except ValueError: pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print("sass instructions", len(data), "executed", tot, "samples", ts)
if chunk > 0:
    for i in range(0, len(data), chunk):
        ch = data[i:i + chunk]
        c = sum(x[0] for x in ch)
        if c: print("%5d %9d %5.1f%% samples %5.1f%%  %s" % (i, c, 100 * c / tot, 100 * sum(x[1] for x in ch) / ts, ch[0][2][:60]))
else:
    for i, d in enumerate(data):
        if d[0] >= -chunk: print(i, d[0], d[3], d[1], d[2])
