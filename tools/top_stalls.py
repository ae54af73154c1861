"""Top stall-sampled SASS instructions of the first kernel in an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]; end = his[1] if len(his) > 1 else len(rows)
hdr = rows[hi]; ie = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(rows[hi + 1:end]):
    try: data.append((int(r[isamp]), int(r[ie]), k, r[isrc].strip(), r))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data)
lo, hi_e = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 60)
print("total samples", tot)
n = 0
for d in sorted(data, reverse=True):
    if not (lo <= d[1] <= hi_e): continue
    ex = sorted(((int(d[4][i]), h) for i, h in cols if d[4][i].isdigit() and int(d[4][i]) > 0), reverse=True)
    print("%5.1f%% exec=%9d #%5d %-46s %s" % (100 * d[0] / tot, d[1], d[2], d[3][:46], ex[:2]))
    n += 1
    if n >= int(sys.argv[4]) if len(sys.argv) > 4 else n >= 30: break
