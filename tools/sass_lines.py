"""ncu source page (`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`) -> one line per SASS instruction:
offset, warp-level executions, threads per execution, stall samples, source file:line, instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'File Path'] + [len(rows)]
ins = {}
for k in range(len(secs) - 1):
    s = secs[k]
    hdr = rows[s + 2]
    ie, ith, isamp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    fn = rows[s][1].split('/')[-1].replace('ballenv_', '').split('.')[0]
    cur = None
    for r in rows[s + 3:secs[k + 1]]:
        if not r:
            continue
        if r[0] != '':
            cur = fn + ':' + r[0]
            continue
        if len(r) > ie and r[2].startswith('0x'):
            a = int(r[2], 16)
            if a not in ins:
                ins[a] = (cur, r[3], int(r[ie] or 0), int(r[ith] or 0), int(r[isamp] or 0))
addrs = sorted(ins)
print("# %d instructions, %d executed" % (len(addrs), sum(ins[a][2] for a in addrs)))
for a in addrs:
    c = ins[a]
    print("%6x %9d %5.1f %5d  %-16s %s" % (a - addrs[0], c[2], c[3] / max(c[2], 1), c[4], c[0], c[1]))
