"""Summarise an `ncu --csv --log-file` launch list: per kernel, per metric -> count / mean / min / max."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value')
d = defaultdict(list)
for r in rows[1:]:
    try: d[(r[ki][:72], r[mi])].append(float(r[vi].replace(',', '')))
    except ValueError: pass
for k, v in sorted(d.items()):
    print("%-72s %-56s n=%3d mean=%14.1f min=%14.1f max=%14.1f" % (k[0], k[1], len(v), sum(v) / len(v), min(v), max(v)))
