"""Instructions executed / stall samples per CUDA source line from `ncu --page source --csv --print-source sass,cuda`.
usage: lines_profile.py <csv> <warp-steps in the capture> [top n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
den = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cur = None; hdr = None; out = []
for r in rows:
    if r and r[0] in ('File Path', 'File Name'):
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
        n = int(r[ie]) if r[ie].isdigit() else 0
        s = int(r[isamp]) if r[isamp].isdigit() else 0
        if n > 0: out.append((n, s, cur, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print('instructions per warp-step %.1f, samples %d' % (tot / den, ts))
for o in sorted(out, reverse=True)[:top]:
    print('%7.1f %5.1f%% %-20s %4d %s' % (o[0] / den, 100.0 * o[1] / max(ts, 1), o[2], o[3], o[4]))
