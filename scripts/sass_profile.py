"""Segment the per-instruction counts of an ncu report (--page source --csv) into loops:
   python scripts/sass_profile.py report.ncu-rep [min_share]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
iE = hdr.index('Instructions Executed'); iT = hdr.index('Avg. Threads Executed'); iS = hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
print('kernel', rows[0][1], 'total warp-instructions', tot, 'SASS lines', len(data))
segs = []; cur = None
for i, r in enumerate(data):
    n = int(r[iE])
    if cur is None or not (0.6 * cur['n'] <= n <= 1.67 * cur['n'] if cur['n'] > 0 else n == 0):
        cur = {'start': i, 'n': n, 'sum': 0, 'cnt': 0, 'thr': 0.0, 'samp': 0}
        segs.append(cur)
    cur['sum'] += n; cur['cnt'] += 1; cur['thr'] += float(r[iT]) * n; cur['samp'] += int(r[iS]); cur['end'] = i
tsamp = sum(s['samp'] for s in segs)
for s in segs:
    if s['sum'] > min_share * tot:
        print('%5d-%5d  ninst %4d  exec/inst %9d  share %5.1f%%  lanes %4.1f  samples %4.1f%%   %s' % (
            s['start'], s['end'], s['cnt'], s['sum'] // s['cnt'], 100 * s['sum'] / tot, s['thr'] / max(s['sum'], 1),
            100.0 * s['samp'] / max(tsamp, 1), data[s['start']][1].strip()[:50]))
if len(sys.argv) > 3:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, b):
        r = data[i]
        print('%4d %8d %5.1f %5s  %s' % (i, int(r[iE]), float(r[iT]), r[iS], r[1].strip()))
