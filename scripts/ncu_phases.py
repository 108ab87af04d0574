"""Instructions executed / stall samples of an ncu report, per source file and per line range."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; h = None; recs = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        h = r; continue
    if h and len(r) == len(h) and r[0] != '':
        dd = dict(zip(h, r))
        try:
            recs.append((cur, int(r[0]), int(dd['# Samples']), int(dd['Instructions Executed']), int(dd.get('Thread Instructions Executed', 0) or 0)))
        except Exception:
            pass
ti = sum(x[3] for x in recs); ts = sum(x[2] for x in recs)
byfile = collections.defaultdict(lambda: [0, 0, 0])
for f, ln, s, i, t in recs:
    byfile[f][0] += s; byfile[f][1] += i; byfile[f][2] += t
print('total warp instr %d samples %d' % (ti, ts))
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print('%-28s instr %5.1f%% samples %5.1f%% lanes %.1f' % (f, 100.0 * v[1] / ti, 100.0 * v[0] / ts, v[2] / max(v[1], 1)))
if len(sys.argv) > 3:
    fname = sys.argv[2]
    bounds = [int(x) for x in sys.argv[3].split(',')]
    names = sys.argv[4].split(',') if len(sys.argv) > 4 else [str(b) for b in bounds]
    agg = [[0, 0, 0] for _ in bounds]
    for f, ln, s, i, t in recs:
        if f != fname: continue
        idx = -1
        for j, b in enumerate(bounds):
            if ln >= b: idx = j
        if idx >= 0:
            agg[idx][0] += s; agg[idx][1] += i; agg[idx][2] += t
    for n, b, v in zip(names, bounds, agg):
        print('  %-14s from line %4d: instr %5.1f%% samples %5.1f%% lanes %.1f' % (n, b, 100.0 * v[1] / ti, 100.0 * v[0] / ts, v[2] / max(v[1], 1)))
