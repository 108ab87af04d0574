"""Summarise an ncu report: headline metrics, stall mix, hottest source lines."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
for k in ('gpu__time_duration.sum', 'launch__block_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):
    print('%-64s %s' % (k, d.get(k)))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; h = None; recs = []; stall = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No':
        h = r; continue
    if h and len(r) == len(h):
        dd = dict(zip(h, r))
        if r[0] != '':
            try:
                recs.append((cur, int(r[0]), r[1].strip(), int(dd['# Samples']), int(dd['Instructions Executed'])))
            except Exception:
                pass
        else:
            for kk, v in dd.items():
                if kk.startswith('stall_') and 'Not Issued' not in kk:
                    try: stall[kk] += int(v)
                    except Exception: pass
ti = sum(x[4] for x in recs); ts = sum(x[3] for x in recs)
s = sum(stall.values()) or 1
print('stalls:', [(k, round(100.0 * v / s, 1)) for k, v in stall.most_common(8)])
print('total instr', ti, 'samples', ts)
print('--- hottest lines by stall samples')
for x in sorted(recs, key=lambda x: -x[3])[:top]:
    print('%-18s %4d %5.1f%% samp %5.1f%% instr | %s' % (x[0], x[1], 100.0 * x[3] / ts, 100.0 * x[4] / ti, x[2][:100]))
