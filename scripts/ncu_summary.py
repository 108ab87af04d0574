"""Summary of one ncu report (first kernel in it): the counters the round's decisions are based on.
Usage: python scripts/ncu_summary.py report.ncu-rep UNITS [out.json] [raw.csv]"""
import csv, json, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active']
CONV = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1, 'Tbyte': 1e12}
rep = sys.argv[1]
units = int(sys.argv[2])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
if len(sys.argv) > 4:
    open(sys.argv[4], 'w').write(raw)
rows = list(csv.reader(raw.splitlines()))
d = dict((a, (b, c)) for a, b, c in zip(rows[0], rows[1], rows[2]))
out = {'kernel': d['Kernel Name'][1]}
for key in KEYS:
    if key in d:
        out[key] = {'unit': d[key][0], 'value': d[key][1]}
stalls = {}
for a, (b, c) in d.items():
    if a.startswith('smsp__average_warps_issue_stalled') and a.endswith('_per_issue_active.ratio'):
        stalls[a.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')] = float(c)
out['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
dram = sum(float(d[m][1]) * CONV[d[m][0]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
out['units_in_launch'] = units
out['dram_bytes_per_unit'] = dram / units
out['warp_instructions_per_unit'] = float(d['smsp__inst_executed.sum'][1].replace(',', '')) / units
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], 'w'), indent=1)
for k, v in out.items():
    if k == 'stalls_per_issue':
        print(k, ' '.join('%s %.2f' % (a, b) for a, b in list(v.items())[:9]))
    else:
        print(k, v if not isinstance(v, dict) else '%s %s' % (v['value'], v['unit']))
