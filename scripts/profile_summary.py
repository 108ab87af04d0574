"""profiles/: raw CSV + summary JSON of the two kernels' ncu captures, and traffic.json.
Usage: python scripts/profile_summary.py gpurun_out/r1_split_blink.ncu-rep gpurun_out/r1_split_primary.ncu-rep UNITS"""
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
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
NAMES = {'blink': 'blink_kernel', 'primary': 'sweep_kernel (primary only)'}
CONV = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
units = int(sys.argv[3])
traffic = {}
for k, rep in (('blink', sys.argv[1]), ('primary', sys.argv[2])):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    open('profiles/r1_split_%s_kernel_full_raw.csv' % k, 'w').write(raw)
    rows = list(csv.reader(raw.splitlines()))
    d = dict((a, (b, c)) for a, b, c in zip(rows[0], rows[1], rows[2]))
    out = {'kernel': d['Kernel Name'][1]}
    for key in KEYS:
        if key in d:
            out[key] = {'unit': d[key][0], 'value': d[key][1]}
    for a, (b, c) in d.items():
        if a.startswith('smsp__average_warps_issue_stalled') and a.endswith('_per_issue_active.ratio'):
            out.setdefault('stalls_per_issue', {})[
                a.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')] = float(c)
    dram = sum(float(d[m][1]) * CONV[d[m][0]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    out['units_in_launch'] = units
    out['dram_bytes_per_unit'] = dram / units
    out['warp_instructions_per_unit'] = float(d['smsp__inst_executed.sum'][1]) / units
    traffic[NAMES[k]] = dram / units
    json.dump(out, open('profiles/r1_split_%s_kernel_summary.json' % k, 'w'), indent=1)
    print(k, d['gpu__time_duration.sum'], out['dram_bytes_per_unit'], out['warp_instructions_per_unit'],
            d['smsp__issue_active.avg.pct_of_peak_sustained_active'][1],
            d['smsp__thread_inst_executed_per_inst_executed.ratio'][1], d['launch__registers_per_thread'][1])
json.dump({'dram_bytes_per_unit_sweep': sum(traffic.values()), 'dram_bytes_per_unit': traffic,
    'source': 'profiles/r1_split_{blink,primary}_kernel_full_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum '
        'over the %d units of each captured launch, cold caches)' % units},
    open('profiles/traffic.json', 'w'), indent=1)
