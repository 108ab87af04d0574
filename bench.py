#!/usr/bin/env python
"""Headline benchmark: Rao-Teh site-sweeps/sec at the p53 codon shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one full Gibbs sweep (primary track + all 20 tolerance tracks,
statistics accumulated) of every (site, chain) unit resident on the GPU(s):
BASELINE.json configs[2] -- the p53-shaped codon blink model (49-node tree, 61
codon states, 20 tolerance classes), synthetic alignment, 15 625 sites x 64
chains = 10^6 units per GPU.  One process per GPU; units are sharded by site with
no data-path collective ("scaling": "weak": every GPU holds 10^6 units); the one
NCCL all-reduce of the statistics happens after the timed region, as it would once
per EM iteration.

Printed JSON (one line, rank 0): see the contract in the task description.  In
short: `value` = whole-job site-sweeps/s with everything resident in HBM, timed
with CUDA events on the launching stream (max over ranks); `e2e` = the same metric
through the public host-buffer API with the step's site data uploaded from pinned
host memory and the statistics read back inside the timed region; `roofline` =
algorithmic HBM bytes of the sweep kernel / its measured duration against the
measured copy bandwidth in MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle
port (oracle/blink_oracle.py, a restatement of the reference's own Python path)
timed on this box's host cores on a bounded sample.

`--impl reference` times that CPU path alone (the reference is pure Python whose
FFBS lives in an un-vendored dependency, so oracle/_ref cannot exist; the oracle
port is the reference arm -- kind "port").
"""
from __future__ import division, print_function

import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'rao_teh_site_sweeps_per_sec_p53'
UNIT = 'site-sweeps/s'
WORKLOAD = 'p53-shaped codon blink model: 49-node tree, 61 states, 20 tolerance classes'


def p53_problem(n_sites, seed=0):
    from nxblink_b200 import p53, packing
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, n_sites, seed=seed, reference_leaf=leaf)
    return m, pm, ls, tt, tf


# ---------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Python path
# ---------------------------------------------------------------------------

def _cpu_worker(args):
    site, burn, sweeps = args
    from oracle import blink_oracle as bo
    from oracle.refcompat.make_golden import MaskData
    m, pm, ls, tt, tf = p53_problem(max(site + 1, 1), seed=0)
    data = MaskData(pm, ls[site], tt[site], tf[site])
    rng = np.random.default_rng(1000 + site)
    tb = bo.build_tables(m)
    pri, tols = bo.make_tracks(tb, data)
    bo._pool_data_for_zero_blen(tb, [pri] + tols)
    bo._init_tracks(rng, tb, pri, tols)
    summary = bo.Summary(tb.T, tb.root, tb.node_to_tm, tb.primary_to_tol, m.get_Q_primary())
    for _ in range(burn):
        bo._sweep(rng, tb, pri, tols)
    t0 = time.perf_counter()
    for _ in range(sweeps):
        bo._sweep(rng, tb, pri, tols)
        summary.on_sample(pri, tols)
    return time.perf_counter() - t0


def host_cores():
    """Host threads this process may use (affinity / cgroup aware), at most 128 workers."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 128))


def cpu_site_sweeps_per_sec(cores, burn, sweeps, pool=None):
    """All `cores` workers sweep one p53-shaped site each; aggregate site-sweeps/s."""
    own = pool is None
    if own:
        pool = multiprocessing.get_context('spawn').Pool(cores)
    try:
        t0 = time.perf_counter()
        elapsed = pool.map(_cpu_worker, [(i, burn, sweeps) for i in range(cores)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    # timed region = the sweeps themselves (model build / initialisation excluded);
    # all workers run concurrently, so the job takes as long as the slowest one
    return cores * sweeps / max(elapsed), wall


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    burn, sweeps = 1, 4
    pool = multiprocessing.get_context('spawn').Pool(cores)
    try:
        for _ in range(args.warmup):
            cpu_site_sweeps_per_sec(cores, 0, 1, pool)
        t0 = time.perf_counter()
        rates = [cpu_site_sweeps_per_sec(cores, burn, sweeps, pool)[0] for _ in range(args.steps)]
        wall = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = float(np.mean(rates))
    sample = ('%d host processes x 1 p53-shaped site x 1 chain x %d sweeps per step '
            '(after %d burn-in), statistics accumulated' % (cores, sweeps, burn))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(args.steps, 1), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ---------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------

class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 10 ms from a thread
    (the `nvidia-smi` query line of B200_PROFILING.md as a fallback when pynvml is missing)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []         # (sm MHz, reasons bit mask)
        self.stop_flag = False
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.gpu
            if vis:
                ids = [x for x in vis.split(',') if x.strip() != '']
                if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
                    idx = int(ids[self.gpu])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll)
            self.thread.daemon = True
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                    ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                        '--format=csv,noheader,nounits', '-lms', '50'],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._pump)
            self.thread.daemon = True
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
                getattr(n, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self.stop_flag:
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                    int(get_reasons(self.handle))))
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            n = self.nvml
            names = (('hw_slowdown', 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                    ('hw_thermal_slowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                    ('sw_thermal_slowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                    ('sw_power_cap', 'nvmlClocksThrottleReasonSwPowerCap', 0x4))
            reasons = set()
            for _, mask in self.samples:
                for name, attr, default in names:
                    if mask & int(getattr(n, attr, default)):
                        reasons.add(name)
            sm = [x for x, _ in self.samples]
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.max_mhz,
                    'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvml, 10 ms period'}
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                    'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi -lms 50'}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------

def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    nb = __graft_entry__.build()
    from nxblink_b200 import packing, distributed as nd

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sites, chains = args.sites, args.chains
    m, pm, ls, tt, tf = p53_problem(sites, seed=rank)      # every rank: its own synthetic sites
    data = packing.pack_alignment(pm, ls, tt, tf)
    sampler = nb.BlinkSampler(pm, device=local_rank)
    sampler.set_data(data)
    sampler.init_chains(chains=chains, seed=2024, unit_offset=rank * sites * chains)
    n_units = sites * chains
    sampler.sweep(args.burnin, accumulate=False)             # reach the stationary event load

    for _ in range(args.warmup):
        sampler.sweep(1, accumulate=True)
    st0 = sampler.stats()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    kind_ms = {}
    for _ in range(args.steps):
        sampler.sweep(1, accumulate=True)
        dev_ms += sampler.last_kernel_ms()[0]
        for k, (ms, n) in sampler.last_kernel_ms_by_kind().items():
            a = kind_ms.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += n
    barrier()
    wall = time.perf_counter() - t0
    clock_info = clocks.stop()
    st1 = sampler.stats()
    launches = st1['kernel_launches'] - st0['kernel_launches']

    # ---- end to end through the host-buffer API --------------------------------
    pin = [torch.from_numpy(a).pin_memory() for a in (
        data.pri_ok.view(np.int64), data.tol_true_ok.view(np.int32), data.tol_false_ok.view(np.int32))]
    pinned = packing.PackedData(pin[0].numpy().view(np.uint64), pin[1].numpy().view(np.uint32),
            pin[2].numpy().view(np.uint32))
    h2d = int(pinned.pri_ok.nbytes + pinned.tol_true_ok.nbytes + pinned.tol_false_ok.nbytes)
    d2h = int(8 * (sampler.layout.n_counts + sampler.layout.n_dwell))
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        sampler.set_data(pinned)
        sampler.sweep(1, accumulate=True)
        counts, dwell = sampler.read_summary()
    barrier()
    wall_e2e = time.perf_counter() - t1

    # ---- the one collective of the path: statistics all-reduce ------------------------
    allreduce_ms = None
    if world > 1:
        c_t, d_t = nd.device_summary_tensors(sampler)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        nd.allreduce_summary(c_t, d_t)
        ev1.record()
        torch.cuda.synchronize()
        allreduce_ms = ev0.elapsed_time(ev1)
        counts, dwell = sampler.read_summary()

    # max over ranks
    t = torch.tensor([dev_ms, wall * 1e3, wall_e2e * 1e3], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max, e2e_ms_max = [float(x) for x in t.tolist()]
    # per-rank device time and SM clock, for reading the scaling numbers
    mine = torch.tensor([dev_ms / args.steps, float(clock_info.get('sm_mhz') or 0.0)],
            dtype=torch.float64, device='cuda')
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(v), 4) for v in x.tolist()] for x in per_rank]

    if rank == 0:
        total_units = n_units * world
        value = total_units * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_units * args.steps / (e2e_ms_max * 1e-3)
        # algorithmic bytes per site-sweep from the ACTUAL live state (never the padded slab)
        state_bytes = st1['state_bytes_used'] / st1['n_units']
        data_bytes_per_site = pm.n_nodes * 16
        bytes_per_unit = 2.0 * state_bytes + data_bytes_per_site / chains
        # One sweep of the batch is two launches (primary-only sweep_kernel, then blink_kernel),
        # each over all units.  The roofline is quoted for the dominant one (blink_kernel) on
        # the bytes IT has to move per unit: the whole trajectory and the tolerance data in,
        # node tolerance states and the new blink events out.  `step` below is the same
        # ratio for the whole sweep on SURVEY 8(d)'s per-sweep figure.
        kernels = dict((k, {'ms_per_step': v[0] / args.steps, 'launches': v[1]})
                for k, v in kind_ms.items() if v[1])
        dom = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        N_nodes = pm.n_nodes
        blk_bytes = 12.0 * st1['n_blk_events'] / st1['n_units']
        if dom == 'blink_kernel':
            dom_bytes = state_bytes + (8 + 4 * N_nodes + blk_bytes) + 8.0 * N_nodes / chains
        elif dom == 'sweep_kernel (primary only)':
            dom_bytes = state_bytes + (state_bytes - 4 * N_nodes - blk_bytes) + 16.0 * N_nodes / chains
        else:
            dom_bytes = bytes_per_unit
        dom_s = kernels[dom]['ms_per_step'] * 1e-3 / max(kernels[dom]['launches'] / args.steps, 1)
        achieved = dom_bytes * n_units / dom_s / 1e9
        step_achieved = bytes_per_unit * n_units * args.steps / (dev_ms * 1e-3) / 1e9
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                peak = float(json.load(f)['hbm_gbs'])
                peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                # ncu capture (one launch at a smaller batch) -> bytes per unit-sweep,
                # scaled to the units one timed launch processes
                tj = json.load(f)
                traffic = tj.get('dram_bytes_per_unit', {}).get(dom, tj['dram_bytes_per_unit_sweep']) * n_units
        except Exception:
            pass
        cores = host_cores()
        if args.cpu_sweeps > 0:
            cpu_val, cpu_wall = cpu_site_sweeps_per_sec(cores, 1, args.cpu_sweeps)
        else:
            cpu_val, cpu_wall = None, 0.0
        R = (st1['n_pri_events'] + st1['n_blk_events']) / st1['n_units']
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dev_ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {
                'workload': WORKLOAD, 'sites_per_gpu': sites, 'chains_per_site': chains,
                'units_per_gpu': n_units, 'burnin_sweeps': args.burnin,
                'rate_on': 1.5, 'rate_off': 0.5, 'mean_retained_events_per_unit': R,
                'l2': 'working set %.1f GB per GPU >> 126 MB L2 (inputs larger than L2)' % (
                    st1['state_bytes_used'] / 1e9),
                'parallelism': 'sites sharded over %d GPU(s), no data-path collective' % world,
                'warps_per_cta': st1['warps_per_cta'], 'ctas': st1['n_ctas'],
                'smem_bytes_per_cta': st1['smem_bytes'],
                'deferred_unit_launch_fraction': (st1['deferred_units'] - st0['deferred_units'])
                    / float(n_units * args.steps),
            },
            'wall_ms_per_step': wall_ms_max / args.steps,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h},
            'gpu_launches': int(launches),
            'clocks': clock_info,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'kernel': dom, 'kernel_bytes_per_unit': dom_bytes,
                'kernel_share_of_step': kernels[dom]['ms_per_step'] / (dev_ms / args.steps),
                'bytes_per_site_sweep': bytes_per_unit,
                'step': {'achieved': step_achieved, 'frac': step_achieved / peak},
                'note': 'the sweep is instruction/latency bound, not HBM bound: see DESIGN.md'},
            'cpu_baseline': {'value': cpu_val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                'sample': '%d host processes x 1 p53-shaped site x 1 chain x %d sweeps (after 1 '
                    'burn-in), %.1f s wall' % (cores, args.cpu_sweeps, cpu_wall)},
            'allreduce_ms': allreduce_ms,
            'kernels': kernels,
            'per_rank_ms_and_sm_mhz': per_rank,
        }
        print(json.dumps(out))
    sampler.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=('b200', 'reference'))
    ap.add_argument('--sites', type=int, default=15625)
    ap.add_argument('--chains', type=int, default=64)
    ap.add_argument('--burnin', type=int, default=12)
    ap.add_argument('--cpu-sweeps', type=int, default=6)
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch one process per GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                '--nproc-per-node', str(args.gpus), '--master-addr', '127.0.0.1',
                '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu_arm(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
