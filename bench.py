#!/usr/bin/env python
"""Headline benchmark: Rao-Teh site-sweeps/sec at the p53 codon shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one full Gibbs sweep (primary track + all 20 tolerance tracks,
statistics accumulated) of every (site, chain) unit, plus -- at N > 1 -- the NCCL
all-reduce of the statistics, all inside the timed region and on one CUDA stream:
BASELINE.json configs[2] -- the p53-shaped codon blink model (49-node tree, 61
codon states, 20 tolerance classes), synthetic alignment, 15 625 sites x 64
chains = 10^6 units IN TOTAL, the same global problem for every N (sites sharded
over the ranks, no data-path collective: "scaling": "strong"; the weak-scaling
figure, 10^6 units on every GPU, is reported beside it in `weak`).  One process
per GPU.

Printed JSON (one line, rank 0): see the contract in the task description.  In
short: `value` = whole-job site-sweeps/s with everything resident in HBM, timed
with CUDA events on the launching stream (max over ranks); `e2e` = the same metric
through the public host-buffer API with the step's site data uploaded from pinned
host memory and the statistics read back inside the timed region; `roofline` =
algorithmic HBM bytes of the dominant kernel / its measured duration against the
measured copy bandwidth in MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle
port (oracle/blink_oracle.py, a restatement of the reference's own Python path)
timed on this box's host cores on a bounded sample; `precision` = the same run
with f64 FFBS messages; `statistics_hash` = hash of the all-reduced integer
statistics (equal for every N); `configs` (N = 1) = BASELINE configs[1] and [4].

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

DTYPE = 'f64 event times, penalty exponents and statistics; f32 FFBS messages (msg_f64=0)'


def fp64_tensor_peak(torch, n=4096, reps=5):
    """FP64 tensor-pipe (DMMA) peak measured on this box: cuBLAS DGEMM n^3, best of `reps`
    (the denominator of the dense kernels' roofline fractions; MEASURED_PEAKS.json has no FP64 entry)."""
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    torch.matmul(a, b)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


class Runner(object):
    """One sampler on this rank's GPU plus the statistics all-reduce of the path, all on ONE
    CUDA stream (nxb_set_stream), so that torch events on that stream bracket sweeps and
    collective alike."""

    def __init__(self, nb, torch, dist, pm, data, chains, unit_offset, local_rank, world, msg_f64=False):
        from nxblink_b200 import distributed as nd
        self.torch, self.dist, self.world = torch, dist, world
        self.stream = torch.cuda.Stream()
        self.s = nb.BlinkSampler(pm, device=local_rank, msg_f64=msg_f64)
        self.s.set_stream(self.stream.cuda_stream)
        self.data = data
        self.s.set_data(data)
        self.s.init_chains(chains=chains, seed=2024, unit_offset=unit_offset)
        self.n_units = data.n_sites * chains
        self.c_t, self.d_t = nd.device_summary_tensors(self.s)
        self.c_red, self.d_red = torch.empty_like(self.c_t), torch.empty_like(self.d_t)
        self.nd = nd

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self):
        """The one collective of the path (run-stochastic.py:172-195 keeps one Summary over all
        sites): all-reduce of COPIES of the two statistic buffers, so the accumulators stay local."""
        with self.torch.cuda.stream(self.stream):
            self.c_red.copy_(self.c_t)
            self.d_red.copy_(self.d_t)
            if self.world > 1:
                self.nd.allreduce_summary(self.c_red, self.d_red)

    def timed(self, steps, e2e=False, pinned=None, host_out=None):
        """K steps = K x (one sweep of every local unit with statistics + the all-reduce)."""
        torch = self.torch
        kind_ms = {}
        ar_ms = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.perf_counter()
        ev0.record(self.stream)
        for _ in range(steps):
            if e2e:
                self.s.set_data(pinned)                    # H2D of this step's site data (pinned host memory)
            self.s.sweep(1, accumulate=True)
            for k, (ms, n) in self.s.last_kernel_ms_by_kind().items():
                a = kind_ms.setdefault(k, [0.0, 0])
                a[0] += ms
                a[1] += n
            if self.world > 1 or e2e:
                a0.record(self.stream)
                self.reduce()
                a1.record(self.stream)
            if e2e:
                with torch.cuda.stream(self.stream):       # D2H of the step's (global) statistics
                    host_out[0].copy_(self.c_red, non_blocking=True)
                    host_out[1].copy_(self.d_red, non_blocking=True)
                self.stream.synchronize()
            elif self.world > 1:
                self.stream.synchronize()
            if self.world > 1 or e2e:
                ar_ms += a0.elapsed_time(a1)
        ev1.record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        return ev0.elapsed_time(ev1), wall * 1e3, kind_ms, ar_ms

    def counts_hash(self):
        """sha256 of the all-reduced integer statistics: identical for every number of GPUs
        (random streams are keyed by the global unit index)."""
        import hashlib
        self.reduce()
        self.stream.synchronize()
        return hashlib.sha256(self.c_red.cpu().numpy().tobytes()).hexdigest()[:16]

    def close(self):
        self.s.close()


def max_over_ranks(torch, dist, world, vals):
    t = torch.tensor(vals, dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def toy_config(nb):
    """BASELINE configs[1]: codon2x3 toy model B, 10 000 sites x 64 chains (synthetic alignment
    forward-sampled from the model itself, full tolerance observation at one leaf)."""
    from nxblink_b200 import toymodel, packing
    pm = packing.PackedModel(toymodel.BlinkModelB)
    s = nb.BlinkSampler(pm)
    node_pri, node_tol = s.forward_sample(10000, seed=0)
    s.close()
    is_leaf = np.ones(pm.n_nodes, dtype=bool)
    is_leaf[pm.parent[1:]] = False
    ls = np.where(is_leaf[None, :], node_pri, -1)
    allk = np.uint32(7)
    tt = np.full(ls.shape, allk, np.uint32)
    tf = np.full(ls.shape, allk, np.uint32)
    ref = pm.node_index['N0']
    tt[:, ref] = node_tol[:, ref]
    tf[:, ref] = (~node_tol[:, ref]) & allk
    data = packing.pack_alignment(pm, ls, tt, tf)
    s = nb.BlinkSampler(pm)
    s.set_data(data)
    s.init_chains(chains=64, seed=1)
    s.sweep(10, accumulate=False)
    best = 1e30
    for _ in range(3):
        s.sweep(10, accumulate=True)
        best = min(best, s.last_kernel_ms()[0])
    s.close()
    return {'workload': 'codon2x3 toy model B (6 primary states, 3 tolerance classes, 6-node tree), '
            '10000 sites x 64 chains, synthetic', 'value': 640000 * 10 / (best * 1e-3),
            'unit': UNIT, 'ms_per_sweep_of_batch': best / 10, 'timing': 'CUDA events, 10 sweeps, best of 3'}


def em_config(nb, torch, fp64_peak):
    """BASELINE configs[4]: one EM iteration at the real p53 size (393 columns x 64 chains,
    10 burn-in + 10 sampling sweeps; run-stochastic.py:154-208, p53em.py:58-140), split into its
    parts, plus the dense kernels at bench size with their FP64 tensor-pipe roofline fractions."""
    from nxblink_b200 import p53, packing, dense, p53em, _lib
    lib = _lib.load()
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 393, seed=1, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    P = pm.n_states
    Q = np.zeros((P, P))
    Q[pm.q_row, pm.q_col] = pm.q_val
    Q -= np.diag(Q.sum(axis=1))
    t = pm.blen[1:] * pm.rate[1:]

    def timed(f):
        t0 = time.perf_counter()
        r = f()
        return r, 1e3 * (time.perf_counter() - t0)
    PARTS = ('expm_48x61x61_ms', 'pruning_393_sites_ms', 'init_chains_ms', 'sweeps_20_ms', 'read_summary_ms', 'mstep_ms')
    s = nb.BlinkSampler(pm)
    s.set_data(data)
    best = None
    for rep in range(4):                       # the first repetition warms everything up; then the fastest of three
        out = {}
        Pe, out['expm_48x61x61_ms'] = timed(lambda: dense.expm_batched(Q, t))
        _, out['pruning_393_sites_ms'] = timed(lambda: dense.pruning_loglik(pm.parent, Pe, pm.prior, data.pri_ok))
        _, out['init_chains_ms'] = timed(lambda: s.init_chains(chains=64, seed=rep))
        _, a = timed(lambda: s.sweep(10, accumulate=False))
        k = s.last_kernel_ms()[0]
        _, b = timed(lambda: s.sweep(10, accumulate=True))
        k += s.last_kernel_ms()[0]
        out['sweeps_20_ms'] = a + b
        out['sweeps_20_kernel_ms'] = k
        sm, out['read_summary_ms'] = timed(lambda: s.summary())
        prm = p53.jeff_params()
        res, out['mstep_ms'] = timed(lambda: p53em.m_step(s, sm, m.genetic_code, prm['kappa'], prm['omega'],
            prm['A'], prm['C'], prm['G'], prm['T'], 1.5, 0.5, np.maximum(pm.rate[1:], 1e-6)))
        out['mstep_where'] = res.get('where', 'host')
        if isinstance(res.get('result'), dict):
            out['mstep_kernel'] = res['result']
        if rep >= 1 and (best is None or sum(v for k, v in out.items() if k in PARTS) < sum(v for k, v in best.items() if k in PARTS)):
            best = dict(out)
    out = best
    s.close()
    out['allreduce_ms'] = None         # one GPU; at N > 1 see `allreduce_ms_per_step` of the main figure
    out['total_ms'] = sum(out[k] for k in PARTS)
    # dense kernels at bench size: kernel time from CUDA events inside the library call
    many = np.tile(t, 64)
    dense.expm_batched(Q, many)
    dense.expm_batched(Q, many)
    k_expm = float(lib.nxb_dense_last_kernel_ms())
    nrm = np.abs(Q).sum(axis=0).max() * np.abs(many)
    squarings = np.where(nrm > 0.5, np.ceil(np.log2(np.maximum(nrm, 1e-300) / 0.5)), 0.0)
    fl_expm = float((2.0 * 64 ** 3 * (13 + squarings)).sum())
    ls2, _, _ = p53.synthetic_alignment(pm, 15625, seed=0)
    d2 = packing.pack_alignment(pm, ls2)
    dense.pruning_loglik(pm.parent, Pe, pm.prior, d2.pri_ok)
    dense.pruning_loglik(pm.parent, Pe, pm.prior, d2.pri_ok)
    k_prune = float(lib.nxb_dense_last_kernel_ms())
    S, N = d2.pri_ok.shape
    fl_prune = 2.0 * 64 * 64 * 32 * (N - 1) * ((S + 31) // 32)
    out['dense'] = {
        'fp64_tensor_peak_tflops': fp64_peak,
        'peak_source': 'measured here: torch.matmul f64 4096^3 (cuBLAS DGEMM), best of 5, CUDA events',
        'expm_kernel': {'matrices': int(len(many)), 'order': P, 'kernel_ms': k_expm,
            'tflops': fl_expm / (k_expm * 1e-3) / 1e12, 'frac': fl_expm / (k_expm * 1e-3) / 1e12 / fp64_peak,
            'flops': 'padded 64^3 products: 13 Horner steps + squarings per matrix'},
        'pruning_kernel': {'sites': int(S), 'nodes': int(N), 'kernel_ms': k_prune,
            'tflops': fl_prune / (k_prune * 1e-3) / 1e12, 'frac': fl_prune / (k_prune * 1e-3) / 1e12 / fp64_peak,
            'flops': '[64x64].[64x32] contraction per edge and tile of 32 sites'}}
    out['workload'] = ('EM iteration, p53-shaped codon blink model, 393 columns x 64 chains, 10 burn-in + 10 '
            'sampling sweeps, wall ms of each host-API call (the fastest of three repetitions after one warm-up)')
    return out


def run_gpu_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    nb = __graft_entry__.build()
    from nxblink_b200 import packing, distributed as nd

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    sites, chains = args.sites, args.chains
    # ---- main figure: ONE global problem (sites x chains units), sharded by site over the ranks ----
    m, pm, ls, tt, tf = p53_problem(sites, seed=0)
    data_all = packing.pack_alignment(pm, ls, tt, tf)
    a, b = nd.shard_range(sites, rank, world)
    local = packing.PackedData(data_all.pri_ok[a:b], data_all.tol_true_ok[a:b], data_all.tol_false_ok[a:b])
    run = Runner(nb, torch, dist, pm, local, chains, a * chains, local_rank, world)
    s = run.s
    n_units = (b - a) * chains
    total_units = sites * chains
    s.sweep(args.burnin, accumulate=False)             # reach the stationary event load
    for _ in range(args.warmup):
        s.sweep(1, accumulate=True)
        run.reduce()
    s.reset_summary()
    st0 = s.stats()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev_ms, wall_ms, kind_ms, ar_ms = run.timed(args.steps)
    clock_info = clocks.stop()
    st1 = s.stats()
    launches = st1['kernel_launches'] - st0['kernel_launches']
    counts_hash = run.counts_hash()
    nsamples = int(run.c_red[s.layout.c_nsamples].item())

    # ---- end to end through the host-buffer API --------------------------------
    pin = [torch.from_numpy(x).pin_memory() for x in (
        local.pri_ok.view(np.int64), local.tol_true_ok.view(np.int32), local.tol_false_ok.view(np.int32))]
    pinned = packing.PackedData(pin[0].numpy().view(np.uint64), pin[1].numpy().view(np.uint32),
            pin[2].numpy().view(np.uint32))
    host_out = (torch.empty(s.layout.n_counts, dtype=torch.int64).pin_memory(),
            torch.empty(s.layout.n_dwell, dtype=torch.float64).pin_memory())
    h2d = int(pinned.pri_ok.nbytes + pinned.tol_true_ok.nbytes + pinned.tol_false_ok.nbytes)
    d2h = int(8 * (s.layout.n_counts + s.layout.n_dwell))
    e2e_ev_ms, e2e_wall_ms, _, _ = run.timed(args.steps, e2e=True, pinned=pinned, host_out=host_out)

    dev_ms_max, wall_ms_max, e2e_ms_max, ar_ms_max = max_over_ranks(torch, dist, world,
            [ev_ms, wall_ms, e2e_wall_ms, ar_ms])
    mine = torch.tensor([ev_ms / args.steps, float(clock_info.get('sm_mhz') or 0.0)],
            dtype=torch.float64, device='cuda')
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(v), 4) for v in x.tolist()] for x in per_rank]
    run.close()

    # ---- second figure at N > 1: weak scaling, args.sites sites on EVERY GPU (own data per rank) ----
    weak = None
    if world > 1:
        m2, pm2, ls2, tt2, tf2 = p53_problem(sites, seed=rank)
        run2 = Runner(nb, torch, dist, pm2, packing.pack_alignment(pm2, ls2, tt2, tf2), chains,
                rank * sites * chains, local_rank, world)
        run2.s.sweep(args.burnin, accumulate=False)
        for _ in range(args.warmup):
            run2.s.sweep(1, accumulate=True)
            run2.reduce()
        w_ev, w_wall, _, w_ar = run2.timed(args.steps)
        w_ev_max, w_ar_max = max_over_ranks(torch, dist, world, [w_ev, w_ar])
        weak = {'value': sites * chains * world * args.steps / (w_ev_max * 1e-3), 'unit': UNIT,
                'ms_per_step': w_ev_max / args.steps, 'units_per_gpu': sites * chains,
                'allreduce_ms_per_step': w_ar_max / args.steps,
                'note': 'every GPU holds %d units of its own synthetic sites; all-reduce inside the step' % (sites * chains)}
        run2.close()

    # ---- N = 1: the FP64-message variant of the same run, and the other BASELINE configs ----
    precision, configs = None, None
    if world == 1:
        run64 = Runner(nb, torch, dist, pm, local, chains, 0, local_rank, world, msg_f64=True)
        run64.s.sweep(args.burnin, accumulate=False)
        for _ in range(min(args.warmup, 3)):
            run64.s.sweep(1, accumulate=True)
        k64 = max(1, min(args.steps, 3))
        ev64, _, kind64, _ = run64.timed(k64)
        precision = {'msg_f32_value': total_units * args.steps / (dev_ms_max * 1e-3),
                'msg_f64_value': total_units * k64 / (ev64 * 1e-3), 'unit': UNIT,
                'msg_f64_ms_per_step': ev64 / k64,
                'msg_f64_kernels_ms_per_step': dict((k, v[0] / k64) for k, v in kind64.items() if v[1]),
                'note': 'msg_f64=1: chunk messages, tolerance (OFF, ON) messages, exp(-penalty), gap logarithm and '
                    'backward-sampling uniforms in f64 (chunking.py:32-36, 264-289); same workload, %d steps' % k64}
        run64.close()
        if not args.no_configs:
            peak64 = fp64_tensor_peak(torch)
            configs = {'codon2x3_10000x64': toy_config(nb), 'em_iteration_p53_393x64': em_config(nb, torch, peak64)}

    if rank == 0:
        value = total_units * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_units * args.steps / (e2e_ms_max * 1e-3)
        # algorithmic bytes per site-sweep from the ACTUAL live state (never the padded slab)
        state_bytes = st1['state_bytes_used'] / st1['n_units']
        data_bytes_per_site = pm.n_nodes * 16
        bytes_per_unit = 2.0 * state_bytes + data_bytes_per_site / chains
        # One sweep of the batch is two launches (primary-only sweep_kernel, then blink_kernel),
        # each over all units.  The roofline is quoted for the dominant one on the bytes IT has to
        # move per unit; `step` below is the same ratio for the whole sweep on SURVEY 8(d)'s figure.
        kernels = dict((k, {'ms_per_step': v[0] / args.steps, 'launches': v[1]})
                for k, v in kind_ms.items() if v[1])
        main_kernels = dict((k, v) for k, v in kernels.items() if k != 'sweep_kernel (fused)')
        dom = max(main_kernels, key=lambda k: main_kernels[k]['ms_per_step'])
        N_nodes = pm.n_nodes
        blk_bytes = 12.0 * st1['n_blk_events'] / st1['n_units']
        if dom == 'blink_kernel':
            dom_bytes = state_bytes + (8 + 4 * N_nodes + blk_bytes) + 8.0 * N_nodes / chains
        else:
            dom_bytes = state_bytes + (state_bytes - 4 * N_nodes - blk_bytes) + 16.0 * N_nodes / chains
        dom_s = kernels[dom]['ms_per_step'] * 1e-3 / max(kernels[dom]['launches'] / args.steps, 1)
        achieved = dom_bytes * n_units / dom_s / 1e9
        kern_total = sum(v['ms_per_step'] for v in kernels.values())
        step_achieved = bytes_per_unit * n_units / (kern_total * 1e-3) / 1e9
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
        if args.cpu_sweeps > 0 and world == 1:
            cpu_val, cpu_wall = cpu_site_sweeps_per_sec(cores, 1, args.cpu_sweeps)
        else:
            cpu_val, cpu_wall = None, 0.0
        R = (st1['n_pri_events'] + st1['n_blk_events']) / st1['n_units']
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dev_ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
            'config': {
                'workload': WORKLOAD + '; %d sites x %d chains = %d units in total, the SAME global problem for '
                    'every N (sites sharded over the GPUs)' % (sites, chains, total_units),
                'global_sites': sites, 'chains_per_site': chains, 'global_units': total_units,
                'units_on_rank0': n_units, 'burnin_sweeps': args.burnin,
                'rate_on': 1.5, 'rate_off': 0.5, 'mean_retained_events_per_unit': R,
                'l2': 'working set %.2f GB per GPU >> 126 MB L2 (inputs larger than L2)' % (
                    st1['state_bytes_used'] / 1e9),
                'parallelism': 'sites sharded over %d GPU(s); no data-path collective; one all-reduce of the '
                    'statistics (0.7 MB) per step INSIDE the timed region' % world,
                'step': 'one Gibbs sweep of every unit with statistics accumulated'
                    + (' + NCCL all-reduce of the statistics' if world > 1 else ''),
                'timing': 'torch CUDA events on the stream the kernels are launched on (nxb_set_stream), '
                    'max over ranks; kernels[] from the library\'s own events around each launch',
                'warps_per_cta': st1['warps_per_cta'], 'ctas': st1['n_ctas'],
                'kernels_note': 'warps_per_cta / smem_bytes_per_cta: the primary-only sweep_kernel; blink_kernel runs 25 warps '
                    '(5 groups of 5 = 40 units x 20 tracks) per CTA at the p53 shape',
                'smem_bytes_per_cta': st1['smem_bytes'],
                'deferred_unit_launch_fraction': (st1['deferred_units'] - st0['deferred_units'])
                    / float(n_units * args.steps),
            },
            'wall_ms_per_step': wall_ms_max / args.steps,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_ms_max / args.steps,
                'note': 'per step: site data H2D from pinned memory, sweep, statistics '
                    + ('all-reduce, ' if world > 1 else '') + 'D2H into pinned memory; wall clock between barriers'},
            'gpu_launches': int(launches),
            'clocks': clock_info,
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'kernel': dom, 'kernel_bytes_per_unit': dom_bytes,
                'kernel_share_of_step': kernels[dom]['ms_per_step'] / kern_total,
                'bytes_per_site_sweep': bytes_per_unit,
                'step': {'achieved': step_achieved, 'frac': step_achieved / peak},
                'note': 'the sweep is instruction/latency bound, not HBM bound: see DESIGN.md'},
            'cpu_baseline': {'value': cpu_val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                'sample': '%d host processes x 1 p53-shaped site x 1 chain x %d sweeps (after 1 '
                    'burn-in), %.1f s wall' % (cores, args.cpu_sweeps, cpu_wall)},
            'allreduce_ms_per_step': (ar_ms_max / args.steps) if world > 1 else None,
            'statistics_hash': {'sha256_16': counts_hash, 'nsamples': nsamples,
                'of': 'all-reduced integer statistics (u64 counts) after the %d timed steps; equal for every N' % args.steps},
            'kernels': kernels,
            'per_rank_ms_and_sm_mhz': per_rank,
        }
        if weak is not None:
            out['weak'] = weak
        if precision is not None:
            out['precision'] = precision
        if configs is not None:
            out['configs'] = configs
        print(json.dumps(out))
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
    ap.add_argument('--no-configs', action='store_true', help='skip the codon2x3 and EM-iteration figures (N=1)')
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
