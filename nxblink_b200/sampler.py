"""Batched Rao-Teh sampler: the host-side handle around the C-ABI.

``BlinkSampler`` runs the Gibbs sweep of nxblink/raoteh.py:364-435 for a whole
batch of alignment sites x independent chains on one B200 and reduces the
sufficient statistics of nxblink/summary.py on the device.  There is no CPU
path: constructing it without the CUDA library or without a device raises.
"""
from __future__ import division, print_function

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NxbError
from .packing import PackedModel, PackedData, pack_data
from .summary import Summary


class Event(object):
    """Same fields as ``nxctmctree.trajectory.Event`` (track, tm, sa, sb)."""
    __slots__ = ('track', 'tm', 'sa', 'sb')

    def __init__(self, track=None, tm=None, sa=None, sb=None):
        self.track, self.tm, self.sa, self.sb = track, tm, sa, sb

    def __repr__(self):
        return 'Event(%r, %r, %r, %r)' % (
                getattr(self.track, 'name', None), self.tm, self.sa, self.sb)


class Track(object):
    """What ``gen_samples`` yields per track: name, history, events (trajectory.py:17-50)."""

    def __init__(self, name):
        self.name = name
        self.history = {}
        self.events = {}


class BlinkSampler(object):
    def __init__(self, model, device=0, msg_f64=False, cap_scale=1.0,
            validate=False, max_warps=0, lockstep_warps=0, split_phases=0):
        self.lib = _lib.load()
        self.device = int(device)
        self.pm = model if isinstance(model, PackedModel) else PackedModel(model)
        self._h = C.c_void_p()
        self._desc_kwargs = dict(msg_f64=msg_f64, cap_scale=cap_scale,
                validate=validate, max_warps=max_warps, lockstep_warps=lockstep_warps,
                split_phases=split_phases)
        desc = self.pm.desc(**self._desc_kwargs)
        rc = self.lib.nxb_create(C.byref(desc), int(device), C.byref(self._h))
        if rc:
            msg = self.lib.nxb_last_error(None).decode()
            self._h = None
            raise NxbError(rc, msg)
        self.layout = _lib.SummaryLayout()
        self._check(self.lib.nxb_summary_layout_get(self._h, C.byref(self.layout)))
        self.data = None
        self.chains = 0
        self.n_units = 0

    def update_model(self, model):
        """Swap in a model of the same shape with new rates / branch lengths / prior (what an EM
        iteration changes) without rebuilding the handle; data and chains are kept (re-initialise
        them with ``init_chains`` for the reference's behaviour, or continue them as a warm start)."""
        pm = model if isinstance(model, PackedModel) else PackedModel(model)
        if pm.nodes != self.pm.nodes or pm.states != self.pm.states or pm.tols != self.pm.tols:
            raise NxbError(5, 'update_model: the model must have the same tree, states and classes')
        desc = pm.desc(**self._desc_kwargs)
        self._check(self.lib.nxb_update_model(self._h, C.byref(desc)))
        self.pm = pm

    def _check(self, rc):
        if rc:
            raise NxbError(rc, self.lib.nxb_last_error(self._h).decode())

    def close(self):
        if getattr(self, '_h', None):
            self.lib.nxb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def set_data(self, data):
        """``data``: PackedData, or a sequence of reference-style data objects."""
        if not isinstance(data, PackedData):
            data = pack_data(self.pm, list(data))
        if self.data is not None and data.n_sites != self.data.n_sites:
            self.chains, self.n_units = 0, 0      # the device drops the chains of the old sites
        self.data = data
        self._check(self.lib.nxb_set_data(self._h, data.n_sites,
            data.pri_ok.ctypes.data, data.tol_true_ok.ctypes.data,
            data.tol_false_ok.ctypes.data))

    def init_chains(self, chains=1, seed=0, unit_offset=0):
        if self.data is None:
            # (after forward_sample the device holds unconstrained data of its own)
            raise NxbError(5, 'init_chains: call set_data first')
        self._check(self.lib.nxb_init_chains(self._h, int(chains), int(seed),
            int(unit_offset)))
        self.chains = int(chains)
        self.n_units = self.data.n_sites * self.chains

    def set_stream(self, cuda_stream=0):
        """Run this sampler's launches and copies on the caller's CUDA stream (an integer
        ``cudaStream_t``, e.g. ``torch.cuda.current_stream().cuda_stream``); 0 = its own."""
        self._check(self.lib.nxb_set_stream(self._h, int(cuda_stream)))

    def forward_sample(self, n_units, seed=0, unit_offset=0, root_pri=None, root_tol=None):
        """Prior simulation of ``n_units`` trajectories on the device (fwdsample.py:67-262).
        ``root_pri`` (state indices) and ``root_tol`` (bit masks of ON classes) fix the root
        states; by default they are drawn from the stationary prior.
        Returns (node_pri [n, N] state indices, node_tol [n, N] bit masks)."""
        if root_pri is None:
            self._check(self.lib.nxb_forward_sample(self._h, int(n_units), int(seed), int(unit_offset)))
        else:
            rp = np.ascontiguousarray(root_pri, dtype=np.int32)
            rt = np.ascontiguousarray(root_tol, dtype=np.uint32)
            if rp.shape != (int(n_units),) or rt.shape != (int(n_units),):
                raise NxbError(5, 'root_pri and root_tol must have one entry per unit')
            self._check(self.lib.nxb_forward_sample_rooted(self._h, int(n_units), int(seed),
                    int(unit_offset), rp.ctypes.data, rt.ctypes.data))
        self.chains, self.n_units = 1, int(n_units)
        self.data = None
        node_pri = np.zeros((self.n_units, self.pm.n_nodes), dtype=np.int32)
        node_tol = np.zeros((self.n_units, self.pm.n_nodes), dtype=np.uint32)
        self._check(self.lib.nxb_read_node_states(self._h, node_pri.ctypes.data, node_tol.ctypes.data))
        return node_pri, node_tol

    def sweep(self, n_sweeps=1, accumulate=True):
        self._check(self.lib.nxb_sweep(self._h, int(n_sweeps), 1 if accumulate else 0))

    def reset_summary(self):
        self._check(self.lib.nxb_summary_reset(self._h))

    def summarize_current(self):
        self._check(self.lib.nxb_summarize_current(self._h))

    def read_summary(self):
        counts = np.zeros(self.layout.n_counts, dtype=np.uint64)
        dwell = np.zeros(self.layout.n_dwell, dtype=np.float64)
        self._check(self.lib.nxb_summary_read(self._h, counts.ctypes.data, dwell.ctypes.data))
        return counts, dwell

    def summary_device_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.lib.nxb_summary_device_ptrs(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def summary(self, counts=None, dwell=None, Q_primary=None):
        if counts is None:
            counts, dwell = self.read_summary()
        return Summary.from_device(self.pm, self.layout, counts, dwell, Q_primary)

    def stats(self):
        st = _lib.Stats()
        self._check(self.lib.nxb_get_stats(self._h, C.byref(st)))
        return dict((f, getattr(st, f)) for f, _ in st._fields_)

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_int32()
        self._check(self.lib.nxb_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    KERNELS = ('sweep_kernel (fused)', 'sweep_kernel (primary only)', 'blink_kernel', 'check')

    def last_kernel_ms_by_kind(self):
        """{kernel: (ms, launches)} of the last ``sweep`` / ``init_chains`` call."""
        ms, n = (C.c_float * 4)(), (C.c_int32 * 4)()
        self._check(self.lib.nxb_last_kernel_ms_by_kind(self._h, ms, n))
        return dict((k, (ms[i], n[i])) for i, k in enumerate(self.KERNELS))

    PHASES = ('load', 'P0 edge-major view', 'P1 candidates', 'P2 chunks', 'P3 upward',
            'P4 downward', 'P5 writeback+stats', 'B0 pool', 'B1 upward', 'B2 downward+stats',
            'BW writeback', 'store')

    # tolpar_kernel books its phases into slots 12..23
    TP_PHASES = ('T load', 'T candidates', 'T merge', 'T thinning', 'T matrices', 'T suffix scan',
            'T tree up', 'T maps', 'T tree down', 'T survivors', 'T event-free', 'T store')

    def profile(self, enable=True):
        """Switch the per-phase cycle counters on/off; returns the counters so far."""
        out = np.zeros(24, dtype=np.uint64)
        self._check(self.lib.nxb_profile(self._h, 1 if enable else 0, out.ctypes.data))
        return dict(zip(self.PHASES + self.TP_PHASES, out.tolist()))

    # ------------------------------------------------------------------
    def export_unit(self, unit, cap=4096):
        """Raw integer-indexed trajectory of one unit."""
        N = self.pm.n_nodes
        node_pri = np.zeros(N, dtype=np.int32)
        node_tol = np.zeros(N, dtype=np.uint32)
        n_pri, n_blk = C.c_int32(), C.c_int32()
        pt = np.zeros(cap); pe = np.zeros(cap, np.int32)
        psa = np.zeros(cap, np.int32); psb = np.zeros(cap, np.int32)
        bt = np.zeros(cap); be = np.zeros(cap, np.int32)
        bk = np.zeros(cap, np.int32); bs = np.zeros(cap, np.int32)
        self._check(self.lib.nxb_export_unit(self._h, int(unit),
            node_pri.ctypes.data, node_tol.ctypes.data, cap,
            C.addressof(n_pri), pt.ctypes.data, pe.ctypes.data, psa.ctypes.data,
            psb.ctypes.data, C.addressof(n_blk), bt.ctypes.data, be.ctypes.data,
            bk.ctypes.data, bs.ctypes.data))
        a, b = n_pri.value, n_blk.value
        return dict(node_pri=node_pri, node_tol=node_tol,
                pri=(pt[:a], pe[:a], psa[:a], psb[:a]),
                blk=(bt[:b], be[:b], bk[:b], bs[:b]))

    def export_tracks(self, unit):
        """(primary_track, tolerance_tracks) in the reference's object format
        (what nxblink/raoteh.py:435 yields)."""
        pm = self.pm
        raw = self.export_unit(unit)
        pri = Track('PRIMARY')
        tols = [Track(name) for name in pm.tols]
        for v, node in enumerate(pm.nodes):
            pri.history[node] = pm.states[raw['node_pri'][v]]
            for k, t in enumerate(tols):
                t.history[node] = bool((int(raw['node_tol'][v]) >> k) & 1)
        for edge in pm.edges:
            pri.events[edge] = []
            for t in tols:
                t.events[edge] = []
        for tm, e, sa, sb in zip(*raw['pri']):
            pri.events[pm.edges[e]].append(
                    Event(pri, float(tm), pm.states[sa], pm.states[sb]))
        for tm, e, k, ns in zip(*raw['blk']):
            t = tols[k]
            t.events[pm.edges[e]].append(Event(t, float(tm), not bool(ns), bool(ns)))
        return pri, tols
