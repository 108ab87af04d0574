"""Sufficient statistics of blink-process trajectories.

Host-side mirror of ``nxblink.summary.Summary`` (nxblink/summary.py:53-243): the
same attribute names, read by ``nxblink.em`` (em.py:58-60,124-126,187-195) and by
the reference's drivers.  The accumulation itself happens on the GPU, fused into
the sweep (csrc/sweep_kernel.cuh); this class only unfolds the flat device
buffers into the reference's shapes.

Device primitives -> reference statistics
    T[e][s]       time on edge e in primary state s
    Doff[e][s][k] time on edge e in state s with tolerance class k OFF
    pri_dwell[e][sa->sb] = T[e][sa] - Doff[e][sa][class(sb)]     (summary.py:111-121)
    off_xon_dwell[e]     = sum_{s,k} Doff[e][s][k]               (summary.py:123-131;
                           an OFF class is never the primary state's own class)
    xon_off_dwell[e]     = (K-1) * sum_s T[e][s] - off_xon_dwell[e]
"""
from __future__ import division, print_function

from collections import defaultdict
from io import StringIO

import numpy as np


class WeightGraph(dict):
    """Minimal stand-in for the ``nx.DiGraph`` used by summary.py:102,107.

    ``G[sa][sb]['weight']``, iteration over source states, ``has_edge`` and
    ``edges()`` are what em.py and the reference drivers use.
    """

    def add_edge(self, sa, sb, weight):
        self.setdefault(sa, {})[sb] = {'weight': weight}
        self.setdefault(sb, {})

    def has_edge(self, sa, sb):
        return sa in self and sb in self[sa]

    def edges(self):
        return [(sa, sb) for sa in self for sb in self[sa]]


class Summary(object):
    """Same constructor and attributes as ``nxblink.summary.Summary``.

    Build it from device statistics with ``Summary.from_device`` (what
    ``BlinkSampler.summary()`` does); ``on_sample`` exists for interface parity
    and accumulates host-side from exported tracks.
    """

    def __init__(self, T, root, node_to_tm, primary_to_tol, Q_primary):
        self._T = T
        self._root = root
        self._node_to_tm = node_to_tm
        self._primary_to_tol = primary_to_tol
        self._Q_primary = Q_primary
        self.nsamples = 0
        self.root_pri_to_count = defaultdict(int)
        self.root_xon_count = 0
        self.root_off_count = 0
        # unambiguous primitives behind root_xon_count
        self.root_on_total = 0
        self.root_xon_intended = 0
        self.edge_to_pri_trans = {}
        self.edge_to_off_xon_trans = defaultdict(int)
        self.edge_to_xon_off_trans = defaultdict(int)
        self.edge_to_pri_dwell = {}
        self.edge_to_off_xon_dwell = defaultdict(float)
        self.edge_to_xon_off_dwell = defaultdict(float)
        self.flat = None

    @classmethod
    def from_device(cls, pm, layout, counts, dwell, Q_primary=None):
        """Unfold flat device buffers (see include/nxblink_b200.h nxb_summary_layout)."""
        self = cls(pm.T, pm.root, pm.node_to_tm, pm.primary_to_tol, Q_primary)
        E, P, K, nnz = pm.n_nodes - 1, pm.n_states, pm.n_classes, pm.nnz
        counts = np.asarray(counts, dtype=np.uint64).astype(np.int64)
        dwell = np.asarray(dwell, dtype=float)
        L = layout
        self.nsamples = int(counts[L.c_nsamples])
        root_pri = counts[L.c_root_pri:L.c_root_pri + P]
        on_cls = counts[L.c_root_on_cls:L.c_root_on_cls + K]
        for i, s in enumerate(pm.states):
            if root_pri[i]:
                self.root_pri_to_count[s] = int(root_pri[i])
        self.root_on_total = int(counts[L.c_root_on])
        self.root_off_count = int(counts[L.c_root_off])
        self.root_xon_intended = self.root_on_total - self.nsamples
        # summary.py:221 compares the tolerance STATE with the class NAME of the
        # root's primary state: an ON track is counted unless True == name, i.e.
        # unless the root state's class is literally named 1 / True.
        xon = 0
        for k, name in enumerate(pm.tols):
            if not (name == True):      # noqa: E712  (reference semantics)
                xon += int(on_cls[k])
        self.root_xon_count = xon
        pri_trans = counts[L.c_pri_trans:L.c_pri_trans + E * nnz].reshape(E, nnz)
        off_on = counts[L.c_off_on:L.c_off_on + E]
        on_off = counts[L.c_on_off:L.c_on_off + E]
        T = dwell[L.d_T:L.d_T + E * P].reshape(E, P)
        Doff = dwell[L.d_off:L.d_off + E * P * K].reshape(E, P, K)
        colcls = pm.state_class[pm.q_col]
        pri_dwell = T[:, pm.q_row] - Doff[:, pm.q_row, colcls]          # [E, nnz]
        # T and Doff are sums of the same time pieces in different orders: a class that was
        # OFF for the whole stay leaves a rounding residue instead of an exact zero
        pri_dwell = np.where(pri_dwell > 1e-11 * T[:, pm.q_row], pri_dwell, 0.0)
        off_xon_dwell = Doff.sum(axis=(1, 2))
        xon_off_dwell = (K - 1) * T.sum(axis=1) - off_xon_dwell
        self.flat = dict(pri_trans=pri_trans, pri_dwell=pri_dwell, off_on=off_on,
                on_off=on_off, off_xon_dwell=off_xon_dwell, xon_off_dwell=xon_off_dwell,
                root_pri=root_pri, T=T, Doff=Doff)
        for e, edge in enumerate(pm.edges):
            gt, gd = WeightGraph(), WeightGraph()
            for i in range(nnz):
                sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
                if pri_trans[e, i]:
                    gt.add_edge(sa, sb, int(pri_trans[e, i]))
                if pri_dwell[e, i] > 0:
                    gd.add_edge(sa, sb, float(pri_dwell[e, i]))
            self.edge_to_pri_trans[edge] = gt
            self.edge_to_pri_dwell[edge] = gd
            self.edge_to_off_xon_trans[edge] = int(off_on[e])
            self.edge_to_xon_off_trans[edge] = int(on_off[e])
            self.edge_to_off_xon_dwell[edge] = float(off_xon_dwell[e])
            self.edge_to_xon_off_dwell[edge] = float(xon_off_dwell[e])
        return self

    # -- host-side accumulation from exported tracks (interface parity) --------
    def on_sample(self, primary_track, tolerance_tracks):
        """summary.py:190-243 on tracks as yielded by ``raoteh.gen_samples``."""
        self.nsamples += 1
        root = self._root
        pri_state = primary_track.history[root]
        self.root_pri_to_count[pri_state] += 1
        for t in tolerance_tracks:
            tol_state = t.history[root]
            if not tol_state:
                self.root_off_count += 1
            else:
                self.root_on_total += 1
                if t.name != self._primary_to_tol[pri_state]:
                    self.root_xon_intended += 1
                if tol_state != self._primary_to_tol[pri_state]:
                    self.root_xon_count += 1
        tracks = [primary_track] + list(tolerance_tracks)
        for edge in primary_track.events:
            gt = self.edge_to_pri_trans.setdefault(edge, WeightGraph())
            gd = self.edge_to_pri_dwell.setdefault(edge, WeightGraph())
            for ev in primary_track.events[edge]:
                w = gt[ev.sa][ev.sb]['weight'] if gt.has_edge(ev.sa, ev.sb) else 0
                gt.add_edge(ev.sa, ev.sb, w + 1)
            for t in tolerance_tracks:
                for ev in t.events[edge]:
                    if ev.sa and not ev.sb:
                        self.edge_to_xon_off_trans[edge] += 1
                    elif not ev.sa and ev.sb:
                        self.edge_to_off_xon_trans[edge] += 1
                    else:
                        raise Exception('blink self transition')
            va, vb = edge
            state = dict((t.name, t.history[va]) for t in tracks)
            seq = sorted(((ev.tm, i, ev) for i, ev in enumerate(
                ev for t in tracks for ev in t.events[edge])), key=lambda x: x[:2])
            tm = self._node_to_tm[va]
            for tm_next, ev in [(a, c) for a, b, c in seq] + [(self._node_to_tm[vb], None)]:
                elapsed = tm_next - tm
                sa = state[primary_track.name]
                for sb in self._Q_primary[sa]:
                    if state[self._primary_to_tol[sb]]:
                        w = gd[sa][sb]['weight'] if gd.has_edge(sa, sb) else 0
                        gd.add_edge(sa, sb, w + elapsed)
                own = self._primary_to_tol[sa]
                for t in tolerance_tracks:
                    if t.name != own:
                        if state[t.name]:
                            self.edge_to_xon_off_dwell[edge] += elapsed
                        else:
                            self.edge_to_off_xon_dwell[edge] += elapsed
                if ev is not None:
                    if state[ev.track.name] != ev.sa:
                        raise Exception('incompatible transition')
                    state[ev.track.name] = ev.sb
                    tm = tm_next

    def __str__(self):
        """Same report as summary.py:133-182."""
        s = StringIO()
        print('nsamples:', self.nsamples, file=s)
        print('root summary:', file=s)
        print('  unforced blinked-on count:', self.root_xon_count, file=s)
        print('  blinked-off count:', self.root_off_count, file=s)
        print('  nonzero primary state counts:', file=s)
        for pri_state, count in sorted(self.root_pri_to_count.items()):
            print('    ', pri_state, ':', count, file=s)
        print('edge-specific summaries:', file=s)
        for edge in self.edge_to_pri_trans:
            print('  edge', edge, ':', file=s)
            pri_trans = self.edge_to_pri_trans[edge]
            pri_dwell = self.edge_to_pri_dwell[edge]
            print('    trans: blinks:', file=s)
            print('      off -> on :', self.edge_to_off_xon_trans[edge], file=s)
            print('      unforced on -> off :', self.edge_to_xon_off_trans[edge], file=s)
            print('    trans: primary:', file=s)
            for sa in sorted(pri_trans):
                for sb in sorted(pri_trans[sa]):
                    print('      ', sa, '->', sb, ':', pri_trans[sa][sb]['weight'], file=s)
            print('    dwell: blinks:', file=s)
            print('      off -> on :', self.edge_to_off_xon_dwell[edge], file=s)
            print('      unforced on -> off :', self.edge_to_xon_off_dwell[edge], file=s)
            print('    dwell: primary:', file=s)
            for sa in sorted(pri_dwell):
                for sb in sorted(pri_dwell[sa]):
                    print('      ', sa, '->', sb, ':', pri_dwell[sa][sb]['weight'], file=s)
        return s.getvalue()
