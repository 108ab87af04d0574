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

import math
from collections import defaultdict
from io import StringIO

import numpy as np

__all__ = ['WeightGraph', 'Summary', 'gen_segments', 'get_ell_init_contrib',
        'get_ell_dwell_contrib', 'get_ell_trans_contrib', 'mean_ell_contribs']


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
        self._pri_trans = {}
        self.edge_to_off_xon_trans = defaultdict(int)
        self.edge_to_xon_off_trans = defaultdict(int)
        self._pri_dwell = {}
        self._lazy = None
        self.edge_to_off_xon_dwell = defaultdict(float)
        self.edge_to_xon_off_dwell = defaultdict(float)
        self.flat = None

    def _unfold(self):
        pm, pri_trans, pri_dwell = self._lazy
        self._lazy = None
        rows, cols = np.nonzero((pri_trans != 0) | (pri_dwell > 0))
        for edge in pm.edges:
            self._pri_trans[edge] = WeightGraph()
            self._pri_dwell[edge] = WeightGraph()
        for e, i in zip(rows.tolist(), cols.tolist()):
            edge = pm.edges[e]
            sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
            if pri_trans[e, i]:
                self._pri_trans[edge].add_edge(sa, sb, int(pri_trans[e, i]))
            if pri_dwell[e, i] > 0:
                self._pri_dwell[edge].add_edge(sa, sb, float(pri_dwell[e, i]))

    @property
    def edge_to_pri_trans(self):
        """{edge: graph of primary transition counts} (summary.py:102)."""
        if self._lazy is not None:
            self._unfold()
        return self._pri_trans

    @property
    def edge_to_pri_dwell(self):
        """{edge: graph of primary dwell times by allowed transition} (summary.py:107)."""
        if self._lazy is not None:
            self._unfold()
        return self._pri_dwell

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
        # the per-edge tables of the reference (nx.DiGraph-like) are unfolded on first access
        self._lazy = (pm, pri_trans, pri_dwell)
        for e, edge in enumerate(pm.edges):
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
            for tm, tm_next, state in gen_segments(edge, self._node_to_tm, tracks):
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


# -- per-sample functionals of exported tracks ---------------------------------------------------

def gen_segments(edge, node_to_tm, tracks):
    """navigation.py:14-46 (golden vector: nxblink/tests/test_navigation.py:30-37): the maximal
    pieces of an edge on which no track changes, as ``(tm, tm_next, {track name: state})``.
    Like the reference, the SAME mutable dict is yielded every time."""
    va, vb = edge
    state = dict((t.name, t.history[va]) for t in tracks)
    seq = sorted(((ev.tm, i, ev) for i, ev in enumerate(
        ev for t in tracks for ev in t.events[edge])), key=lambda x: x[:2])
    tm = node_to_tm[va]
    for tm_next, i, ev in seq:
        yield tm, tm_next, state
        if state[ev.track.name] != ev.sa:
            raise Exception('incompatible transition')
        state[ev.track.name] = ev.sb
        tm = tm_next
    yield tm, node_to_tm[vb], state


def _rate(G, a, b):
    """G[a][b]['weight'] of an nx.DiGraph-like table, or None."""
    row = G[a] if a in G else None
    if row is None or b not in row:
        return None
    w = row[b]
    return w['weight'] if isinstance(w, dict) else w


def _edges_of(T):
    return list(T.edges()) if hasattr(T, 'edges') else [(a, b) for a in T for b in T[a]]


def get_ell_init_contrib(root, primary_distn, blink_distn,
        primary_track, tolerance_tracks, primary_to_tol):
    """summary.py:247-261: log prior of the root state of one sampled trajectory (the own
    class of the root's primary state is forced ON and contributes nothing)."""
    primary_state = primary_track.history[root]
    own = primary_to_tol[primary_state]
    ll = math.log(primary_distn[primary_state])
    for t in tolerance_tracks:
        if t.name != own:
            ll += math.log(blink_distn[t.history[root]])
    return ll


def get_ell_dwell_contrib(T, root, node_to_tm, edge_to_rate, Q_primary, Q_blink, Q_meta,
        primary_track, tolerance_tracks, primary_to_tol):
    """summary.py:265-317: per edge, minus (edge rate x total exit rate x time) summed over the
    all-constant segments of one sampled trajectory.  ``Q_primary`` is used as given (the
    reference does not rescale it here: SURVEY quirk 5)."""
    tracks = [primary_track] + list(tolerance_tracks)
    rate_off, rate_on = _rate(Q_blink, True, False), _rate(Q_blink, False, True)
    out = {}
    for edge in _edges_of(T):
        contrib = 0
        for tma, tmb, state in gen_segments(edge, node_to_tm, tracks):
            pri = state[primary_track.name]
            own = primary_to_tol[pri]
            rate = 0
            for t in tolerance_tracks:
                into = _rate(Q_meta, pri, t.name)
                if t.name == own:
                    if into is not None:
                        rate += into
                elif state[t.name]:
                    rate += rate_off
                    if into is not None:
                        rate += into
                else:
                    rate += rate_on
            contrib += -(edge_to_rate[edge] * rate * (tmb - tma))
        out[edge] = contrib
    return out


def get_ell_trans_contrib(T, root, edge_to_rate, Q_primary, Q_blink,
        primary_track, tolerance_tracks):
    """summary.py:321-348: per edge, the sum of log(edge rate x transition rate) over the
    events of one sampled trajectory."""
    out = {}
    for edge in _edges_of(T):
        contrib = 0
        for ev in primary_track.events[edge]:
            rate = _rate(Q_primary, ev.sa, ev.sb)
            if rate is None:
                raise Exception('primary transition that the rate matrix does not have')
            contrib += math.log(edge_to_rate[edge] * rate)
        for t in tolerance_tracks:
            for ev in t.events[edge]:
                if bool(ev.sa) == bool(ev.sb):
                    raise Exception('blink self transition')
                contrib += math.log(edge_to_rate[edge] * _rate(Q_blink, ev.sa, ev.sb))
        out[edge] = contrib
    return out


def mean_ell_contribs(summary, pm, edge_to_rate, Q_primary, rate_on, rate_off,
        primary_distn, blink_distn):
    """Per-sample AVERAGES of the three functionals above, from the statistics the sweep
    kernel reduced on the device -- what examples/codon2x3/toymodel-stochastic.py:137-196
    obtains by calling them on every sample in a Python loop.

    summary : ``Summary.from_device`` result; pm : its ``PackedModel``.
    Returns ``(ell_init, {edge: ell_dwell}, {edge: ell_trans})``.
    """
    f, n = summary.flat, float(summary.nsamples)
    P, K = pm.n_states, pm.n_classes
    q = np.array([_rate(Q_primary, pm.states[a], pm.states[b]) or 0.0
        for a, b in zip(pm.q_row, pm.q_col)], dtype=float)
    qmeta = np.zeros((P, K))
    np.add.at(qmeta, (pm.q_row, pm.state_class[pm.q_col]), q)
    rates = np.array([edge_to_rate[e] for e in pm.edges], dtype=float)
    # exit rate of a segment = sum_k Qmeta[s][k] [k ON] + off * #(non-own ON) + on * #OFF
    exit_time = (f['T'] * qmeta.sum(axis=1)[None, :]).sum(axis=1) \
            - (f['Doff'] * qmeta[None, :, :]).sum(axis=(1, 2)) \
            + rate_off * f['xon_off_dwell'] + rate_on * f['off_xon_dwell']
    dwell = -rates * exit_time / n
    with np.errstate(divide='ignore'):
        logq, logr = np.log(q), np.log(rates)
    ntrans = f['pri_trans'].sum(axis=1) + f['off_on'] + f['on_off']
    trans = np.where(ntrans > 0, ntrans * np.where(ntrans > 0, logr, 0.0), 0.0) \
            + np.where(f['pri_trans'] > 0, f['pri_trans'] * np.where(f['pri_trans'] > 0, logq, 0.0),
                    0.0).sum(axis=1) \
            + f['off_on'] * math.log(rate_on) + f['on_off'] * math.log(rate_off)
    trans = trans / n
    pd = np.array([primary_distn.get(st, 0.0) if hasattr(primary_distn, 'get') else primary_distn[st]
        for st in pm.states], dtype=float)
    with np.errstate(divide='ignore'):
        lp = np.where(f['root_pri'] > 0, f['root_pri'] * np.log(np.where(pd > 0, pd, 1.0)), 0.0)
    init = (lp.sum() + summary.root_off_count * math.log(blink_distn[0])
            + summary.root_xon_intended * math.log(blink_distn[1])) / n
    return (init, dict(zip(pm.edges, dwell.tolist())), dict(zip(pm.edges, trans.tolist())))
