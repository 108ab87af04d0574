"""Vanilla (non-CTBN) Rao-Teh sampling of the explicit compound process: every
(primary state, tolerance bits) combination is one state of a single track.

Mirror of nxblink/compoundb.py:23-296.  The reference enumerates the compound states,
builds their rate matrix and hands both to ``nxctmctree.raoteh.gen_raoteh_trajectories``
(an un-vendored dependency); here the same chain is sampled by the device sweep kernel:
a one-track Rao-Teh chain is the blink model with a single tolerance class that every
state belongs to (the class is then forced ON everywhere and its track never moves), so
``BlinkSampler`` runs it unchanged.  Because none of the tolerance-update code takes part,
its statistics are an independent statistical cross-check of the CTBN sampler at toy size
(at most 64 compound states).
"""
from __future__ import division, print_function

import numpy as np

from .dense import CompoundProcess
from .packing import PackedModel, PackedData
from .sampler import BlinkSampler, Track, Event
from .summary import WeightGraph

__all__ = ['State', 'gen_states', 'get_Q', 'state_is_consistent', 'gen_successors',
        'CompoundSampler', 'gen_raoteh_samples']


class State(object):
    """The reference's compound state (compoundb.py:23-41: a primary state and one 0/1 state per
    tolerance class), kept here in packed form: the tolerance states are the bits of ``mask``.
    Same public surface -- ``pri``, ``tol`` (a list), ``copy()``, ``tolist()``, ordering and
    hashing by ``[pri] + tol`` -- so it can stand in wherever the reference's objects are used."""
    __slots__ = ('pri', 'mask', 'ntol')

    def __init__(self, pri, tol):
        bits = [1 if x else 0 for x in tol]
        self.pri, self.ntol = pri, len(bits)
        self.mask = sum(bit << k for k, bit in enumerate(bits))

    @classmethod
    def packed(cls, pri, mask, ntol):
        s = cls.__new__(cls)
        s.pri, s.mask, s.ntol = pri, mask, ntol
        return s

    @property
    def tol(self):
        return [(self.mask >> k) & 1 for k in range(self.ntol)]

    def flipped(self, k):
        """The state with tolerance class k switched."""
        return State.packed(self.pri, self.mask ^ (1 << k), self.ntol)

    def copy(self):
        return State.packed(self.pri, self.mask, self.ntol)

    def tolist(self):
        return [self.pri] + self.tol

    def _key(self):
        return (self.pri, tuple(self.tol))

    def __lt__(self, other):
        return self._key() < other._key()

    def __eq__(self, other):
        return isinstance(other, State) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return 'State(%r, %r)' % (self.pri, self.tol)


def state_is_consistent(pri_to_tol, s):
    """A compound state exists only with the class of its primary state switched on
    (compoundb.py:150-157)."""
    return (s.mask >> pri_to_tol[s.pri]) & 1 == 1


def _n_classes(pri_to_tol):
    classes = set(pri_to_tol.values())
    if classes != set(range(len(classes))):
        raise Exception('tolerance classes must be numbered 0..ntol-1')
    return len(classes)


def gen_states(pri_to_tol):
    """Every consistent compound state (compoundb.py:114-124), tolerance bits counting up with
    the LAST class fastest, as ``itertools.product((0, 1), repeat=ntol)`` orders them."""
    ntol = _n_classes(pri_to_tol)
    for pri in pri_to_tol:
        own = 1 << pri_to_tol[pri]
        for code in range(1 << ntol):
            # bit k of `mask` is digit k of the product tuple = bit (ntol - 1 - k) of `code`
            mask = sum(((code >> (ntol - 1 - k)) & 1) << k for k in range(ntol))
            if mask & own:
                yield State.packed(pri, mask, ntol)


def gen_successors(Q_pri, on_rate, off_rate, pri_to_tol, s):
    """One-step moves of the compound chain with their rates (compoundb.py:160-193): a primary
    transition into a state whose class is on, or one tolerance class switching (never the class
    of the current primary state off)."""
    for pri, w in Q_pri[s.pri].items():
        if (s.mask >> pri_to_tol[pri]) & 1:
            yield State.packed(pri, s.mask, s.ntol), (w['weight'] if isinstance(w, dict) else w)
    own = pri_to_tol[s.pri]
    for k in range(s.ntol):
        is_on = (s.mask >> k) & 1
        if is_on and k == own:
            continue
        yield s.flipped(k), (off_rate if is_on else on_rate)


def get_Q(Q_pri, on_rate, off_rate, pri_to_tol):
    """The compound rate matrix as a weighted digraph over ``State`` objects (compoundb.py:127-134)."""
    Q = WeightGraph()
    for a in gen_states(pri_to_tol):
        for b, rate in gen_successors(Q_pri, on_rate, off_rate, pri_to_tol, a):
            Q.add_edge(a, b, rate)
    return Q


class _ChainModel(object):
    """The compound chain presented through the reference's model getters."""

    def __init__(self, cp):
        pm = cp.pm
        n = len(cp.states)
        off = cp.Q - np.diag(np.diag(cp.Q))
        self._Q = WeightGraph()
        for i, j in zip(*np.nonzero(off)):
            self._Q.add_edge(int(i), int(j), float(off[i, j]))
        # raoteh.py:90-102 rescales the rate matrix it is given to expected rate 1;
        # the edge rates carry the factor back
        er = float(cp.prior.dot(off.sum(axis=1)))
        self._edge_to_rate = dict((e, float(pm.rate[i + 1]) * er) for i, e in enumerate(pm.edges))
        self._edge_to_blen = dict((e, float(pm.blen[i + 1])) for i, e in enumerate(pm.edges))
        self._distn = dict((i, float(cp.prior[i])) for i in range(n))
        self._pm = pm
        self._n = n

    def get_T_and_root(self):
        return self._pm.T, self._pm.root

    def get_edge_to_blen(self):
        return self._edge_to_blen

    def get_edge_to_rate(self):
        return self._edge_to_rate

    def get_primary_to_tol(self):
        return dict((i, 0) for i in range(self._n))

    def get_Q_primary(self):
        return self._Q

    def get_primary_distn(self):
        return self._distn

    def get_rate_on(self):
        return 1.0

    def get_rate_off(self):
        return 1.0

    def get_blink_distn(self):
        return {0: 0.5, 1: 0.5}


class CompoundSampler(object):
    """Batched one-track Rao-Teh chains of the compound process on the device."""

    def __init__(self, model, device=0, **kwargs):
        self.cp = CompoundProcess(model)
        self.chain_pm = PackedModel(_ChainModel(self.cp))
        # PackedModel numbers nodes by subtree size only, so both packings agree
        assert self.chain_pm.nodes == self.cp.pm.nodes
        self.sampler = BlinkSampler(self.chain_pm, device=device, **kwargs)

    def set_data(self, packed_data):
        """packed_data: per-track masks of the blink model (packing.pack_data)."""
        m = self.cp.masks(packed_data)
        self.sampler.set_data(PackedData(m, np.ones(m.shape, np.uint32), np.zeros(m.shape, np.uint32)))

    def init_chains(self, chains=1, seed=0):
        self.sampler.init_chains(chains=chains, seed=seed)

    def sweep(self, n=1, accumulate=True):
        self.sampler.sweep(n, accumulate=accumulate)

    def close(self):
        self.sampler.close()

    def blink_summary(self):
        """Fold the compound-state statistics into the blink model's sufficient statistics
        (what compoundb.Summary.on_track, compoundb.py:44-111, accumulates), per sample and
        keyed like ``dense.CompoundProcess.expected_summary``."""
        cp, pm, cm = self.cp, self.cp.pm, self.chain_pm
        sm = self.sampler.summary()
        n = float(sm.nsamples)
        K = pm.n_classes
        pri = np.array([p for p, b in cp.states])
        bits = np.array([b for p, b in cp.states])
        nbits = bits.sum(axis=1)
        T, tr, root = sm.flat['T'], sm.flat['pri_trans'], sm.flat['root_pri']
        ia, ib = cm.q_row, cm.q_col              # chain states are the compound indices
        out = dict(nsamples=sm.nsamples,
            root_pri_to_count=dict((pm.states[p], float(root[pri == p].sum()) / n)
                for p in range(pm.n_states)),
            root_on_total=float(root.dot(nbits)) / n, root_off_count=float(root.dot(K - nbits)) / n,
            root_xon_intended=float(root.dot(nbits - 1)) / n,
            edge_to_pri_trans={}, edge_to_pri_dwell={}, edge_to_off_xon_trans={},
            edge_to_xon_off_trans={}, edge_to_off_xon_dwell={}, edge_to_xon_off_dwell={})
        moved = pri[ia] != pri[ib]
        up = ~moved & (nbits[ib] > nbits[ia])
        down = ~moved & (nbits[ib] < nbits[ia])
        for e, edge in enumerate(pm.edges):
            t, d = {}, {}
            for j in range(pm.nnz):
                a, b = pm.q_row[j], pm.q_col[j]
                sel = moved & (pri[ia] == a) & (pri[ib] == b)
                t.setdefault(pm.states[a], {})[pm.states[b]] = float(tr[e][sel].sum()) / n
                d.setdefault(pm.states[a], {})[pm.states[b]] = float(
                        T[e][(pri == a) & (bits[:, pm.state_class[b]] == 1)].sum()) / n
            out['edge_to_pri_trans'][edge] = t
            out['edge_to_pri_dwell'][edge] = d
            out['edge_to_off_xon_trans'][edge] = float(tr[e][up].sum()) / n
            out['edge_to_xon_off_trans'][edge] = float(tr[e][down].sum()) / n
            out['edge_to_off_xon_dwell'][edge] = float(T[e].dot(K - nbits)) / n
            out['edge_to_xon_off_dwell'][edge] = float(T[e].dot(nbits - 1)) / n
        return out

    def export_track(self, unit):
        """One compound track: ``history[node]`` and ``events[edge]`` hold ``State`` objects."""
        pm = self.cp.pm
        pri, tols = self.sampler.export_tracks(unit)
        label = [State(pm.states[p], b) for p, b in self.cp.states]
        tr = Track('COMPOUND')
        tr.history = dict((v, label[s]) for v, s in pri.history.items())
        tr.events = dict((e, [Event(tr, ev.tm, label[ev.sa], label[ev.sb]) for ev in evs])
                for e, evs in pri.events.items())
        return tr


def gen_raoteh_samples(model, data, nburnin, nsamples, seed=None, device=0):
    """compoundb.py:196-296: yield ``nsamples`` compound tracks after ``nburnin`` sweeps
    (one site, one chain, every sweep a sample)."""
    from .packing import pack_data
    cs = CompoundSampler(model, device=device)
    try:
        cs.set_data(pack_data(cs.cp.pm, [data]))
        cs.init_chains(chains=1, seed=np.random.randint(1 << 30) if seed is None else seed)
        cs.sweep(nburnin, accumulate=False)
        for _ in range(nsamples):
            cs.sweep(1, accumulate=False)
            yield cs.export_track(0)
    finally:
        cs.close()
