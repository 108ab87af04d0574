"""Shapes the reference's fixtures never reach: a 150-node random tree with polytomies and
zero-length branches, 30 primary states in 8 classes; and the limits K = 32 classes / P = 64
states.  No exact answer exists, so the checks are the trajectory invariants after every phase,
the conservation laws, and prior invariance under the sweep (forward-simulated vs swept
statistics agree)."""
import numpy as np
import pytest

import golden_util as gu
from nxblink_b200.summary import WeightGraph
from nxblink_b200.toymodel import Tree
from nxblink_b200 import packing

pytestmark = pytest.mark.gpu


class BigModel(object):
    def __init__(self, n_nodes=150, P=30, K=8, seed=0, deg=3):
        rng = np.random.default_rng(seed)
        self.P, self.K = P, K
        self.code = dict((s, 'c%02d' % (s % K)) for s in range(P))
        w = rng.random(P) + 0.3
        self.distn = dict((i, float(x)) for i, x in enumerate(w / w.sum()))
        pairs = set((i, (i + 1) % P) for i in range(P))
        while len(pairs) < deg * P // 2 + P:
            a, b = rng.integers(0, P, size=2)
            if a != b and (int(b), int(a)) not in pairs:
                pairs.add((int(a), int(b)))
        self.Q = {}
        for a, b in sorted(pairs):
            sym = float(rng.random() + 0.2)
            self.Q[a, b] = sym * self.distn[b] * P
            self.Q[b, a] = sym * self.distn[a] * P
        self.edges, self.rates, self.blens = [], {}, {}
        for v in range(1, n_nodes):
            par = int(rng.integers(max(0, v - 12), v))      # polytomies and long chains both occur
            e = ('n%d' % par, 'n%d' % v)
            self.edges.append(e)
            self.rates[e] = 0.0 if rng.random() < 0.05 else float(rng.random() * 0.3)
            self.blens[e] = 0.0 if rng.random() < 0.03 else float(0.5 + rng.random())

    def get_rate_on(self):
        return 0.9

    def get_rate_off(self):
        return 1.7

    def get_blink_distn(self):
        return {0: 1.7 / 2.6, 1: 0.9 / 2.6}

    def get_primary_to_tol(self):
        return dict(self.code)

    def get_T_and_root(self):
        T = Tree()
        for a, b in self.edges:
            T.setdefault(a, []).append(b)
            T.setdefault(b, [])
        return T, 'n0'

    def get_edge_to_blen(self):
        return dict(self.blens)

    def get_edge_to_rate(self):
        return dict(self.rates)

    def get_primary_distn(self):
        return dict(self.distn)

    def get_Q_primary(self):
        Q = WeightGraph()
        for (a, b), r in sorted(self.Q.items()):
            Q.add_edge(a, b, r)
        return Q


def _prior_invariance(nb, M, n_units, sweeps):
    pm = packing.PackedModel(M)
    fwd, mcmc = [], []
    for seed in range(4):
        s = nb.BlinkSampler(pm, validate=True)
        s.forward_sample(n_units, seed=seed)
        s.reset_summary(); s.summarize_current()
        sm = s.summary()
        fwd.append(gu.scalar_functionals(sm))
        E, K = pm.n_nodes - 1, pm.n_classes
        np.testing.assert_allclose(sm.flat['T'].sum(axis=1), pm.blen[1:] * sm.nsamples, rtol=1e-10, atol=1e-9)
        s.reset_summary()
        s.sweep(sweeps, accumulate=True)
        sm = s.summary()
        assert sm.nsamples == n_units * sweeps
        assert sm.root_on_total + sm.root_off_count == K * sm.nsamples
        mcmc.append(gu.scalar_functionals(sm))
        st = s.stats()
        s.close()
    for name in fwd[0]:
        a = np.array([r[name] for r in fwd]); b = np.array([r[name] for r in mcmc])
        se = np.hypot(a.std(ddof=1), b.std(ddof=1)) / np.sqrt(len(a))
        assert abs(a.mean() - b.mean()) <= 5 * se + 3e-3 * max(1.0, abs(a.mean())), (name, a.mean(), b.mean(), se)
    return st


def test_large_irregular_tree(nb):
    st = _prior_invariance(nb, BigModel(), n_units=4096, sweeps=5)
    assert st['warps_per_cta'] >= 2


def test_maximum_classes_and_states(nb):
    """K = 32 tolerance classes (every lane a track, full-width masks), P = 64 states."""
    M = BigModel(n_nodes=20, P=64, K=32, seed=3, deg=2)
    _prior_invariance(nb, M, n_units=2048, sweeps=4)


class StarModel(BigModel):
    """A root with `n_leaves` long branches: the root's chunk collects one factor per branch."""

    def __init__(self, n_leaves=100, P=12, K=3, seed=5):
        BigModel.__init__(self, n_nodes=2, P=P, K=K, seed=seed, deg=3)
        self.edges, self.rates, self.blens = [], {}, {}
        for v in range(1, n_leaves + 1):
            e = ('n0', 'n%d' % v)
            self.edges.append(e)
            self.rates[e] = 1.0
            self.blens[e] = 1.5


def test_chunk_with_a_hundred_child_chunks_does_not_underflow_f32(nb):
    """The message of a chunk is the product of the factors its child chunks push up
    (chunking.py:264-289 -> nxmctree FFBS).  With 100 observed leaves on long branches the root
    chunk receives ~100 factors of order 1e-2: the raw product leaves the f32 range (round-1
    advisor finding: a spurious INFEASIBLE on feasible data).  The product is rescaled by a power
    of two every 4th fold; f32 and f64 statistics agree."""
    M = StarModel()
    pm = packing.PackedModel(M)
    rng = np.random.default_rng(1)
    n_sites = 6
    ls = np.full((n_sites, pm.n_nodes), -1, dtype=np.int64)
    ls[:, 1:] = rng.integers(0, pm.n_states, size=(n_sites, pm.n_nodes - 1))
    data = packing.pack_alignment(pm, ls)
    res = {}
    for f64 in (False, True):
        vals = []
        for seed in range(3):
            s = nb.BlinkSampler(pm, validate=True, msg_f64=f64)
            s.set_data(data)
            s.init_chains(chains=64, seed=seed)
            s.sweep(6, accumulate=False)          # raises NxbError(INFEASIBLE) if a message underflows
            s.sweep(6, accumulate=True)
            vals.append(gu.scalar_functionals(s.summary()))
            s.close()
        res[f64] = vals
    for name in res[False][0]:
        a = np.array([r[name] for r in res[False]]); b = np.array([r[name] for r in res[True]])
        se = np.hypot(a.std(ddof=1), b.std(ddof=1)) / np.sqrt(len(a))
        assert abs(a.mean() - b.mean()) <= 5 * se + 5e-3 * max(1.0, abs(a.mean())), (name, a.mean(), b.mean(), se)
