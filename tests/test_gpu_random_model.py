"""A randomly generated model that exercises what the toy grid does not: more than 32
primary states (both halves of the state lanes), uneven class sizes, a deeper unbalanced
tree with a zero-rate edge, observed and unobserved leaves and a forced-off class -- checked
statistic by statistic against the EXACT compound-process posterior (160 compound states,
oracle/dense_oracle.py; fixture tests/golden/exact_random40.json)."""
import ast
import json
import os

import numpy as np
import pytest

from oracle.dense_oracle import DensePosterior
from nxblink_b200.summary import WeightGraph
from nxblink_b200.toymodel import Tree
from test_gpu_toy_exact import _flatten, _flatten_exact

pytestmark = pytest.mark.gpu


class RandomModel(object):
    def __init__(self, seed=0, P=40, K=3):
        rng = np.random.default_rng(seed)
        self.P, self.K = P, K
        sizes = [20, 14, 6]
        self.code = {}
        s = 0
        for k, n in enumerate(sizes):
            for _ in range(n):
                self.code[s] = 'C%d' % k
                s += 1
        w = rng.random(P) + 0.2
        self.distn = dict((i, float(x)) for i, x in enumerate(w / w.sum()))
        # reversible sparse rates: ring + random chords, q_ab = s_ab * pi_b
        pairs = set((i, (i + 1) % P) for i in range(P))
        while len(pairs) < 2 * P:
            a, b = rng.integers(0, P, size=2)
            if a != b and (b, a) not in pairs:
                pairs.add((int(a), int(b)))
        self.Q = {}
        for a, b in sorted(pairs):
            sym = float(rng.random() + 0.3)
            self.Q[a, b] = sym * self.distn[b] * P
            self.Q[b, a] = sym * self.distn[a] * P
        self.edges = [('r', 'a'), ('r', 'b'), ('a', 'c'), ('a', 'd'), ('c', 'e'), ('c', 'f'),
                      ('e', 'g'), ('e', 'h')]
        self.rates = dict(zip(self.edges, [0.3, 0.7, 0.0, 0.5, 0.4, 0.2, 0.6, 0.35]))
        self.blens = dict(zip(self.edges, [1.0, 0.5, 1.0, 2.0, 1.0, 1.5, 1.0, 1.0]))

    def get_rate_on(self):
        return 1.3

    def get_rate_off(self):
        return 0.8

    def get_blink_distn(self):
        return {0: 0.8 / 2.1, 1: 1.3 / 2.1}

    def get_primary_to_tol(self):
        return dict(self.code)

    def get_T_and_root(self):
        T = Tree()
        for a, b in self.edges:
            T.setdefault(a, []).append(b)
            T.setdefault(b, [])
        return T, 'r'

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


class RandomData(object):
    def __init__(self, model):
        self.m = model

    def get_data(self):
        m = self.m
        nodes = ['r', 'a', 'b', 'c', 'd', 'e', 'f', 'g', 'h']
        allp = set(range(m.P))
        pri = dict((v, set(allp)) for v in nodes)
        pri['b'] = {3}           # class C0
        pri['d'] = {25}          # class C1
        pri['g'] = {36, 37, 5}   # ambiguous observation
        pri['h'] = {22}
        data = {'PRIMARY': pri}
        for name in ('C0', 'C1', 'C2'):
            data[name] = dict((v, {False, True}) for v in nodes)
        data['C0']['b'] = {True}
        data['C1']['d'] = {True}
        data['C1']['h'] = {True}
        data['C2']['b'] = {False}     # "lethal" class at leaf b
        data['C2']['f'] = {True}
        return data


def test_forty_state_model_against_exact_posterior(nb):
    M = RandomModel()
    D = RandomData(M)
    exact = None
    runs = []
    for seed in range(6):
        s = nb.BlinkSampler(M, validate=True)
        s.set_data([D])
        s.init_chains(chains=2048, seed=77 + seed)
        s.sweep(60, accumulate=False)
        s.sweep(40, accumulate=True)
        sm = s.summary()
        if exact is None:
            # precomputed by tests/golden/make_exact_random40.py (4 minutes of expm_frechet)
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'exact_random40.json')
            with open(path) as f:
                exact = dict((ast.literal_eval(k), v) for k, v in json.load(f)['values'])
        runs.append(_flatten(sm, s.pm))
        s.close()
    worst = (0.0, None)
    for key, ex in exact.items():
        vals = np.array([r[key] for r in runs])
        se = vals.std(ddof=1) / np.sqrt(len(runs))
        z = abs(vals.mean() - ex) / (5 * se + 2e-3)
        if z > worst[0]:
            worst = (z, (key, vals.mean(), ex, se))
    assert worst[0] <= 1.0, worst
