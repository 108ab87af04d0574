"""Hard-coded toy blink models (the fixtures of nxblink/toymodel.py:80-241).

Six primary states in three tolerance classes (a "codon2x3" code), a six-node
tree rooted at N1.  The tables below restate the reference's three models; the
accessor names are the nine getters that nxblink/raoteh.py:33-44 requires.
"""
from __future__ import division, print_function

from .summary import WeightGraph

__all__ = ['BlinkModelA', 'BlinkModelB', 'BlinkModelC', 'Tree']


class Tree(dict):
    """Rooted tree as {node: [children]}; reads like the nx.DiGraph the reference uses."""

    def edges(self):
        return [(a, b) for a in self for b in self[a]]


_TREE_EDGES = (('N1', 'N0'), ('N1', 'N2'), ('N1', 'N5'), ('N2', 'N3'), ('N2', 'N4'))
_CODE = {0: 'T0', 1: 'T0', 2: 'T1', 3: 'T1', 4: 'T2', 5: 'T2'}

# (sa, sb): rate.  Model A: symmetric unit rates on the codon2x3 graph (toymodel.py:99-122).
_Q_A = dict(((a, b), 1.0) for a, b in (
    (0, 1), (0, 2), (1, 0), (1, 3), (2, 0), (2, 3), (2, 4),
    (3, 1), (3, 2), (3, 5), (4, 2), (4, 5), (5, 3), (5, 4)))
# Model B: 4<->5 removed, state 1 gets doubled in-rates and halved out-rates (toymodel.py:178-203).
_Q_B = {(0, 1): 2.0, (0, 2): 1.0, (1, 0): 0.5, (1, 3): 0.5, (2, 0): 1.0, (2, 3): 1.0,
        (2, 4): 1.0, (3, 1): 2.0, (3, 2): 1.0, (3, 5): 1.0, (4, 2): 1.0, (5, 3): 1.0}


class _BlinkModel(object):
    RATE_ON = RATE_OFF = None
    Q = None
    DISTN = None
    EDGE_RATES = None

    @classmethod
    def get_rate_on(cls):
        return cls.RATE_ON

    @classmethod
    def get_rate_off(cls):
        return cls.RATE_OFF

    @classmethod
    def get_blink_distn(cls):
        total = cls.RATE_ON + cls.RATE_OFF
        return {0: cls.RATE_OFF / total, 1: cls.RATE_ON / total}

    @classmethod
    def get_Q_blink(cls):
        Q = WeightGraph()
        Q.add_edge(0, 1, cls.RATE_ON)
        Q.add_edge(1, 0, cls.RATE_OFF)
        return Q

    @classmethod
    def get_primary_to_tol(cls):
        return dict(_CODE)

    @classmethod
    def get_T_and_root(cls):
        T = Tree()
        for a, b in _TREE_EDGES:
            T.setdefault(a, []).append(b)
            T.setdefault(b, [])
        return T, 'N1'

    @classmethod
    def get_edge_to_blen(cls):
        return dict((e, 1.0) for e in _TREE_EDGES)

    @classmethod
    def get_edge_to_rate(cls):
        return dict(zip(_TREE_EDGES, cls.EDGE_RATES))

    @classmethod
    def get_primary_distn(cls):
        return dict(cls.DISTN)

    @classmethod
    def get_Q_primary(cls):
        Q = WeightGraph()
        for (a, b), r in sorted(cls.Q.items()):
            Q.add_edge(a, b, r)
        return Q


class BlinkModelA(_BlinkModel):
    """Symmetric model (toymodel.py:80-133)."""
    RATE_ON = 1.0
    RATE_OFF = 1.0
    Q = _Q_A
    DISTN = dict((i, 1 / 6) for i in range(6))
    EDGE_RATES = (0.5, 0.5, 0.5, 0.5, 0.5)


class BlinkModelB(_BlinkModel):
    """Asymmetric model (toymodel.py:136-223)."""
    RATE_ON = 2.0
    RATE_OFF = 0.5
    Q = _Q_B
    DISTN = {0: 1 / 9, 1: 4 / 9, 2: 1 / 9, 3: 1 / 9, 4: 1 / 9, 5: 1 / 9}
    EDGE_RATES = (0.5, 0.5, 0.5, 1.0, 0.25)


class BlinkModelC(BlinkModelB):
    """Model B with two zero-rate branches (toymodel.py:226-241)."""
    EDGE_RATES = (0, 0, 0.5, 1.0, 0.25)
