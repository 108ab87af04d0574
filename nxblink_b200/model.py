"""Functions specifying the blinking model (mirror of nxblink/model.py:22-106).

Rate matrices are nested dicts ``Q[sa][sb] = {'weight': rate}`` held in a
``WeightGraph``, which reads like the reference's ``nx.DiGraph``.
"""
from __future__ import division, print_function

from collections import defaultdict

from .summary import WeightGraph

__all__ = ['get_interaction_map', 'get_Q_blink', 'get_Q_meta']


def get_interaction_map(primary_to_part):
    """model.py:77-81 (golden: nxblink/tests/test_interaction_map.py:12-67).

    interaction['PRIMARY'][part][tol state] -> allowed primary states;
    interaction[part]['PRIMARY'][primary state] -> allowed tolerance states.
    """
    all_primary = set(primary_to_part)
    part_to_primary = defaultdict(set)
    for primary, part in primary_to_part.items():
        part_to_primary[part].add(primary)
    interaction = {'PRIMARY': {}}
    for part, members in part_to_primary.items():
        interaction['PRIMARY'][part] = {
                True: set(all_primary), False: all_primary - members}
        interaction[part] = {'PRIMARY': dict(
            (p, {True} if p in members else {False, True}) for p in all_primary)}
    return interaction


def get_Q_blink(rate_on=None, rate_off=None):
    """model.py:84-90."""
    Q = WeightGraph()
    Q.add_edge(False, True, rate_on)
    Q.add_edge(True, False, rate_off)
    return Q


def get_Q_meta(Q_primary, primary_to_tol):
    """model.py:93-106: rates from primary states into tolerance classes."""
    Q_meta = WeightGraph()
    for sa in Q_primary:
        for sb in Q_primary[sa]:
            w = Q_primary[sa][sb]
            rate = w['weight'] if isinstance(w, dict) else w
            tol = primary_to_tol[sb]
            if Q_meta.has_edge(sa, tol):
                Q_meta[sa][tol]['weight'] += rate
            else:
                Q_meta.add_edge(sa, tol, rate)
    return Q_meta
