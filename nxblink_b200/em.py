"""Expectation-maximisation objective pieces (mirror of nxblink/em.py:15-198).

Same names and argument order as the reference.  The reference evaluates these
through ``algopy`` so that AD types can flow through; here they are plain
numpy (values), plus vectorised variants on the flat device statistics
(``*_flat``) that avoid the per-edge Python loops.
"""
from __future__ import division, print_function

import numpy as np

__all__ = ['get_ll_dwell', 'get_ll_trans', 'get_ll_root', 'get_expected_rate']


def _items(G):
    for sa in G:
        for sb in G[sa]:
            w = G[sa][sb]
            yield sa, sb, (w['weight'] if isinstance(w, dict) else w)


def get_expected_rate(pre_Q, distn):
    """em.py:15-16."""
    return np.dot(np.asarray(pre_Q).sum(axis=1), distn)


def get_ll_dwell(summary, pre_Q, distn, blink_on, blink_off, edges, edge_rates):
    """em.py:19-83."""
    expected_rate = get_expected_rate(pre_Q, distn)
    ll = 0.0
    for edge, edge_rate in zip(edges, edge_rates):
        ll_dwell = 0.0
        for sa, sb, elapsed in _items(summary.edge_to_pri_dwell[edge]):
            if elapsed:
                ll_dwell = ll_dwell + elapsed * pre_Q[sa, sb]
        off_xon_dwell = summary.edge_to_off_xon_dwell[edge]
        xon_off_dwell = summary.edge_to_xon_off_dwell[edge]
        if off_xon_dwell:
            ll_dwell = ll_dwell + off_xon_dwell * blink_on
        if xon_off_dwell:
            ll_dwell = ll_dwell + xon_off_dwell * blink_off
        ll = ll - ll_dwell * (edge_rate / expected_rate)
    return ll / summary.nsamples


def get_ll_trans(summary, pre_Q, distn, blink_on, blink_off, edges, edge_rates):
    """em.py:86-154."""
    expected_rate = get_expected_rate(pre_Q, distn)
    ll = 0.0
    for edge, edge_rate in zip(edges, edge_rates):
        transition_count_sum = 0
        for sa, sb, count in _items(summary.edge_to_pri_trans[edge]):
            transition_count_sum += count
            if count:
                ll = ll + count * np.log(pre_Q[sa, sb])
        off_xon_trans = summary.edge_to_off_xon_trans[edge]
        xon_off_trans = summary.edge_to_xon_off_trans[edge]
        if off_xon_trans:
            transition_count_sum += off_xon_trans
            ll = ll + off_xon_trans * np.log(blink_on)
        if xon_off_trans:
            transition_count_sum += xon_off_trans
            ll = ll + xon_off_trans * np.log(blink_off)
        if transition_count_sum:
            ll = ll + transition_count_sum * np.log(edge_rate / expected_rate)
    return ll / summary.nsamples


def get_ll_root(summary, distn, blink_on, blink_off):
    """em.py:157-198 (uses ``root_xon_count`` exactly as the reference does)."""
    p_off = blink_off / (blink_on + blink_off)
    p_on = blink_on / (blink_on + blink_off)
    ll = 0.0
    for state, count in summary.root_pri_to_count.items():
        if count:
            ll = ll + count * np.log(distn[state])
    if summary.root_off_count:
        ll = ll + summary.root_off_count * np.log(p_off)
    if summary.root_xon_count:
        ll = ll + summary.root_xon_count * np.log(p_on)
    return ll / summary.nsamples
