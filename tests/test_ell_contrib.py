"""Per-sample functionals of exported trajectories (nxblink/summary.py:247-348,
nxblink/navigation.py:14-46): the host mirrors against values the REFERENCE computed on its
own sampled trajectories (tests/golden/ref_ell_contrib.json, made by
oracle/refcompat/make_golden.py ell) and against the reference's known-answer vector."""
import json
import os

import pytest
from numpy.testing import assert_allclose, assert_equal

from nxblink_b200 import toymodel, model as nbmodel
from nxblink_b200.sampler import Track, Event
from nxblink_b200.summary import (WeightGraph, gen_segments, get_ell_init_contrib,
        get_ell_dwell_contrib, get_ell_trans_contrib)

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'ref_ell_contrib.json')


def _tracks(rec):
    out = []
    for t in rec['tracks']:
        tr = Track(t['name'])
        tr.history = dict(t['history'])
        tr.events = dict((tuple(e), [Event(tr, tm, sa, sb) for tm, sa, sb in evs])
                for e, evs in t['events'])
        out.append(tr)
    return out[0], out[1:]


def test_gen_segments_known_answer():
    # nxblink/tests/test_navigation.py:11-39
    edge = ('A', 'B')
    node_to_tm = {'A': 4, 'B': 10}
    ta, tb, tc = Track('ta'), Track('tb'), Track('tc')
    ta.history, tb.history, tc.history = {'A': 1, 'B': 4}, {'A': 'x', 'B': 'x'}, {'A': 'a', 'B': 'c'}
    ta.events = {edge: [Event(ta, 7, 3, 2), Event(ta, 5.5, 1, 3), Event(ta, 9, 2, 4)]}
    tb.events = {edge: []}
    tc.events = {edge: [Event(tc, 5, 'a', 'b'), Event(tc, 6, 'b', 'c')]}
    expected = [
        (4.0, 5.0, {'ta': 1, 'tb': 'x', 'tc': 'a'}),
        (5.0, 5.5, {'ta': 1, 'tb': 'x', 'tc': 'b'}),
        (5.5, 6.0, {'ta': 3, 'tb': 'x', 'tc': 'b'}),
        (6.0, 7.0, {'ta': 3, 'tb': 'x', 'tc': 'c'}),
        (7.0, 9.0, {'ta': 2, 'tb': 'x', 'tc': 'c'}),
        (9.0, 10.0, {'ta': 4, 'tb': 'x', 'tc': 'c'}),
    ]
    got = [(a, b, dict(s)) for a, b, s in gen_segments(edge, node_to_tm, [ta, tb, tc])]
    assert_equal(got, expected)


def test_gen_segments_rejects_incompatible_transition():
    edge = ('A', 'B')
    t = Track('t')
    t.history = {'A': 0, 'B': 1}
    t.events = {edge: [Event(t, 1.0, 1, 0)]}
    with pytest.raises(Exception):
        list(gen_segments(edge, {'A': 0.0, 'B': 2.0}, [t]))


@pytest.mark.parametrize('key', ['BlinkModelB/DataC', 'BlinkModelC/DataA'])
def test_contribs_equal_the_reference(key):
    with open(GOLDEN) as f:
        fx = json.load(f)[key]
    M = getattr(toymodel, key.split('/')[0])
    T, root = M.get_T_and_root()
    blen = M.get_edge_to_blen()
    node_to_tm = {root: 0.0}
    for a, b in T.edges():          # parents precede children in the toy tree's edge order
        node_to_tm[b] = node_to_tm[a] + blen[(a, b)]
    Q_primary = WeightGraph()
    for a, b, w in fx['Q_primary']:
        Q_primary.add_edge(a, b, w)
    primary_to_tol = M.get_primary_to_tol()
    Q_meta = nbmodel.get_Q_meta(Q_primary, primary_to_tol)
    Q_blink = M.get_Q_blink()
    edge_to_rate = M.get_edge_to_rate()
    for rec in fx['samples']:
        pri, tols = _tracks(rec)
        got = get_ell_init_contrib(root, M.get_primary_distn(), M.get_blink_distn(), pri, tols,
                primary_to_tol)
        assert_allclose(got, rec['ell_init'], rtol=1e-13)
        d = get_ell_dwell_contrib(T, root, node_to_tm, edge_to_rate, Q_primary, Q_blink, Q_meta,
                pri, tols, primary_to_tol)
        for e, want in rec['ell_dwell']:
            assert_allclose(d[tuple(e)], want, rtol=1e-12, atol=1e-14)
        tr = get_ell_trans_contrib(T, root, edge_to_rate, Q_primary, Q_blink, pri, tols)
        for e, want in rec['ell_trans']:
            assert_allclose(tr[tuple(e)], want, rtol=1e-12, atol=1e-14)
        e0 = tuple(rec['ell_dwell'][0][0])
        segs = [[a, b, sorted((str(k), int(v)) for k, v in st.items())]
                for a, b, st in gen_segments(e0, node_to_tm, [pri] + tols)]
        assert len(segs) == len(rec['segments_edge0'])
        for g, w in zip(segs, rec['segments_edge0']):
            assert g[:2] == w[:2]
            assert [list(x) for x in g[2]] == w[2]
