"""The oracle (and the product's host-side mirrors) against the golden vectors
held by the reference's own tests."""
import numpy as np
from numpy.testing import assert_allclose, assert_equal

from oracle import blink_oracle as bo
from nxblink_b200 import model as product_model
from nxblink_b200 import maxlikelihood as product_ml


class _T(object):
    def __init__(self, name, history):
        self.name, self.history, self.events = name, history, {}


def test_gen_segments_known_answer():
    # nxblink/tests/test_navigation.py:11-39
    va, vb = 'A', 'B'
    edge = (va, vb)
    node_to_tm = {va: 4, vb: 10}
    ta = _T('ta', {va: 1, vb: 4})
    tb = _T('tb', {va: 'x', vb: 'x'})
    tc = _T('tc', {va: 'a', vb: 'c'})
    ta.events = {edge: [bo.Event(ta, 7, 3, 2), bo.Event(ta, 5.5, 1, 3), bo.Event(ta, 9, 2, 4)]}
    tb.events = {edge: []}
    tc.events = {edge: [bo.Event(tc, 5, 'a', 'b'), bo.Event(tc, 6, 'b', 'c')]}
    expected = [
        (4.0, 5.0, {'ta': 1, 'tb': 'x', 'tc': 'a'}),
        (5.0, 5.5, {'ta': 1, 'tb': 'x', 'tc': 'b'}),
        (5.5, 6.0, {'ta': 3, 'tb': 'x', 'tc': 'b'}),
        (6.0, 7.0, {'ta': 3, 'tb': 'x', 'tc': 'c'}),
        (7.0, 9.0, {'ta': 2, 'tb': 'x', 'tc': 'c'}),
        (9.0, 10.0, {'ta': 4, 'tb': 'x', 'tc': 'c'}),
    ]
    got = [(a, b, dict(s)) for a, b, s in bo.gen_segments(edge, node_to_tm, [ta, tb, tc])]
    assert_equal(got, expected)


# nxblink/tests/test_maxlikelihood.py:12-17 and :42-47
_ML_A = (6528, 6272, 15776, 16139, 15893.1803168, 16106.8196832)
_ML_B = (8811, 3989, 22025, 21989, 10069.0006186, 21930.9993814)


def _check_ml(mod):
    for stats, target in ((_ML_A, (1.0, 1.0)), (_ML_B, (2.22, 1.0))):
        on, off = mod.get_blink_rate_mle(*stats)
        xon, xoff = mod.get_blink_rate_analytical_mle(*stats)
        assert_allclose(on, xon, rtol=1e-7)
        assert_allclose(off, xoff, rtol=1e-7)
        assert_allclose(on, target[0], rtol=1e-1)
        assert_allclose(off, target[1], rtol=1e-1)
        # the closed form is a stationary point of the log likelihood
        g = mod.ll_blink_contribution_gradient(*(stats + (xon, xoff)))
        assert_allclose(g, 0, atol=1e-6)


def test_blink_rate_mle_oracle():
    _check_ml(bo)


def test_blink_rate_mle_product():
    _check_ml(product_ml)


def test_maxlikelihood_product_matches_oracle_to_1e10():
    for stats in (_ML_A, _ML_B):
        for rates in ((1.0, 1.0), (2.2, 0.9), (0.3, 4.0)):
            a = bo.ll_blink_contribution(*(stats + rates))
            b = product_ml.ll_blink_contribution(*(stats + rates))
            assert_allclose(b, a, rtol=1e-10)
            assert_allclose(product_ml.ll_blink_contribution_gradient(*(stats + rates)),
                    bo.ll_blink_contribution_gradient(*(stats + rates)), rtol=1e-10)
        assert_allclose(product_ml.get_blink_rate_analytical_mle(*stats),
                bo.get_blink_rate_analytical_mle(*stats), rtol=1e-10)


def _expected_interaction_map():
    # nxblink/tests/test_interaction_map.py:12-67
    allp = {'P0', 'P1', 'P2', 'P3', 'P4', 'P5'}
    code = {'P0': 'T0', 'P1': 'T0', 'P2': 'T1', 'P3': 'T1', 'P4': 'T2', 'P5': 'T2'}
    exp = {'PRIMARY': {}}
    for t in ('T0', 'T1', 'T2'):
        members = set(p for p, c in code.items() if c == t)
        exp['PRIMARY'][t] = {True: set(allp), False: allp - members}
        exp[t] = {'PRIMARY': dict((p, {True} if p in members else {False, True}) for p in allp)}
    return code, exp


def test_interaction_map_known_answer():
    code, exp = _expected_interaction_map()
    assert_equal(bo.get_interaction_map(code), exp)
    assert_equal(product_model.get_interaction_map(code), exp)


def test_q_meta_and_q_blink():
    from nxblink_b200 import toymodel
    Q = toymodel.BlinkModelB.get_Q_primary()
    code = toymodel.BlinkModelB.get_primary_to_tol()
    got = product_model.get_Q_meta(Q, code)
    exp = bo.get_Q_meta(bo.graph_to_dict(Q), code)
    for sa in exp:
        for tol, r in exp[sa].items():
            assert_allclose(got[sa][tol]['weight'], r, rtol=1e-15)
    qb = product_model.get_Q_blink(rate_on=2.0, rate_off=0.5)
    assert qb[False][True]['weight'] == 2.0 and qb[True][False]['weight'] == 0.5
