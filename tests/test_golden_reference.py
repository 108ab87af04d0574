"""Fixtures produced by the REFERENCE ITSELF (its own raoteh / summary / em /
maxlikelihood files executed through oracle/refcompat; generator:
oracle/refcompat/make_golden.py) pin the oracle and the product's host-side
deterministic functions.

Deterministic quantities: <= 1e-10 relative.  Sampled quantities: Monte Carlo
error (5 combined standard errors, batch means)."""
import ast

import numpy as np
import pytest
from numpy.testing import assert_allclose

import golden_util as gu
from oracle import blink_oracle as bo
from oracle.dense_oracle import DensePosterior
from nxblink_b200 import em, maxlikelihood, toymodel, toydata

TOY = gu.load('ref_toy.json')
MODELS = dict((m.__name__, m) for m in (toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC))
DATA = dict((d.__name__, d) for d in (toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD))


def _model_arrays(M):
    Q = M.get_Q_primary()
    pre_Q = np.zeros((6, 6))
    for a, b in Q.edges():
        pre_Q[a, b] = Q[a][b]['weight']
    pd = M.get_primary_distn()
    return pre_Q, np.array([pd[i] for i in range(6)])


@pytest.mark.parametrize('key', sorted(TOY))
def test_em_and_maxlikelihood_match_reference_values(key):
    rec = TOY[key]
    M = MODELS[key.split('/')[0]]
    summary = gu.summary_from_record(rec)
    pre_Q, distn = _model_arrays(M)
    on, off = M.get_rate_on(), M.get_rate_off()
    edges = [(ast.literal_eval(a), ast.literal_eval(b)) for a, b in rec['em']['edges']]
    rates = rec['em']['edge_rates']
    for mod in (em, bo):
        assert_allclose(mod.get_ll_root(summary, distn, on, off), rec['em']['ll_root'], rtol=1e-10)
        assert_allclose(mod.get_ll_dwell(summary, pre_Q, distn, on, off, edges, rates),
                rec['em']['ll_dwell'], rtol=1e-10)
        assert_allclose(mod.get_ll_trans(summary, pre_Q, distn, on, off, edges, rates),
                rec['em']['ll_trans'], rtol=1e-10)
    six = tuple(rec['maxlikelihood']['six'])
    for mod in (maxlikelihood, bo):
        assert_allclose(mod.ll_blink_contribution(*(six + (on, off))),
                rec['maxlikelihood']['ll_at_model'], rtol=1e-10)
        assert_allclose(mod.ll_blink_contribution_gradient(*(six + (on, off))),
                rec['maxlikelihood']['grad_at_model'], rtol=1e-10)
        assert_allclose(mod.get_blink_rate_analytical_mle(*six),
                rec['maxlikelihood']['analytic_mle'], rtol=1e-10)
        # optimiser iterates agree with the closed form only to ~1e-5
        assert_allclose(mod.get_blink_rate_mle(*six), rec['maxlikelihood']['numeric_mle'], rtol=1e-4)


@pytest.mark.parametrize('key', sorted(TOY))
def test_exact_posterior_agrees_with_reference_sampler(key):
    """Pins oracle/dense_oracle.py to the reference's own sampler output."""
    rec = TOY[key]
    mname, dname = key.split('/')
    M, D = MODELS[mname], DATA[dname]
    dp = DensePosterior(M, D)
    ex = dp.expected_summary()
    edges = dp.tb.edges
    exact = dict(
        off_root=ex['root_off_count'],
        # summary.py:221 with string class names counts every ON track (quirk 1)
        xon_root=ex['root_on_total'],
        pri_trans=sum(c for e in edges for a in ex['edge_to_pri_trans'][e]
            for c in ex['edge_to_pri_trans'][e][a].values()),
        off_xon_trans=sum(ex['edge_to_off_xon_trans'].values()),
        xon_off_trans=sum(ex['edge_to_xon_off_trans'].values()),
        off_xon_dwell=sum(ex['edge_to_off_xon_dwell'].values()),
        xon_off_dwell=sum(ex['edge_to_xon_off_dwell'].values()))
    for name, bm in rec['batch_means'].items():
        assert abs(bm['mean'] - exact[name]) <= 5 * bm['se'] + 1e-9, (name, bm, exact[name])


@pytest.mark.parametrize('key', ['BlinkModelA/DataB', 'BlinkModelB/DataC', 'BlinkModelC/DataA'])
def test_oracle_sampler_agrees_with_reference_sampler(key):
    rec = TOY[key]
    mname, dname = key.split('/')
    M, D = MODELS[mname], DATA[dname]
    tb = bo.build_tables(M)
    rng = np.random.default_rng(99)
    nchains, nsamp = 5, 300
    vals = []
    for _ in range(nchains):
        s = bo.Summary(tb.T, tb.root, tb.node_to_tm, tb.primary_to_tol, M.get_Q_primary())
        for pri, tols in bo.gen_samples(M, D, 30, nsamp, rng):
            s.on_sample(pri, tols)
        vals.append(gu.scalar_functionals(s))
    for name, bm in rec['batch_means'].items():
        x = np.array([v[name] for v in vals])
        se = np.hypot(bm['se'], x.std(ddof=1) / np.sqrt(nchains))
        assert abs(x.mean() - bm['mean']) <= 5 * se + 5e-3, (name, x.mean(), bm)
