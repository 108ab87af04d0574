"""The oracle sampler against the exact compound-process posterior
(oracle/dense_oracle.py restating compound.py / denseutil.py /
toymodel-analytic.py), and internal consistency of the dense oracle."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import blink_oracle as bo
from oracle.dense_oracle import DensePosterior
from nxblink_b200 import toymodel, toydata


def test_dense_prior_is_stationary_and_sizes():
    dp = DensePosterior(toymodel.BlinkModelA, toydata.DataA)
    assert len(dp.comp) == 6 * 4           # P * 2**(K-1) feasible compound states
    assert_allclose(dp.prior.dot(dp.Q), 0, atol=1e-12)
    assert_allclose(dp.likelihood, 1.0, rtol=1e-12)     # no data
    ex = dp.expected_summary()
    # no data: the root posterior is the prior; dwell on an edge sums to (K-1) * blen
    for s, p in toymodel.BlinkModelA.get_primary_distn().items():
        assert_allclose(ex['root_pri_to_count'][s], p, rtol=1e-10)
    for e in dp.tb.edges:
        assert_allclose(ex['edge_to_off_xon_dwell'][e] + ex['edge_to_xon_off_dwell'][e],
                2.0 * dp.tb.edge_to_blen[e], rtol=1e-9)


def test_dense_stationary_flux():
    # no data + stationarity: expected blink transitions per edge = rate * blen * flux
    M = toymodel.BlinkModelB
    dp = DensePosterior(M, toydata.DataA)
    ex = dp.expected_summary()
    for e in dp.tb.edges:
        rate = dp.tb.edge_to_rate[e]
        off_dwell = ex['edge_to_off_xon_dwell'][e]
        xon_dwell = ex['edge_to_xon_off_dwell'][e]
        assert_allclose(ex['edge_to_off_xon_trans'][e], rate * M.get_rate_on() * off_dwell, rtol=1e-8, atol=1e-12)
        assert_allclose(ex['edge_to_xon_off_trans'][e], rate * M.get_rate_off() * xon_dwell, rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize('M,D', [
    (toymodel.BlinkModelA, toydata.DataB),
    (toymodel.BlinkModelB, toydata.DataC),
    (toymodel.BlinkModelC, toydata.DataD),
])
def test_oracle_sampler_matches_exact_posterior(M, D):
    ex = DensePosterior(M, D).expected_summary()
    rng = np.random.default_rng(2024)
    tb = bo.build_tables(M)
    nchains, nburn, nsamp = 6, 30, 250
    per_chain = []
    for _ in range(nchains):
        s = bo.Summary(tb.T, tb.root, tb.node_to_tm, tb.primary_to_tol, M.get_Q_primary())
        for pri, tols in bo.gen_samples(M, D, nburn, nsamp, rng):
            s.on_sample(pri, tols)
        per_chain.append(s)

    def check(get, exact, floor):
        vals = np.array([get(s) / s.nsamples for s in per_chain])
        se = max(vals.std(ddof=1) / np.sqrt(nchains), floor)
        assert abs(vals.mean() - exact) < 5 * se, (vals.mean(), exact, se)
    for e in tb.edges:
        check(lambda s: s.edge_to_off_xon_trans[e], ex['edge_to_off_xon_trans'][e], 0.01)
        check(lambda s: s.edge_to_xon_off_trans[e], ex['edge_to_xon_off_trans'][e], 0.01)
        check(lambda s: s.edge_to_off_xon_dwell[e], ex['edge_to_off_xon_dwell'][e], 0.01)
        check(lambda s: sum(c for a in s.edge_to_pri_trans[e] for c in s.edge_to_pri_trans[e][a].values()),
                sum(c for a in ex['edge_to_pri_trans'][e] for c in ex['edge_to_pri_trans'][e][a].values()), 0.015)
    check(lambda s: s.root_off_count, ex['root_off_count'], 0.01)
    check(lambda s: s.root_xon_intended, ex['root_xon_intended'], 0.01)
