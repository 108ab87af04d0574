"""Parity gate for the sampled quantities: the CUDA sweep against the EXACT
posterior expectations of the 48-state compound process (oracle/dense_oracle.py,
restating compound.py / denseutil.py / toymodel-analytic.py), for the reference's
full smoke grid (test_smoke.py:19-30: 3 models x 4 data levels).

Every statistic that nxblink.summary.Summary accumulates is tested individually
with a z-test whose standard error comes from 32 independent seeds.  Tolerance:
|mean - exact| <= 5 * SE + 2e-4 (Monte Carlo error only; the additive term is there
for statistics whose standard error is ~0, not as slack: typical SE is 1e-4..1e-3;
there is no FP tolerance to state because nothing here is deterministic).  The
chains run 100 burn-in sweeps from the reference-style initial trajectories
(raoteh.py:337-361); the grids WITHOUT data mix slowly from there (a 1e-3 residue
after 100 sweeps), so they start from exact prior draws instead (device forward
sampler, fwdsample.py:67-262) and the sweep has to keep them on the prior."""
import numpy as np
import pytest

from oracle.dense_oracle import DensePosterior
from nxblink_b200 import toymodel, toydata

pytestmark = pytest.mark.gpu

MODELS = [toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC]
DATA = [toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD]


def _flatten(sm, pm):
    n = float(sm.nsamples)
    out = {}
    for s in pm.states:
        out['root_pri', s] = sm.root_pri_to_count.get(s, 0) / n
    out['root_off'] = sm.root_off_count / n
    out['root_xon'] = sm.root_xon_intended / n
    for e in pm.edges:
        out['off_xon_trans', e] = sm.edge_to_off_xon_trans[e] / n
        out['xon_off_trans', e] = sm.edge_to_xon_off_trans[e] / n
        out['off_xon_dwell', e] = sm.edge_to_off_xon_dwell[e] / n
        out['xon_off_dwell', e] = sm.edge_to_xon_off_dwell[e] / n
        for i in range(pm.nnz):
            sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
            out['pri_trans', e, sa, sb] = sm.flat['pri_trans'][pm.edge_index[e], i] / n
            out['pri_dwell', e, sa, sb] = sm.flat['pri_dwell'][pm.edge_index[e], i] / n
    return out


def _flatten_exact(ex, pm):
    out = {}
    for s in pm.states:
        out['root_pri', s] = ex['root_pri_to_count'][s]
    out['root_off'] = ex['root_off_count']
    out['root_xon'] = ex['root_xon_intended']
    for e in pm.edges:
        for key in ('off_xon_trans', 'xon_off_trans', 'off_xon_dwell', 'xon_off_dwell'):
            out[key, e] = ex['edge_to_' + key][e]
        for i in range(pm.nnz):
            sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
            out['pri_trans', e, sa, sb] = ex['edge_to_pri_trans'][e][sa][sb]
            out['pri_dwell', e, sa, sb] = ex['edge_to_pri_dwell'][e][sa][sb]
    return out


NSEEDS = 32


def exact_posterior_gate(nb, M, D, nseeds=NSEEDS, **kw):
    """Runs `nseeds` independent batches of 1024 chains x 40 accumulated sweeps and returns
    (worst, rms z): worst = max over the statistics of |mean - exact| / (5 SE + 2e-4)."""
    exact = None
    runs = []
    for seed in range(nseeds):
        s = nb.BlinkSampler(M, validate=(seed < 4), **kw)  # (every invariant after every phase, 4 of the runs)
        if D is toydata.DataA:
            s.forward_sample(1024, seed=1000 + seed)       # exact draws from the prior = the posterior without data
            s.reset_summary()
            s.sweep(20, accumulate=False)
        else:
            s.set_data([D])
            s.init_chains(chains=1024, seed=1000 + seed)
            s.sweep(100, accumulate=False)
        s.sweep(40, accumulate=True)
        sm = s.summary()
        assert sm.nsamples == 1024 * 40
        if exact is None:
            exact = _flatten_exact(DensePosterior(M, D).expected_summary(), s.pm)
        runs.append(_flatten(sm, s.pm))
        s.close()
    worst = (0.0, None)
    zs = []
    for key, ex in exact.items():
        vals = np.array([r[key] for r in runs])
        se = vals.std(ddof=1) / np.sqrt(nseeds)
        z = abs(vals.mean() - ex) / (5 * se + 2e-4)
        if se > 0:
            zs.append(abs(vals.mean() - ex) / se)
        if z > worst[0]:
            worst = (z, (key, vals.mean(), ex, se))
    return worst, float(np.sqrt(np.mean(np.square(zs))))


@pytest.mark.parametrize('D', DATA, ids=lambda d: d.__name__)
@pytest.mark.parametrize('M', MODELS, ids=lambda m: m.__name__)
def test_summary_statistics_match_exact_posterior(nb, M, D):
    worst, rms = exact_posterior_gate(nb, M, D)
    assert worst[0] <= 1.0, worst
    # unbiased estimates: the z-scores of the ~150 statistics have unit scale
    assert rms < 1.5


def test_reference_compatible_root_xon_count(nb):
    """summary.py:221 compares a bool with the class NAME: with the
    toy models' string names every ON track is counted, so ref = intended + nsamples."""
    s = nb.BlinkSampler(toymodel.BlinkModelA)
    s.set_data([toydata.DataA])
    s.init_chains(chains=256, seed=3)
    s.sweep(8, accumulate=True)
    sm = s.summary()
    assert sm.root_xon_count + sm.root_off_count == 3 * sm.nsamples
    assert sm.root_xon_count == sm.root_xon_intended + sm.nsamples
    s.close()
