"""Sharp parity gate at the p53 shape: the CUDA sweep against a LONG run of the reference's own
sampler (tests/golden/ref_p53_long_site*.json: 10 000 samples per site after 50 burn-in sweeps,
per-edge statistics with batch-means standard errors over 50 batches; generator
oracle/refcompat/make_golden.py::make_p53_long, which imports the reference's raoteh.gen_samples
and summary.Summary).

Per site 48 edges x 6 statistics + 2 root statistics = 290 comparisons.  Criterion, per
statistic: |gpu - ref| <= 5 * sqrt(se_ref^2 + se_gpu^2) + 5e-4 (no slack proportional to the
value).  Cells in which the reference run saw NO event at all (mean = se = 0: primary transitions
on short edges) carry no error estimate: the reference is ONE chain of 10 000 consecutive sweeps in
which rare events arrive in clusters (a primary event on such an edge survives many sweeps), so
"none seen" only bounds the rate at the 1e-3 level; those cells must stay below 5e-3 per sample."""
import os

import numpy as np
import pytest

import golden_util as gu
from nxblink_b200 import p53, packing
from test_gpu_precision import _edge_matrix

pytestmark = pytest.mark.gpu

FILES = [f for f in ('ref_p53_long_site0.json', 'ref_p53_long_site1.json', 'ref_p53_long_site2.json')
        if os.path.exists(os.path.join(gu.GOLDEN, f))]


def test_fixtures_present():
    assert len(FILES) == 3


@pytest.mark.parametrize('f64', [False, True], ids=['f32', 'f64'])
@pytest.mark.parametrize('fname', FILES)
def test_per_edge_statistics_against_long_reference_run(nb, fname, f64):
    rec = gu.load(fname)
    assert rec['nsamples'] >= 10000 and rec['stats'] == ['pri_trans', 'pri_dwell', 'off_xon_trans',
            'xon_off_trans', 'off_xon_dwell', 'xon_off_dwell']
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    assert [repr(n) for n in pm.nodes] == rec['node_order']
    data = packing.pack_alignment(pm, np.array([rec['leaf_states']]),
            np.array([rec['tol_true_ok']], dtype=np.uint32), np.array([rec['tol_false_ok']], dtype=np.uint32))
    mats, roots = [], []
    nseeds = 16
    for seed in range(nseeds):
        s = nb.BlinkSampler(pm, validate=(seed < 2), msg_f64=f64)
        s.set_data(data)
        s.init_chains(chains=1024, seed=900 + seed)
        s.sweep(50, accumulate=False)
        s.sweep(25, accumulate=True)
        sm = s.summary()
        assert sm.nsamples == 1024 * 25
        mats.append(_edge_matrix(sm))
        roots.append((sm.root_off_count / float(sm.nsamples), sm.root_xon_count / float(sm.nsamples)))
        s.close()
    g = np.array(mats)
    ref, ref_se = np.array(rec['edge_mean']), np.array(rec['edge_se'])
    se = np.hypot(ref_se, g.std(axis=0, ddof=1) / np.sqrt(nseeds))
    diff = np.abs(g.mean(axis=0) - ref)
    unseen = (ref == 0.0) & (ref_se == 0.0)
    bad = np.where(unseen, g.mean(axis=0) > 5e-3, diff > 5 * se + 5e-4)
    assert unseen.sum() < 0.1 * unseen.size
    assert not bad.any(), (np.argwhere(bad), g.mean(axis=0)[bad], ref[bad], se[bad])
    # the comparison has power: combined SE is a few percent of the typical blink statistic
    assert np.median(se[:, 2:] / np.maximum(ref[:, 2:], 1e-12)) < 0.1
    r = np.array(roots)
    for i, name in enumerate(('off_root', 'xon_root')):
        se_r = np.hypot(rec['root_se'][name], r[:, i].std(ddof=1) / np.sqrt(nseeds))
        assert abs(r[:, i].mean() - rec['root_mean'][name]) <= 5 * se_r + 5e-4, (name, r[:, i].mean(), rec['root_mean'][name], se_r)
