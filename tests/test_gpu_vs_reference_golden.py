"""The CUDA sweep against statistics recorded from the REFERENCE's own sampler
(tests/golden/ref_toy.json, ref_p53_site*.json; generator
oracle/refcompat/make_golden.py).  Monte Carlo tolerance: 5 combined standard
errors (reference batch means; the GPU side uses far more samples)."""
import os

import numpy as np
import pytest

import golden_util as gu
from nxblink_b200 import toymodel, toydata, p53, packing

pytestmark = pytest.mark.gpu

TOY = gu.load('ref_toy.json')
MODELS = dict((m.__name__, m) for m in (toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC))
DATA = dict((d.__name__, d) for d in (toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD))


def _gpu_functionals(nb, model, data, nseeds=4, chains=2048, burn=40, sweeps=30, **kw):
    vals = []
    for seed in range(nseeds):
        s = nb.BlinkSampler(model, **kw)
        s.set_data(data)
        s.init_chains(chains=chains, seed=500 + seed)
        s.sweep(burn, accumulate=False)
        s.sweep(sweeps, accumulate=True)
        vals.append(gu.scalar_functionals(s.summary()))
        s.close()
    return vals


@pytest.mark.parametrize('key', sorted(TOY))
def test_toy_grid_against_reference_sampler(nb, key):
    rec = TOY[key]
    mname, dname = key.split('/')
    vals = _gpu_functionals(nb, MODELS[mname], [DATA[dname]])
    for name, bm in rec['batch_means'].items():
        x = np.array([v[name] for v in vals])
        se = np.hypot(bm['se'], x.std(ddof=1) / np.sqrt(len(x)))
        assert abs(x.mean() - bm['mean']) <= 5 * se + 1e-3, (name, x.mean(), bm)


P53_FILES = [f for f in ('ref_p53_site0.json', 'ref_p53_site1.json', 'ref_p53_site2.json')
        if os.path.exists(os.path.join(gu.GOLDEN, f))]


@pytest.mark.parametrize('fname', P53_FILES)
def test_p53_shape_against_reference_sampler(nb, fname):
    rec = gu.load(fname)
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    assert [repr(n) for n in pm.nodes] == rec['node_order']
    ls = np.array([rec['leaf_states']])
    tt = np.array([rec['tol_true_ok']], dtype=np.uint32)
    tf = np.array([rec['tol_false_ok']], dtype=np.uint32)
    data = packing.pack_alignment(pm, ls, tt, tf)
    vals = _gpu_functionals(nb, pm, data, nseeds=4, chains=512, burn=30, sweeps=20, validate=True)
    for name, bm in rec['batch_means'].items():
        x = np.array([v[name] for v in vals])
        se = np.hypot(bm['se'], x.std(ddof=1) / np.sqrt(len(x)))
        assert abs(x.mean() - bm['mean']) <= 5 * se + 1e-3, (name, x.mean(), bm)
