"""The FP64 message path (msg_f64=1: chunk messages, tolerance (OFF, ON) messages, exp(-penalty),
gap logarithm and backward-sampling uniforms all in f64, as chunking.py:32-36,264-289 computes
them) and the event-parallel tolerance kernel (NXB_TOLPAR=1) get the same parity gates as the
default path: exact posterior at toy size, prior invariance at the p53 shape, and f32 against
f64 statistics within Monte Carlo error at the p53 shape.

Tolerances are Monte Carlo tolerances (5 standard errors from independent seeds, plus the
additive floor written next to each assert; the exact-posterior gates are those of
test_gpu_toy_exact.py: 32 seeds, 5 SE + 2e-4); nothing compared here is deterministic."""
import os

import numpy as np
import pytest

import golden_util as gu
from nxblink_b200 import toymodel, toydata, p53, packing
from test_gpu_toy_exact import exact_posterior_gate, MODELS, DATA

pytestmark = pytest.mark.gpu


class _env(object):
    """NXB_TOLPAR is read by nxb_create: set it around the construction of a sampler."""

    def __init__(self, tolpar):
        self.tolpar = tolpar

    def __enter__(self):
        self.old = os.environ.get('NXB_TOLPAR')
        os.environ['NXB_TOLPAR'] = '1' if self.tolpar else '0'

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop('NXB_TOLPAR', None)
        else:
            os.environ['NXB_TOLPAR'] = self.old


@pytest.mark.parametrize('D', DATA, ids=lambda d: d.__name__)
@pytest.mark.parametrize('M', MODELS, ids=lambda m: m.__name__)
def test_f64_messages_match_exact_posterior(nb, M, D):
    with _env(False):
        worst, rms = exact_posterior_gate(nb, M, D, msg_f64=True)
    assert worst[0] <= 1.0, worst
    assert rms < 1.5


@pytest.mark.parametrize('f64', [False, True], ids=['f32', 'f64'])
@pytest.mark.parametrize('M,D', [(toymodel.BlinkModelB, toydata.DataC), (toymodel.BlinkModelA, toydata.DataA),
    (toymodel.BlinkModelC, toydata.DataB), (toymodel.BlinkModelC, toydata.DataD)],
    ids=['B-C', 'A-A', 'C-B', 'C-D'])
def test_event_parallel_kernel_matches_exact_posterior(nb, M, D, f64):
    with _env(True):
        worst, rms = exact_posterior_gate(nb, M, D, msg_f64=f64)
    assert worst[0] <= 1.0, worst
    assert rms < 1.5


def _edge_matrix(sm):
    """[E, 6] per-sample per-edge statistics in the order of make_golden.EDGE_STATS."""
    n = float(sm.nsamples)
    f = sm.flat
    return np.stack([f['pri_trans'].sum(axis=1), f['pri_dwell'].sum(axis=1), f['off_on'], f['on_off'],
        f['off_xon_dwell'], f['xon_off_dwell']], axis=1).astype(np.float64) / n


@pytest.mark.parametrize('tolpar', [False, True], ids=['walk', 'event-parallel'])
def test_prior_invariance_p53_shape_f64(nb, tolpar):
    """No data: the sweep must leave the prior invariant (forward-simulated trajectories before
    and after 6 sweeps), with f64 messages; validate=1 checks every invariant after every phase."""
    pm = packing.PackedModel(p53.p53_model())
    fwd, mcmc = [], []
    for seed in range(16):          # (16 seeds x 4096 units: enough degrees of freedom for a 5-SE rule)
        with _env(tolpar):
            s = nb.BlinkSampler(pm, validate=(seed < 4), msg_f64=True)
        s.forward_sample(4096, seed=200 + seed)
        s.reset_summary()
        s.summarize_current()
        fwd.append(_edge_matrix(s.summary()))
        s.reset_summary()
        s.sweep(6, accumulate=True)
        sm = s.summary()
        assert sm.nsamples == 4096 * 6
        mcmc.append(_edge_matrix(sm))
        s.close()
    a, b = np.array(fwd), np.array(mcmc)
    se = np.hypot(a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)) / np.sqrt(len(fwd))
    # per edge and statistic (288 comparisons); floor 1e-3 relative to the edge's scale
    lim = 5 * se + 1e-3 * np.maximum(1.0, np.abs(a.mean(axis=0)))
    bad = np.abs(a.mean(axis=0) - b.mean(axis=0)) > lim
    assert not bad.any(), (np.argwhere(bad), a.mean(axis=0)[bad], b.mean(axis=0)[bad], se[bad])


def test_f32_and_f64_messages_agree_p53_shape(nb):
    """Same data, same seeds: the f32-message sweep and the f64-message sweep sample the same
    posterior -- per-edge statistics within 5 combined standard errors (8 sites x 64 chains x
    20 sweeps per seed, 16 seeds each)."""
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    ls, tt, tf = p53.synthetic_alignment(pm, 8, seed=3, reference_leaf=pm.node_index[m.name_to_leaf['Has']])
    data = packing.pack_alignment(pm, ls, tt, tf)
    res = {}
    for f64 in (False, True):
        mats, launched = [], None
        for seed in range(16):
            with _env(False):
                s = nb.BlinkSampler(pm, validate=(seed < 2), msg_f64=f64)
            s.set_data(data)
            s.init_chains(chains=64, seed=40 + seed)
            s.sweep(30, accumulate=False)
            s.sweep(20, accumulate=True)
            mats.append(_edge_matrix(s.summary()))
            s.close()
        res[f64] = np.array(mats)
    a, b = res[False], res[True]
    se = np.hypot(a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)) / np.sqrt(a.shape[0])
    lim = 5 * se + 1e-3 * np.maximum(1.0, np.abs(a.mean(axis=0)))
    bad = np.abs(a.mean(axis=0) - b.mean(axis=0)) > lim
    assert not bad.any(), (np.argwhere(bad), a.mean(axis=0)[bad], b.mean(axis=0)[bad], se[bad])


def test_large_penalty_exponents_do_not_make_f32_infeasible(nb):
    """Toy model B with every edge rate x100 and slow blinking (on 0.02, off 0.005): few blink
    events per edge, so the penalty exponent rate * Q_meta * dt between two live events reaches
    ~100, beyond the f32 range of exp().  A penalty factor must never turn into a hard zero
    (chunking.py:32-36 computes exp(-x) in f64, which is tiny but positive): the f32 path clamps it
    to the smallest normal number, so data that exclude OFF stay feasible.  f32 and f64 statistics
    agree within Monte Carlo error."""
    class Fast(toymodel.BlinkModelB):
        @classmethod
        def get_edge_to_rate(cls):
            return dict((e, 100.0 * r) for e, r in toymodel.BlinkModelB.get_edge_to_rate().items())

        @classmethod
        def get_rate_on(cls):
            return 0.02

        @classmethod
        def get_rate_off(cls):
            return 0.005
    res = {}
    for f64 in (False, True):
        mats = []
        for seed in range(16):                      # (enough seeds for a 5-SE rule: t_15 tails)
            with _env(False):
                s = nb.BlinkSampler(Fast, validate=(seed < 2), msg_f64=f64)
            s.set_data([toydata.DataC, toydata.DataD])
            s.init_chains(chains=256, seed=70 + seed)
            s.sweep(30, accumulate=False)           # raises NxbError on an infeasible unit
            s.sweep(30, accumulate=True)
            mats.append(_edge_matrix(s.summary()))
            s.close()
        res[f64] = np.array(mats)
    a, b = res[False], res[True]
    se = np.hypot(a.std(axis=0, ddof=1), b.std(axis=0, ddof=1)) / np.sqrt(a.shape[0])
    lim = 5 * se + 2e-3 * np.maximum(1.0, np.abs(a.mean(axis=0)))
    assert (np.abs(a.mean(axis=0) - b.mean(axis=0)) <= lim).all()
