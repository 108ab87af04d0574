"""M-step objective of the codon blinking model: the closed-form gradient against finite
differences and against the reference formulas of nxblink/em.py (through the host mirror).
CPU only: the statistics are synthetic, no sweep is run."""
import numpy as np
from numpy.testing import assert_allclose
import scipy.optimize

from nxblink_b200 import p53, p53em
from nxblink_b200.packing import PackedModel


class _FakeSummary(object):
    pass


def _mstep(seed=0):
    m = p53.p53_model()
    pm = PackedModel(m)
    rng = np.random.RandomState(seed)
    E, P, K, nnz = pm.n_nodes - 1, pm.n_states, pm.n_classes, pm.nnz
    s = _FakeSummary()
    s.nsamples = 500
    s.root_off_count = 2100
    s.root_xon_count = 7000
    s.flat = dict(
            pri_trans=rng.poisson(0.4, size=(E, nnz)),
            pri_dwell=rng.gamma(2.0, 3.0, size=(E, nnz)),
            off_on=rng.poisson(300, size=E), on_off=rng.poisson(300, size=E),
            off_xon_dwell=rng.gamma(5.0, 40.0, size=E), xon_off_dwell=rng.gamma(5.0, 90.0, size=E),
            root_pri=rng.multinomial(500, np.ones(P) / P))
    return p53em.MStep(pm, s, m.genetic_code), E


def test_gradient_matches_finite_differences():
    mstep, E = _mstep()
    rng = np.random.RandomState(1)
    for trial in range(3):
        x = np.concatenate([np.log([2.0, 0.5, 0.1, 0.2, 0.3, 0.4, 1.5, 0.5]) + 0.3 * rng.randn(8),
                np.log(0.15) + 0.5 * rng.randn(E)])
        f, g = mstep.neg_ll_and_grad(x)
        assert_allclose(f, mstep.neg_ll(x), rtol=1e-13)
        num = scipy.optimize.approx_fprime(x, mstep.neg_ll, 1e-6)
        assert_allclose(g, num, rtol=2e-5, atol=2e-5 * np.abs(num).max())


def test_gradient_is_invariant_to_nucleotide_scale():
    # the nucleotide weights are normalised inside the objective (p53em.py:160-166): the
    # gradient has no component along a common scaling of the four weights
    mstep, E = _mstep(3)
    x = np.concatenate([np.log([2.0, 0.5, 0.1, 0.2, 0.3, 0.4, 1.5, 0.5]), np.full(E, np.log(0.1))])
    f, g = mstep.neg_ll_and_grad(x)
    assert abs(g[2:6].sum()) < 1e-9 * np.abs(g[2:6]).max()


def test_edge_rate_stationary_point():
    mstep, E = _mstep(5)
    k, w, nt, on, off = 2.0, 0.5, np.array([0.1, 0.2, 0.3, 0.4]), 1.5, 0.5
    rates = mstep.closed_form_edge_rates(k, w, nt, on, off)
    x = np.concatenate([np.log([k, w]), np.log(nt), np.log([on, off]), np.log(rates)])
    f, g = mstep.neg_ll_and_grad(x)
    assert np.abs(g[8:]).max() < 1e-9 * mstep.trans_total.max()


def test_maximization_step_converges_with_the_analytic_gradient():
    mstep, E = _mstep(7)
    new = p53em.maximization_step(mstep, 2.0, 0.5, 0.25, 0.25, 0.25, 0.25, 1.0, 1.0, np.full(E, 0.1))
    x = np.concatenate([np.log([new['kappa'], new['omega'], new['A'], new['C'], new['G'], new['T'],
            new['rate_on'], new['rate_off']]), np.log(new['edge_rates'])])
    f, g = mstep.neg_ll_and_grad(x)
    assert new['result'].success
    assert np.abs(g).max() < 1e-3
