"""Deterministic parity of the dense path (FP64 tensor-pipe kernels behind
nxb_expm_batched / nxb_pruning_loglik): tolerance 1e-10 relative, as BASELINE.json
states for expm and pruning log-likelihoods.  Oracle: scipy.linalg.expm (the
reference's own numerical authority, toymodel-analytic.py:119) and the restated
pruning recursion of oracle/dense_oracle.py."""
import numpy as np
import pytest
import scipy.linalg
from numpy.testing import assert_allclose

from oracle.dense_oracle import DensePosterior
from nxblink_b200 import dense, packing, p53, toymodel, toydata

pytestmark = pytest.mark.gpu


def _rate_matrix(pm):
    P = pm.n_states
    Q = np.zeros((P, P))
    Q[pm.q_row, pm.q_col] = pm.q_val
    return Q - np.diag(Q.sum(axis=1))


def test_expm_batched_matches_scipy(nb):
    pm = packing.PackedModel(p53.p53_model())
    Q = _rate_matrix(pm)
    t = np.concatenate([pm.blen[1:] * pm.rate[1:], [0.0, 1e-9, 3.0, 25.0]])
    got = dense.expm_batched(Q, t)
    for b, tb in enumerate(t):
        exp = scipy.linalg.expm(Q * tb)
        assert_allclose(got[b], exp, rtol=1e-10, atol=1e-14)
        assert_allclose(got[b].sum(axis=1), 1.0, rtol=1e-12)
    cp = dense.CompoundProcess(toymodel.BlinkModelB)
    got = dense.expm_batched(cp.Q, [0.25, 0.5, 1.0])
    for b, tb in enumerate([0.25, 0.5, 1.0]):
        assert_allclose(got[b], scipy.linalg.expm(cp.Q * tb), rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize('D', [toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD],
        ids=lambda d: d.__name__)
@pytest.mark.parametrize('M', [toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC],
        ids=lambda m: m.__name__)
def test_compound_likelihood_matches_dense_oracle(nb, M, D):
    """toymodel-analytic.py:112-126: likelihood of the data under the compound process."""
    cp = dense.CompoundProcess(M)
    assert len(cp.states) == 24
    data = packing.pack_data(cp.pm, [D, D, D])
    ll = cp.loglik(data)
    exact = np.log(DensePosterior(M, D).likelihood)
    assert_allclose(ll, exact, rtol=1e-10, atol=1e-12)


def test_primary_pruning_loglik_p53_shape(nb):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    ls, tt, tf = p53.synthetic_alignment(pm, 200, seed=3)
    data = packing.pack_alignment(pm, ls, tt, tf)
    got = dense.primary_loglik(pm, data)
    # numpy / scipy restatement of the pruning recursion
    Q = _rate_matrix(pm)
    Pe = [scipy.linalg.expm(Q * (pm.blen[v] * pm.rate[v])) for v in range(1, pm.n_nodes)]
    P, N = pm.n_states, pm.n_nodes
    exp = np.zeros(200)
    for s in range(200):
        L = np.ones((N, P))
        for v in range(N):
            L[v] = [(int(data.pri_ok[s, v]) >> i) & 1 for i in range(P)]
        for v in range(N - 1, 0, -1):
            L[pm.parent[v]] *= Pe[v - 1].dot(L[v])
        exp[s] = np.log(pm.prior.dot(L[0]))
    assert_allclose(got, exp, rtol=1e-10)


def test_infeasible_site_is_minus_infinity(nb):
    cp = dense.CompoundProcess(toymodel.BlinkModelA)

    class Impossible(toydata.DataB):
        # state 4 observed at N3 but its class T2 declared off there
        TOL = dict(toydata.DataB.TOL, T2=dict(N3=(False,), N4=(True,)))
    data = packing.pack_data(cp.pm, [toydata.DataB, Impossible])
    ll = cp.loglik(data)
    assert np.isfinite(ll[0]) and ll[1] == -np.inf


def test_expm_frechet_batched_matches_scipy(nb):
    """nxb_expm_frechet_batched against scipy.linalg.expm_frechet, the numerical authority
    of nxblink/denseutil.py:148-176 (tolerance 1e-10 relative to the largest entry)."""
    rng = np.random.RandomState(5)
    for n, scale in ((5, 1.0), (24, 1.0), (48, 4.0), (64, 0.3)):
        Q = rng.gamma(1.0, 1.0, size=(n, n)) * (rng.uniform(size=(n, n)) < 0.3)
        np.fill_diagonal(Q, 0.0)
        Q -= np.diag(Q.sum(axis=1))
        t = np.array([0.0, 1e-6, 0.1, 1.0, scale * 3.0])
        E = rng.randn(len(t), n, n)
        P, L = dense.expm_frechet_batched(Q, t, E)
        for b, tb in enumerate(t):
            Pe, Le = scipy.linalg.expm_frechet(Q * tb, E[b])
            assert_allclose(P[b], Pe, rtol=1e-10, atol=1e-13)
            assert_allclose(L[b], Le, rtol=0, atol=1e-10 * np.abs(Le).max())
        L2 = dense.expm_frechet_batched(Q, t, E, compute_expm=False)
        assert (L2 == L).all()


@pytest.mark.parametrize('D', [toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD],
        ids=lambda d: d.__name__)
@pytest.mark.parametrize('M', [toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC],
        ids=lambda m: m.__name__)
def test_expected_summary_matches_dense_oracle(nb, M, D):
    """toymodel-analytic.py:128-226 on the device path (one Frechet derivative per edge)
    against the restated oracle (one scipy expm_frechet per edge and statistic)."""
    cp = dense.CompoundProcess(M)
    data = packing.pack_data(cp.pm, [D])
    got = cp.expected_summary(data)
    dp = DensePosterior(M, D)
    exact = dp.expected_summary()
    assert_allclose(got['likelihood'], dp.likelihood, rtol=1e-10)
    for s, v in exact['root_pri_to_count'].items():
        assert_allclose(got['root_pri_to_count'][s], v, rtol=1e-9, atol=1e-12)
    for key in ('root_on_total', 'root_off_count', 'root_xon_intended'):
        assert_allclose(got[key], exact[key], rtol=1e-9, atol=1e-12)
    for edge in cp.pm.edges:
        for key in ('edge_to_off_xon_trans', 'edge_to_xon_off_trans', 'edge_to_off_xon_dwell',
                'edge_to_xon_off_dwell'):
            assert_allclose(got[key][edge], exact[key][edge], rtol=1e-9, atol=1e-11, err_msg=key)
        for key in ('edge_to_pri_trans', 'edge_to_pri_dwell'):
            for sa in exact[key][edge]:
                for sb, v in exact[key][edge][sa].items():
                    assert_allclose(got[key][edge][sa][sb], v, rtol=1e-9, atol=1e-11, err_msg=key)
