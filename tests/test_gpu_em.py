"""The EM / ML loop on p53-shaped data: the vectorised M-step
objective equals the reference-mirror em.get_ll_* (<= 1e-10), and one EM iteration runs
end to end on the device and improves its own objective."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from nxblink_b200 import p53, p53em, packing, em

pytestmark = pytest.mark.gpu


def test_em_iteration_p53_shape(nb):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 64, seed=1, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    prm = p53.jeff_params()
    params = dict(kappa=prm['kappa'], omega=prm['omega'], A=prm['A'], C=prm['C'], G=prm['G'],
            T=prm['T'], rate_on=1.5, rate_off=0.5, edge_rates=np.maximum(pm.rate[1:], 1e-6))
    new = p53em.em_iteration(m, data, params, chains=8, nburnin=8, nsamples=8, seed=3)
    mstep, summary = new['mstep'], new['summary']
    assert summary.nsamples == 64 * 8 * 8
    # (1) vectorised objective == the em.py mirror on the same summary
    x0 = np.log(np.concatenate([[params[k] for k in ('kappa', 'omega', 'A', 'C', 'G', 'T',
        'rate_on', 'rate_off')], params['edge_rates']]))
    kappa, omega, nt, on, off, rates, _ = p53em.unpack_params(x0)
    P = pm.n_states
    pre_Q = np.zeros((P, P))
    pre_Q[pm.q_row, pm.q_col] = mstep.pre_Q(kappa, omega, nt)
    distn = mstep.distn(nt)
    ref = -(em.get_ll_root(summary, distn, on, off)
            + em.get_ll_dwell(summary, pre_Q, distn, on, off, pm.edges, rates)
            + em.get_ll_trans(summary, pre_Q, distn, on, off, pm.edges, rates))
    assert_allclose(mstep.neg_ll(x0), ref, rtol=1e-10)
    # (2) the M-step does not increase the objective and lands on the closed-form
    #     stationary point of the edge rates
    assert new['neg_ll'] <= mstep.neg_ll(x0) + 1e-12
    cf = mstep.closed_form_edge_rates(new['kappa'], new['omega'],
            np.array([new['A'], new['C'], new['G'], new['T']]), new['rate_on'], new['rate_off'])
    big = (mstep.trans_total > 50) & (cf > 0.01)
    assert_allclose(new['edge_rates'][big], cf[big], rtol=5e-2)
    x_cf = np.array(new['result'].x)
    x_cf[8:] = np.log(np.maximum(cf, 1e-12))
    assert abs(mstep.neg_ll(x_cf) - new['neg_ll']) < 1e-4 * abs(new['neg_ll'])
    assert 0.2 < new['rate_on'] < 10 and 0.05 < new['rate_off'] < 5
