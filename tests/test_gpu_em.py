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
    new = p53em.em_iteration(m, data, params, chains=8, nburnin=8, nsamples=8, seed=3, mstep='host')
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


def test_update_model_equals_a_fresh_handle(nb):
    """nxb_update_model swaps the model values in place: after re-initialising the chains the
    sweep gives exactly what a handle created for the new model gives (same seed), and the kept
    chains can also be continued under the new rates (warm start) with every invariant holding."""
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 16, seed=2, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    prm = p53.jeff_params()
    edge_to_rate = dict((e, 0.7 * r + 0.01) for e, r in m.get_edge_to_rate().items()
            if True)
    edge_to_rate = dict((e, (r if m.get_edge_to_rate()[e] > 0 else 0.0)) for e, r in edge_to_rate.items())
    m2 = p53.Model(1.3 * prm['kappa'], 0.8 * prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'],
            1.1, 0.7, m._tree, m._root, m.get_edge_to_blen(), edge_to_rate, m.genetic_code)
    pm2 = packing.PackedModel(m2)

    def run(s):
        s.set_data(data)
        s.init_chains(chains=8, seed=11)
        s.sweep(3, accumulate=False)
        s.reset_summary()
        s.sweep(4, accumulate=True)
        return s.read_summary()

    fresh = nb.BlinkSampler(pm2, validate=True)
    c_fresh, d_fresh = run(fresh)
    fresh.close()
    s = nb.BlinkSampler(pm, validate=True)
    run(s)
    s.update_model(pm2)
    c_upd, d_upd = run(s)
    assert (c_fresh == c_upd).all()
    assert_allclose(d_fresh, d_upd, rtol=1e-11, atol=1e-9)
    # warm start: continue the chains under the first model's rates again
    s.update_model(pm)
    s.sweep(3, accumulate=True)                 # validate=True checks every invariant
    assert s.read_summary()[0][s.layout.c_nsamples] == 16 * 8 * 7
    # a model of another shape is refused
    with pytest.raises(Exception):
        s.update_model(packing.PackedModel(__import__('nxblink_b200').toymodel.BlinkModelA))
    s.close()


def test_em_iterations_reuse_the_sampler(nb):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 32, seed=4, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    prm = p53.jeff_params()
    params = dict(kappa=prm['kappa'], omega=prm['omega'], A=prm['A'], C=prm['C'], G=prm['G'],
            T=prm['T'], rate_on=1.5, rate_off=0.5, edge_rates=np.maximum(pm.rate[1:], 1e-6))
    fresh = p53em.em_iteration(m, data, params, chains=4, nburnin=4, nsamples=4, seed=9)
    sampler = nb.BlinkSampler(pm)
    first = p53em.em_iteration(m, data, params, chains=4, nburnin=4, nsamples=4, seed=9, sampler=sampler)
    assert first['sampler'] is sampler
    # (dwell sums are floating-point atomics: equal to ~1e-12, the optimiser amplifies that)
    assert_allclose(first['neg_ll'], fresh['neg_ll'], rtol=1e-5)
    keys = ('kappa', 'omega', 'A', 'C', 'G', 'T', 'rate_on', 'rate_off', 'edge_rates')
    second = p53em.em_iteration(m, data, dict((k, first[k]) for k in keys), chains=4, nburnin=4,
            nsamples=4, seed=10, sampler=sampler)
    assert second['summary'].nsamples == 32 * 4 * 4 and np.isfinite(second['neg_ll'])
    sampler.close()


def test_device_mstep_matches_the_host_objective_and_optimum(nb):
    """nxb_mstep (csrc/mstep_kernel.cu) on the device-resident statistics against the numpy
    objective / closed-form gradient (p53em.MStep) and the em.py mirror: values and gradients to
    1e-10, and the on-device L-BFGS lands on the optimum scipy's L-BFGS-B finds."""
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 64, seed=1, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    prm = p53.jeff_params()
    s = nb.BlinkSampler(pm)
    s.set_data(data)
    s.init_chains(chains=8, seed=3)
    s.sweep(8, accumulate=False)
    s.sweep(8, accumulate=True)
    summary = s.summary()
    host = p53em.MStep(pm, summary, m.genetic_code)
    dev = p53em.DeviceMStep(s, m.genetic_code)
    x0 = np.log(np.concatenate([[prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'], 1.5, 0.5],
        np.maximum(pm.rate[1:], 1e-6)]))
    rng = np.random.default_rng(0)
    for x in (x0, x0 + 0.3 * rng.standard_normal(len(x0))):
        fh, gh = host.neg_ll_and_grad(x)
        fd, gd = dev.neg_ll_and_grad(x)
        assert_allclose(fd, fh, rtol=1e-10)
        assert_allclose(gd, gh, rtol=1e-8, atol=1e-11 * max(1.0, np.abs(gh).max()))
    # the em.py mirror (nxblink/em.py:19-198) at x0
    kappa, omega, nt, on, off, rates, _ = p53em.unpack_params(x0)
    P = pm.n_states
    pre_Q = np.zeros((P, P))
    pre_Q[pm.q_row, pm.q_col] = host.pre_Q(kappa, omega, nt)
    distn = host.distn(nt)
    ref = -(em.get_ll_root(summary, distn, on, off)
            + em.get_ll_dwell(summary, pre_Q, distn, on, off, pm.edges, rates)
            + em.get_ll_trans(summary, pre_Q, distn, on, off, pm.edges, rates))     # (per sample)
    assert_allclose(dev.neg_ll_and_grad(x0)[0], ref, rtol=1e-10)
    # optimum
    sci = p53em.maximization_step(host, prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'],
            1.5, 0.5, np.maximum(pm.rate[1:], 1e-6))
    x, f, g, info = dev.minimize(x0, maxiter=200)
    assert info['status'] in (1, 2), info
    assert_allclose(host.neg_ll(x), f, rtol=1e-10)              # the device's own value is right
    assert abs(f - sci['neg_ll']) <= 2e-6 * abs(sci['neg_ll']), (f, sci['neg_ll'], info)
    # (both optimisers stop on scipy's relative-decrease rule, 2.2e-9: the objective is flat in the
    # blink rates, so the remaining gradient is small but not tiny)
    assert np.abs(g).max() < 5e-2 and f <= sci['neg_ll'] + 1e-7 * abs(sci['neg_ll'])
    new = p53em.m_step(s, None, m.genetic_code, prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'],
            prm['T'], 1.5, 0.5, np.maximum(pm.rate[1:], 1e-6))
    assert new['where'].startswith('device')
    for k in ('kappa', 'omega', 'rate_on', 'rate_off'):
        assert_allclose(new[k], sci[k], rtol=2e-2)
    # the edge rates sit exactly on their closed-form stationary point at the device's optimum ...
    cf = host.closed_form_edge_rates(new['kappa'], new['omega'],
            np.array([new['A'], new['C'], new['G'], new['T']]), new['rate_on'], new['rate_off'])
    has = host.trans_total > 0
    assert_allclose(new['edge_rates'][has], cf[has], rtol=1e-9)
    # ... and agree with where scipy's 200 L-BFGS-B iterations (not converged in the small rates) ended
    big = (host.trans_total > 50) & (cf > 0.01)
    assert_allclose(new['edge_rates'][big], sci['edge_rates'][big], rtol=5e-2)
    s.close()


def test_em_iteration_on_device_reduced_statistics(nb):
    """The multi-GPU form of the EM iteration: the statistics are all-reduced on the device
    (copies; a no-op sum in a single process) and the device M-step runs on the reduced buffers --
    same result as on the sampler's own accumulators."""
    torch = pytest.importorskip('torch')
    from nxblink_b200 import distributed as nd
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 32, seed=4, reference_leaf=leaf)
    data = packing.pack_alignment(pm, ls, tt, tf)
    prm = p53.jeff_params()
    params = dict(kappa=prm['kappa'], omega=prm['omega'], A=prm['A'], C=prm['C'], G=prm['G'],
            T=prm['T'], rate_on=1.5, rate_off=0.5, edge_rates=np.maximum(pm.rate[1:], 1e-6))
    own = p53em.em_iteration(m, data, params, chains=4, nburnin=4, nsamples=4, seed=9)
    red = p53em.em_iteration(m, data, params, chains=4, nburnin=4, nsamples=4, seed=9,
            reduce_device_fn=nd.allreduce_device)
    assert own['where'].startswith('device') and red['where'].startswith('device')
    # (dwell sums are floating-point atomics: equal to ~1e-12, the optimiser amplifies that)
    assert_allclose(red['neg_ll'], own['neg_ll'], rtol=1e-6)
    for k in ('kappa', 'omega', 'rate_on', 'rate_off'):
        assert_allclose(red[k], own[k], rtol=1e-4)
