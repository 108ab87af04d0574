"""Deterministic parity: on FIXED trajectories (exported from the device) the
device statistics, the host-side Summary and the oracle's Summary
(summary.py:190-243 restated) must agree -- integers exactly, dwell sums to
1e-10 relative -- and so must the EM objective pieces (em.py) built on them."""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from oracle import blink_oracle as bo
from nxblink_b200 import toymodel, toydata, em, maxlikelihood, p53, packing
from nxblink_b200.summary import Summary

pytestmark = pytest.mark.gpu


def _oracle_summary(pm, model, tracks_list):
    s = bo.Summary(pm.T, pm.root, pm.node_to_tm, pm.primary_to_tol, model.get_Q_primary())
    for pri, tols in tracks_list:
        s.on_sample(pri, tols)
    return s


def _compare(dev, orc, pm):
    assert dev.nsamples == orc.nsamples
    assert dict(dev.root_pri_to_count) == dict((k, v) for k, v in orc.root_pri_to_count.items() if v)
    assert dev.root_off_count == orc.root_off_count
    assert dev.root_xon_count == orc.root_xon_count
    assert dev.root_on_total == orc.root_on_total
    assert dev.root_xon_intended == orc.root_xon_intended
    for e in pm.edges:
        assert dev.edge_to_off_xon_trans[e] == orc.edge_to_off_xon_trans[e]
        assert dev.edge_to_xon_off_trans[e] == orc.edge_to_xon_off_trans[e]
        assert_allclose(dev.edge_to_off_xon_dwell[e], orc.edge_to_off_xon_dwell[e], rtol=1e-10, atol=1e-11)
        assert_allclose(dev.edge_to_xon_off_dwell[e], orc.edge_to_xon_off_dwell[e], rtol=1e-10, atol=1e-11)
        got = dict(((a, b), dev.edge_to_pri_trans[e][a][b]['weight']) for a, b in dev.edge_to_pri_trans[e].edges())
        exp = dict(((a, b), c) for a in orc.edge_to_pri_trans[e] for b, c in orc.edge_to_pri_trans[e][a].items())
        assert got == exp
        gd = dict(((a, b), dev.edge_to_pri_dwell[e][a][b]['weight']) for a, b in dev.edge_to_pri_dwell[e].edges())
        ed = dict(((a, b), c) for a in orc.edge_to_pri_dwell[e] for b, c in orc.edge_to_pri_dwell[e][a].items() if c > 0)
        assert set(gd) == set(ed)
        for k in ed:
            assert_allclose(gd[k], ed[k], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize('M,D', [(toymodel.BlinkModelA, toydata.DataA),
    (toymodel.BlinkModelB, toydata.DataC), (toymodel.BlinkModelC, toydata.DataD)])
def test_device_statistics_equal_oracle_on_fixed_trajectories(nb, M, D):
    s = nb.BlinkSampler(M, validate=True)
    s.set_data([D])
    chains = 24
    s.init_chains(chains=chains, seed=42)
    s.sweep(5, accumulate=False)
    # (1) fused-in-sweep statistics of ONE sweep == statistics of the state it left
    s.reset_summary()
    s.sweep(1, accumulate=True)
    fused = s.summary(Q_primary=M.get_Q_primary())
    tracks = [s.export_tracks(u) for u in range(chains)]
    orc = _oracle_summary(s.pm, M, tracks)
    _compare(fused, orc, s.pm)
    # (2) the independent (non-fused) kernel
    s.reset_summary()
    s.summarize_current()
    _compare(s.summary(Q_primary=M.get_Q_primary()), orc, s.pm)
    # (3) the host-side Summary.on_sample of the product (interface parity)
    host = Summary(s.pm.T, s.pm.root, s.pm.node_to_tm, s.pm.primary_to_tol, M.get_Q_primary())
    for pri, tols in tracks:
        host.on_sample(pri, tols)
    assert host.root_xon_count == orc.root_xon_count
    for e in s.pm.edges:
        assert_allclose(host.edge_to_off_xon_dwell[e], orc.edge_to_off_xon_dwell[e], rtol=1e-12, atol=1e-12)
    # (4) EM objective pieces on the device summary vs the oracle's em on its own summary
    P = 6
    Q = M.get_Q_primary()
    pre_Q = np.zeros((P, P))
    for a, b in Q.edges():
        pre_Q[a, b] = Q[a][b]['weight']
    pd = M.get_primary_distn()
    distn = np.array([pd[i] for i in range(P)])
    on, off = M.get_rate_on(), M.get_rate_off()
    use = [(e, r) for e, r in M.get_edge_to_rate().items() if r]
    edges, rates = [e for e, r in use], [r for e, r in use]
    assert_allclose(em.get_ll_root(fused, distn, on, off), bo.get_ll_root(orc, distn, on, off), rtol=1e-10)
    assert_allclose(em.get_ll_dwell(fused, pre_Q, distn, on, off, edges, rates),
            bo.get_ll_dwell(orc, pre_Q, distn, on, off, edges, rates), rtol=1e-10)
    assert_allclose(em.get_ll_trans(fused, pre_Q, distn, on, off, edges, rates),
            bo.get_ll_trans(orc, pre_Q, distn, on, off, edges, rates), rtol=1e-10)
    six = (fused.root_xon_count, fused.root_off_count,
            sum(fused.edge_to_off_xon_trans.values()), sum(fused.edge_to_xon_off_trans.values()),
            sum(fused.edge_to_off_xon_dwell.values()), sum(fused.edge_to_xon_off_dwell.values()))
    six_o = (orc.root_xon_count, orc.root_off_count,
            sum(orc.edge_to_off_xon_trans.values()), sum(orc.edge_to_xon_off_trans.values()),
            sum(orc.edge_to_off_xon_dwell.values()), sum(orc.edge_to_xon_off_dwell.values()))
    assert_allclose(maxlikelihood.ll_blink_contribution(*(six + (on, off))),
            bo.ll_blink_contribution(*(six_o + (on, off))), rtol=1e-10)
    s.close()


def test_p53_shape_fixed_trajectory_statistics(nb):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 3, seed=0, reference_leaf=leaf)
    s = nb.BlinkSampler(pm, validate=True)
    s.set_data(packing.pack_alignment(pm, ls, tt, tf))
    s.init_chains(chains=2, seed=1)
    s.sweep(4, accumulate=False)
    s.reset_summary()
    s.sweep(1, accumulate=True)
    fused = s.summary(Q_primary=m.get_Q_primary())
    tracks = [s.export_tracks(u) for u in range(6)]
    orc = _oracle_summary(pm, m, tracks)
    _compare(fused, orc, pm)
    # p53 class names are the integers 0..19: summary.py:221 then skips samples whose
    # root codon is in class 1 (True == 1) -- reproduced bit for bit
    assert fused.root_xon_count == orc.root_xon_count
    s.close()
