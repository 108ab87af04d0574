"""The reference's call surface on the GPU path: raoteh.gen_samples yields the
same kind of objects as nxblink/raoteh.py:25-157, and the reference's own smoke
test (nxblink/tests/test_smoke.py:19-30) passes against it."""
import pytest

from nxblink_b200 import raoteh, toymodel, toydata
from nxblink_b200.summary import Summary
from nxblink_b200.packing import PackedModel

pytestmark = pytest.mark.gpu


def test_ctbn_raoteh_sampling_smoke(nb):
    k = 4
    for model in (toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC):
        for data in (toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD):
            nsampled = 0
            for primary_track, tolerance_tracks in raoteh.gen_samples(
                    model, data, k, k * k, seed=1):
                nsampled += 1
            assert nsampled == k * k


def test_yielded_tracks_have_the_reference_shape(nb):
    M, D = toymodel.BlinkModelB, toydata.DataC
    pm = PackedModel(M)
    T, root = M.get_T_and_root()
    summary = Summary(T, root, pm.node_to_tm, M.get_primary_to_tol(), M.get_Q_primary())
    first = None
    for pri, tols in raoteh.gen_samples(M, D, 5, 20, seed=7, validate=True):
        if first is None:
            first = (pri, tols)
        assert pri is first[0] and tols is first[1]          # same mutable objects (raoteh.py:152)
        assert pri.name == 'PRIMARY' and sorted(t.name for t in tols) == ['T0', 'T1', 'T2']
        assert set(pri.history) == set(T)
        for edge in T.edges():
            for ev in pri.events[edge]:
                assert ev.track is pri and ev.sa != ev.sb
                assert M.get_Q_primary().has_edge(ev.sa, ev.sb)
            for t in tols:
                for ev in t.events[edge]:
                    assert ev.track is t and {ev.sa, ev.sb} == {False, True}
        # data constraints hold (DataC: N0 observed 0, T1 off at N0)
        assert pri.history['N0'] == 0 and pri.history['N5'] == 1
        assert [t for t in tols if t.name == 'T1'][0].history['N0'] is False
        summary.on_sample(pri, tols)
    assert summary.nsamples == 20
    assert summary.root_xon_count + summary.root_off_count == 3 * 20


def test_infeasible_data_raises(nb):
    """raoteh.py:201 'neither True nor False is feasible' / FFBS with no support."""
    class Empty(toydata.DataB):
        TOL = dict(toydata.DataB.TOL, T1=dict(N3=()))
    with pytest.raises(Exception):
        list(raoteh.gen_samples(toymodel.BlinkModelA, Empty, 1, 1, seed=1))

    class Contradiction(toydata.DataB):
        # N3 observed in state 4 (class T2) while T2 is declared OFF there
        TOL = dict(toydata.DataB.TOL, T2=dict(N3=(False,), N4=(True,)))
    with pytest.raises(Exception):
        list(raoteh.gen_samples(toymodel.BlinkModelA, Contradiction, 1, 1, seed=1))


def test_mean_ell_contribs_from_device_statistics(nb):
    """summary.mean_ell_contribs (device statistics) == the average of the reference's
    per-sample functionals (summary.py:247-348) over the same trajectories, exported."""
    import numpy as np
    from numpy.testing import assert_allclose
    from nxblink_b200 import model as nbmodel
    from nxblink_b200.summary import (WeightGraph, get_ell_init_contrib, get_ell_dwell_contrib,
            get_ell_trans_contrib, mean_ell_contribs)
    for M, D in ((toymodel.BlinkModelB, toydata.DataC), (toymodel.BlinkModelA, toydata.DataA)):
        pm = PackedModel(M)
        sampler = nb.BlinkSampler(pm)
        sampler.set_data([D])
        sampler.init_chains(chains=24, seed=5)
        sampler.sweep(6, accumulate=False)
        sampler.reset_summary()
        sampler.summarize_current()
        summary = sampler.summary()
        T, root = M.get_T_and_root()
        p2t = M.get_primary_to_tol()
        Q = WeightGraph()                      # the normalised matrix the sampler used
        for a, b, w in zip(pm.q_row, pm.q_col, pm.q_val):
            Q.add_edge(pm.states[a], pm.states[b], float(w))
        Q_meta, Q_blink = nbmodel.get_Q_meta(Q, p2t), M.get_Q_blink()
        rates = M.get_edge_to_rate()
        pos = dict((e, r) for e, r in rates.items())
        init, dwell, trans = mean_ell_contribs(summary, pm, pos, Q, M.get_rate_on(),
                M.get_rate_off(), M.get_primary_distn(), M.get_blink_distn())
        h_init, h_dwell, h_trans = 0.0, dict((e, 0.0) for e in pm.edges), dict((e, 0.0) for e in pm.edges)
        for u in range(24):
            pri, tols = sampler.export_tracks(u)
            h_init += get_ell_init_contrib(root, M.get_primary_distn(), M.get_blink_distn(),
                    pri, tols, p2t) / 24
            d = get_ell_dwell_contrib(T, root, pm.node_to_tm, rates, Q, Q_blink, Q_meta, pri, tols, p2t)
            t = get_ell_trans_contrib(T, root, rates, Q, Q_blink, pri, tols)
            for e in pm.edges:
                h_dwell[e] += d[e] / 24
                h_trans[e] += t[e] / 24
        assert_allclose(init, h_init, rtol=1e-12)
        for e in pm.edges:
            assert_allclose(dwell[e], h_dwell[e], rtol=1e-11, atol=1e-13)
            assert_allclose(trans[e], h_trans[e], rtol=1e-11, atol=1e-13)
        sampler.close()
