"""Vanilla compound-state Rao-Teh (nxblink/compoundb.py:196-296) on the device as an
independent statistical cross-check: the one-track chain over the 24 feasible compound
states never runs the tolerance-update code, yet its folded statistics must agree with the
exact posterior expectations (and hence with the CTBN sampler, tests/test_gpu_toy_exact.py).
Tolerance: |mean - exact| <= 5 SE + 2e-3 over independent seeds (Monte Carlo)."""
import numpy as np
import pytest

from nxblink_b200 import compoundb, dense, packing, toymodel, toydata

pytestmark = pytest.mark.gpu


def _flat(d, pm):
    out = {}
    for s in pm.states:
        out['root_pri', s] = d['root_pri_to_count'][s]
    for key in ('root_on_total', 'root_off_count', 'root_xon_intended'):
        out[key] = d[key]
    for e in pm.edges:
        for key in ('off_xon_trans', 'xon_off_trans', 'off_xon_dwell', 'xon_off_dwell'):
            out[key, e] = d['edge_to_' + key][e]
        for key in ('pri_trans', 'pri_dwell'):
            for sa in d['edge_to_' + key][e]:
                for sb, v in d['edge_to_' + key][e][sa].items():
                    out[key, e, sa, sb] = v
    return out


@pytest.mark.parametrize('M,D', [(toymodel.BlinkModelB, toydata.DataC),
    (toymodel.BlinkModelA, toydata.DataA), (toymodel.BlinkModelC, toydata.DataD)],
    ids=['B-C', 'A-A', 'C-D'])
def test_compound_chain_matches_exact_posterior(nb, M, D):
    cp = dense.CompoundProcess(M)
    data = packing.pack_data(cp.pm, [D])
    exact = _flat(cp.expected_summary(data), cp.pm)
    runs = []
    nseeds = 6
    for seed in range(nseeds):
        cs = compoundb.CompoundSampler(M, validate=True)
        cs.set_data(data)
        cs.init_chains(chains=1024, seed=50 + seed)
        cs.sweep(40, accumulate=False)
        cs.sweep(40, accumulate=True)
        runs.append(_flat(cs.blink_summary(), cp.pm))
        cs.close()
    worst = (0.0, None)
    for key, ex in exact.items():
        vals = np.array([r[key] for r in runs])
        se = vals.std(ddof=1) / np.sqrt(nseeds)
        z = abs(vals.mean() - ex) / (5 * se + 2e-3)
        if z > worst[0]:
            worst = (z, (key, vals.mean(), ex, se))
    assert worst[0] <= 1.0, worst


def test_gen_raoteh_samples_yields_consistent_compound_tracks(nb):
    M, D = toymodel.BlinkModelB, toydata.DataC
    p2t = dict((p, int(t[1:])) for p, t in M.get_primary_to_tol().items())
    n = 0
    for track in compoundb.gen_raoteh_samples(M, D, 5, 12, seed=3):
        n += 1
        for v, s in track.history.items():
            assert compoundb.state_is_consistent(p2t, s)
        for edge, evs in track.events.items():
            for ev in evs:
                assert ev.sa != ev.sb and compoundb.state_is_consistent(p2t, ev.sb)
                # exactly one component changes per event
                assert sum(a != b for a, b in zip(ev.sa.tolist(), ev.sb.tolist())) == 1
    assert n == 12
