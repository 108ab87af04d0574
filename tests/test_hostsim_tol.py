"""The per-lane code of the tolerance update (nxblink_b200/csrc/tol_update.cuh), compiled for the
host (oracle/hostsim) and run as a Markov chain on ONE unit with a fixed primary trajectory,
against the EXACT conditional posterior of every tolerance track (oracle/tol_exact.py): node
marginals, expected off->on / on->off transitions per edge, OFF dwell per (edge, primary state,
track).  This is the target of raoteh.py:413-432; it checks the device algorithm itself (candidate
generation by thinning, the upward walk with its node accumulators, the backward draw) without a GPU.  The GPU parity tests run the same
source on the device.

Tolerance: |mean - exact| <= 5.5 batch-means standard errors (20 batches); quantities without
Monte Carlo spread (hard constraints, or events that never occurred) must match to 8 / nsweep."""
import numpy as np
import pytest

from nxblink_b200 import toymodel, packing, p53
from oracle import tol_exact
import hostsim_util as hu


def _qm(pm):
    qm = np.zeros((pm.n_states, pm.n_classes))
    for s in range(pm.n_states):
        for i in range(pm.q_ptr[s], pm.q_ptr[s + 1]):
            qm[s, pm.state_class[pm.q_col[i]]] += pm.q_val[i]
    return qm


def _leaf_constraints(pm, node_pri, rng, p_off=0.3, p_on=0.2, min_rate=0.0):
    N, K = pm.n_nodes, pm.n_classes
    allk = (1 << K) - 1
    tt = np.full(N, allk, dtype=np.uint32)
    tf = np.full(N, allk, dtype=np.uint32)
    kids = set(int(x) for x in pm.parent[1:])
    for v in range(1, N):
        if v in kids or pm.rate[v] * pm.blen[v] < min_rate:
            continue
        own = pm.state_class[node_pri[v]]
        for k in range(K):
            if k == own:
                continue
            u = rng.random()
            if u < p_off:
                tt[v] &= ~np.uint32(1 << k)       # observed OFF
            elif u < p_off + p_on:
                tf[v] &= ~np.uint32(1 << k)       # observed ON
    return tt, tf


def _check_trajectory(ch):
    """Blink events sorted by (track, edge, time), alternate, and connect the node states."""
    pm = ch.pm
    t, e, k, ns = ch.blink_events()
    key = list(zip(k.tolist(), e.tolist(), t.tolist()))
    assert key == sorted(key) and len(set(key)) == len(key)
    for kk in range(pm.n_classes):
        for ed in range(pm.n_nodes - 1):
            x = (int(ch.node_tol[pm.parent[ed + 1]]) >> kk) & 1
            lo, hi = pm.node_tm[pm.parent[ed + 1]], pm.node_tm[ed + 1]
            for tm, ee, k2, n2 in zip(t, e, k, ns):
                if k2 == kk and ee == ed:
                    assert lo < tm < hi
                    assert n2 != x
                    x = n2
            assert x == (int(ch.node_tol[ed + 1]) >> kk) & 1


def _run(pm, msg_f64, nsweep, seed, nsub=32, nbatch=20, min_rate=0.0, burn=100):
    rng = np.random.default_rng(seed)
    node_pri, ev = hu.forward_primary_path(pm, rng)
    tt, tf = _leaf_constraints(pm, node_pri, rng, min_rate=min_rate)
    ch = hu.HostTolChain(pm, node_pri, ev, tt, tf, msg_f64=msg_f64, seed=seed)
    N, E, P, K = pm.n_nodes, pm.n_nodes - 1, pm.n_states, pm.n_classes
    for _ in range(burn):
        ch.sweep(False)
    per = nsweep // nbatch
    node_on = np.zeros((nbatch, N, K))
    doff = np.zeros((nbatch, E * P * K))
    up = np.zeros((nbatch, E))
    dn = np.zeros((nbatch, E))
    shifts = np.arange(K)[None, :]
    for b in range(nbatch):
        d0, u0, n0 = ch.doff.copy(), ch.off_on.astype(float), ch.on_off.astype(float)
        for _ in range(per):
            ch.sweep(True)
            node_on[b] += (ch.node_tol[:, None] >> shifts) & 1
        _check_trajectory(ch)
        doff[b] = (ch.doff - d0) / per
        up[b] = (ch.off_on.astype(float) - u0) / per
        dn[b] = (ch.on_off.astype(float) - n0) / per
    node_on /= per
    qm = _qm(pm)
    ex_on = np.zeros((N, K))
    ex_up, ex_dn = np.zeros(E), np.zeros(E)
    ex_doff = np.zeros((E, P, K))
    for k in range(K):
        ex = tol_exact.track_posterior(pm.parent, pm.blen, pm.rate, node_pri, ev, k,
                pm.state_class, qm, pm.rate_on, pm.rate_off, tt, tf, nsub=nsub)
        ex_on[:, k] = ex['node_on']
        ex_up += ex['off_on']
        ex_dn += ex['on_off']
        for (e, s), v in ex['off_dwell'].items():
            ex_doff[e, s, k] = v

    def worst(got, exact, what):
        m = got.mean(axis=0)
        se = got.std(axis=0, ddof=1) / np.sqrt(nbatch)
        # no spread over the batches: a hard constraint (exact match), or an event so rare that it
        # never happened in nsweep draws (its exact expectation must then be below ~8 / nsweep)
        # (the same bound where fewer than 5 of the batches saw the event at all: a standard
        # error from one or two occurrences is no basis for a z-score)
        det = (se < 1e-12) | ((got != 0).sum(axis=0) < 5)
        assert np.all(np.abs(m - exact)[det] < 8.0 / nsweep), (what, np.abs(m - exact)[det].max())
        z = np.abs(m - exact)[~det] / se[~det]
        return float(z.max()) if z.size else 0.0

    zs = dict(node_on=worst(node_on.reshape(nbatch, -1), ex_on.ravel(), 'node states'),
            off_on=worst(up, ex_up, 'off->on'), on_off=worst(dn, ex_dn, 'on->off'),
            off_dwell=worst(doff, ex_doff.ravel(), 'OFF dwell'))
    return zs


@pytest.mark.parametrize('msg_f64', [False, True], ids=['f32', 'f64'])
@pytest.mark.parametrize('M', [toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC],
        ids=lambda m: m.__name__)
def test_tolerance_update_matches_exact_two_state_posterior(M, msg_f64):
    pm = packing.PackedModel(M)
    zs = _run(pm, msg_f64, nsweep=40000, seed=17 + int(msg_f64))
    assert max(zs.values()) <= 5.5, zs


def test_tolerance_update_p53_shape():
    """49-node tree, 61 states, 20 classes: every (node, class) marginal and every OFF dwell cell.
    (Observations only at leaves whose branch carries rate x length >= 0.05: conflicting
    observations across near-zero branches make the posterior multi-modal, and the chain -- like
    the reference's -- then needs ~10^3 sweeps per mode switch, which a unit test cannot afford.)"""
    pm = packing.PackedModel(p53.p53_model())
    zs = _run(pm, False, nsweep=8000, seed=4, nsub=8, min_rate=0.05, burn=1000)
    assert max(zs.values()) <= 5.5, zs


@pytest.mark.parametrize('msg_f64', [False, True], ids=['f32', 'f64'])
def test_event_driven_downward_pass_samples_the_same_trajectory_as_the_full_walk(msg_f64):
    """tol_down_ev (visits only the edges with live events, the rest is filled in per unit) against
    tol_down_walk (every edge of every track): same random draws in the same order, so node states,
    blink events and integer statistics are IDENTICAL sweep after sweep; the OFF dwell sums agree to
    rounding (the additions come in another order)."""
    for pm, seed in ((packing.PackedModel(p53.p53_model()), 11), (packing.PackedModel(toymodel.BlinkModelC), 5)):
        rng = np.random.default_rng(seed)
        node_pri, ev = hu.forward_primary_path(pm, rng)
        tt, tf = _leaf_constraints(pm, node_pri, rng, min_rate=0.05)
        a = hu.HostTolChain(pm, node_pri, ev, tt, tf, msg_f64=msg_f64, seed=seed)
        b = hu.HostTolChain(pm, node_pri, ev, tt, tf, msg_f64=msg_f64, seed=seed, walk=True)
        for i in range(300):
            a.sweep(True)
            b.sweep(True)
            assert np.array_equal(a.node_tol, b.node_tol), i
            ea, eb = a.blink_events(), b.blink_events()
            assert all(np.array_equal(x, y) for x, y in zip(ea, eb)), i
        assert np.array_equal(a.off_on, b.off_on) and np.array_equal(a.on_off, b.on_off)
        assert a.doff.sum() > 0
        np.testing.assert_allclose(a.doff, b.doff, rtol=1e-11, atol=1e-12)
