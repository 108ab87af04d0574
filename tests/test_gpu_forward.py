"""Device forward sampler (nxblink/fwdsample.py:67-262 on the GPU) and the prior/posterior
consistency check it enables: with no data the Rao-Teh chain must leave the prior invariant, so
statistics of forward-simulated trajectories and of the same trajectories after Gibbs sweeps
must agree -- at the p53 shape, where no exact answer exists."""
import numpy as np
import pytest

import golden_util as gu
from oracle.dense_oracle import DensePosterior
from nxblink_b200 import toymodel, toydata, p53, packing
from test_gpu_toy_exact import _flatten, _flatten_exact

pytestmark = pytest.mark.gpu


def test_forward_statistics_match_exact_prior(nb):
    M = toymodel.BlinkModelB
    runs = []
    for seed in range(24):
        s = nb.BlinkSampler(M)
        node_pri, node_tol = s.forward_sample(40000, seed=seed)
        assert node_pri.shape == (40000, 6)
        # own class is on at every node of every sampled trajectory
        own = s.pm.state_class[node_pri]
        assert (((node_tol >> own.astype(np.uint32)) & 1) == 1).all()
        s.reset_summary()
        s.summarize_current()
        sm = s.summary()
        assert sm.nsamples == 40000
        runs.append(_flatten(sm, s.pm))
        pm = s.pm
        s.close()
    exact = _flatten_exact(DensePosterior(M, toydata.DataA).expected_summary(), pm)
    worst = (0.0, None)
    for key, ex in exact.items():
        vals = np.array([r[key] for r in runs])
        se = vals.std(ddof=1) / np.sqrt(len(runs))
        z = abs(vals.mean() - ex) / (5 * se + 2e-4)          # (24 seeds; the floor is for statistics with SE ~ 0)
        if z > worst[0]:
            worst = (z, (key, vals.mean(), ex, se))
    assert worst[0] <= 1.0, worst


def test_prior_is_invariant_under_the_sweep_p53_shape(nb):
    pm = packing.PackedModel(p53.p53_model())
    fwd, mcmc = [], []
    for seed in range(16):
        s = nb.BlinkSampler(pm, validate=(seed < 4))
        s.forward_sample(4096, seed=100 + seed)
        s.reset_summary()
        s.summarize_current()
        fwd.append(gu.scalar_functionals(s.summary()))
        s.reset_summary()
        s.sweep(6, accumulate=True)              # also validates every invariant after every phase
        sm = s.summary()
        assert sm.nsamples == 4096 * 6
        mcmc.append(gu.scalar_functionals(sm))
        s.close()
    for name in fwd[0]:
        a = np.array([r[name] for r in fwd])
        b = np.array([r[name] for r in mcmc])
        se = np.hypot(a.std(ddof=1), b.std(ddof=1)) / np.sqrt(len(a))
        assert abs(a.mean() - b.mean()) <= 5 * se + 2e-3 * max(1.0, abs(a.mean())), (name, a.mean(), b.mean(), se)
    # stationary blink flux: 2 on off / (on + off) per class and unit of (rate * length)
    flux = (np.mean([r['off_xon_trans'] for r in fwd]) + np.mean([r['xon_off_trans'] for r in fwd]))
    expect = 19 * (pm.blen * pm.rate).sum() * 2 * 1.5 * 0.5 / 2.0
    assert abs(flux - expect) < 0.02 * expect


def test_forward_sampled_alignment_can_be_analysed(nb):
    """Synthetic data straight from the blink process: sample, read the leaves, condition on them."""
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    s = nb.BlinkSampler(pm)
    node_pri, node_tol = s.forward_sample(64, seed=5)
    s.close()
    is_leaf = np.ones(pm.n_nodes, dtype=bool)
    is_leaf[pm.parent[1:]] = False
    leaf_states = np.where(is_leaf[None, :], node_pri, -1)
    data = packing.pack_alignment(pm, leaf_states)
    s = nb.BlinkSampler(pm, validate=True)
    s.set_data(data)
    s.init_chains(chains=4, seed=1)
    s.sweep(5, accumulate=True)
    assert s.summary().nsamples == 64 * 4 * 5
    s.close()


def test_rooted_forward_samples_follow_the_transition_matrix(nb):
    """fwdsample.gen_forward_samples takes the root states from the caller: with the root
    fixed, the joint (primary, tolerance) state at a child of the root is distributed as
    the row of expm(Q_compound * rate * blen) (scipy.linalg.expm, the reference's authority)."""
    import scipy.linalg
    from nxblink_b200 import dense
    M = toymodel.BlinkModelB
    cp = dense.CompoundProcess(M)
    pm = cp.pm
    n = 60000
    root_state = (2, (1, 1, 0))                   # primary 2 (class T1) with T0 on, T2 off
    r = cp.states.index(root_state)
    s = nb.BlinkSampler(pm)
    rp = np.full(n, root_state[0], dtype=np.int32)
    rt = np.full(n, sum(b << k for k, b in enumerate(root_state[1])), dtype=np.uint32)
    node_pri, node_tol = s.forward_sample(n, seed=9, root_pri=rp, root_tol=rt)
    s.close()
    assert (node_pri[:, 0] == root_state[0]).all() and (node_tol[:, 0] == rt[0]).all()
    idx = dict((c, i) for i, c in enumerate(cp.states))
    for v in range(1, pm.n_nodes):
        if pm.parent[v] != 0:
            continue
        t = pm.blen[v] * pm.rate[v]
        exact = scipy.linalg.expm(cp.Q * t)[r]
        counts = np.zeros(len(cp.states))
        codes = node_pri[:, v].astype(np.int64) * 8 + node_tol[:, v]
        for code, c in zip(*np.unique(codes, return_counts=True)):
            p, m = int(code) // 8, int(code) % 8
            counts[idx[(p, tuple((m >> k) & 1 for k in range(3)))]] = c
        freq = counts / n
        se = np.sqrt(exact * (1 - exact) / n)
        assert (np.abs(freq - exact) <= 5 * se + 1e-4).all(), (v, freq, exact)


def test_gen_forward_samples_facade(nb):
    from nxblink_b200 import fwdsample
    M = toymodel.BlinkModelA
    roots = [{'PRIMARY': 0, 'T0': 1, 'T1': 0, 'T2': 1}, {'PRIMARY': 5, 'T0': 0, 'T1': 0, 'T2': 1},
            {'PRIMARY': 3, 'T0': 1, 'T1': 1, 'T2': 1}]
    T, root = M.get_T_and_root()
    p2t = M.get_primary_to_tol()
    out = list(fwdsample.gen_forward_samples(M, roots, seed=4))
    assert len(out) == 3
    for want, (pri, tols) in zip(roots, out):
        assert pri.history[root] == want['PRIMARY']
        for t in tols:
            assert bool(t.history[root]) == bool(want[t.name])
        assert set(pri.history) == set(T)
        # every event is a move the model allows, in time order along its edge
        for edge in T.edges():
            state = dict((t.name, bool(t.history[edge[0]])) for t in tols)
            cur = pri.history[edge[0]]
            evs = sorted([ev for tr in [pri] + tols for ev in tr.events[edge]], key=lambda e: e.tm)
            for ev in evs:
                if ev.track is pri:
                    assert ev.sa == cur and state[p2t[ev.sb]] and M.get_Q_primary().has_edge(ev.sa, ev.sb)
                    cur = ev.sb
                else:
                    assert bool(ev.sa) == state[ev.track.name] and bool(ev.sb) != bool(ev.sa)
                    assert ev.sb or ev.track.name != p2t[cur]      # own class never switches off
                    state[ev.track.name] = bool(ev.sb)
            assert cur == pri.history[edge[1]]
    with pytest.raises(Exception):
        list(fwdsample.gen_forward_samples(M, [{'PRIMARY': 0, 'T0': 0, 'T1': 1, 'T2': 1}]))
