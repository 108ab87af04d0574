"""Generate the golden fixtures under tests/golden/ by executing the REFERENCE's
own files (oracle/refcompat/loader.py) in this container.

TEST INFRASTRUCTURE ONLY.  Usage (from the repo root, where /root/reference exists):

    python -m oracle.refcompat.make_golden toy      # tests/golden/ref_toy.json
    python -m oracle.refcompat.make_golden p53 0    # tests/golden/ref_p53_site0.json
    python -m oracle.refcompat.make_golden p53long 0 [nsamples]   # tests/golden/ref_p53_long_site0.json
    python -m oracle.refcompat.make_golden mg94     # tests/golden/ref_mg94.json
    python -m oracle.refcompat.make_golden ell      # tests/golden/ref_ell_contrib.json

The sampled statistics are single recorded runs of the reference sampler (its RNG
is the global numpy.random, seeded here, but its consumption order depends on
Python set iteration order, so the run is a fixture rather than a reproducible
stream); every fixture records its sample count and batch-means standard errors.
The deterministic values (em.get_ll_*, maxlikelihood, MG94 tables) are exact.
"""
from __future__ import division, print_function

import json
import os
import sys
import time

import numpy as np
import networkx as nx

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from oracle.refcompat import loader  # noqa: E402


class NxAdapter(object):
    """Present a nxblink_b200-style model (dict graphs) with networkx objects."""

    def __init__(self, model):
        self._m = model

    def __getattr__(self, name):
        return getattr(self._m, name)

    def get_T_and_root(self):
        T, root = self._m.get_T_and_root()
        G = nx.DiGraph()
        G.add_nodes_from(T)
        for a in T:
            for b in T[a]:
                G.add_edge(a, b)
        return G, root

    def get_Q_primary(self):
        Q = self._m.get_Q_primary()
        G = nx.DiGraph()
        for a in Q:
            for b in Q[a]:
                G.add_edge(a, b, weight=Q[a][b]['weight'])
        return G


def summary_to_json(summary, edges, states, per_sample=None):
    out = dict(nsamples=summary.nsamples,
            root_pri_to_count=dict((repr(k), int(v)) for k, v in summary.root_pri_to_count.items()),
            root_xon_count=int(summary.root_xon_count),
            root_off_count=int(summary.root_off_count), edges=[])
    for e in edges:
        tr, dw = summary.edge_to_pri_trans[e], summary.edge_to_pri_dwell[e]
        out['edges'].append(dict(
            edge=[repr(e[0]), repr(e[1])],
            pri_trans=[[repr(a), repr(b), tr[a][b]['weight']] for a, b in tr.edges()],
            pri_dwell=[[repr(a), repr(b), dw[a][b]['weight']] for a, b in dw.edges()],
            off_xon_trans=int(summary.edge_to_off_xon_trans[e]),
            xon_off_trans=int(summary.edge_to_xon_off_trans[e]),
            off_xon_dwell=float(summary.edge_to_off_xon_dwell[e]),
            xon_off_dwell=float(summary.edge_to_xon_off_dwell[e])))
    if per_sample is not None:
        out['per_sample'] = per_sample
    return out


def scalar_series(summary, edges):
    """A few scalar functionals of the running summary, for batch-means errors."""
    return dict(
        off_root=summary.root_off_count,
        xon_root=summary.root_xon_count,
        pri_trans=sum(d['weight'] for e in edges
            for a, b, d in summary.edge_to_pri_trans[e].edges(data=True)),
        off_xon_trans=sum(summary.edge_to_off_xon_trans[e] for e in edges),
        xon_off_trans=sum(summary.edge_to_xon_off_trans[e] for e in edges),
        off_xon_dwell=sum(summary.edge_to_off_xon_dwell[e] for e in edges),
        xon_off_dwell=sum(summary.edge_to_xon_off_dwell[e] for e in edges))


def run_reference(model, data, nburnin, nsamples, seed):
    pkg = loader.load()
    from nxblink.raoteh import gen_samples
    from nxblink.summary import Summary
    from nxblink.util import get_node_to_tm
    np.random.seed(seed)
    T, root = model.get_T_and_root()
    node_to_tm = get_node_to_tm(T, root, model.get_edge_to_blen())
    summary = Summary(T, root, node_to_tm, model.get_primary_to_tol(), model.get_Q_primary())
    edges = list(T.edges())
    series = []
    for pri, tols in gen_samples(model, data, nburnin, nsamples):
        summary.on_sample(pri, tols)
        series.append(scalar_series(summary, edges))
    return summary, edges, series


def batch_means(series, nbatches=20):
    """Per-sample mean and batch-means standard error of cumulative scalar series."""
    n = len(series)
    out = {}
    for key in series[0]:
        cum = np.array([0.0] + [s[key] for s in series])
        per = np.diff(cum)
        b = n // nbatches
        means = per[:b * nbatches].reshape(nbatches, b).mean(axis=1)
        out[key] = dict(mean=float(per.mean()), se=float(means.std(ddof=1) / np.sqrt(nbatches)))
    return out


def make_toy(nsamples=4000, nburnin=100):
    loader.load()
    from nxblink.toymodel import BlinkModelA, BlinkModelB, BlinkModelC
    from nxblink.toydata import DataA, DataB, DataC, DataD
    from nxblink import em, maxlikelihood
    out = {}
    for M in (BlinkModelA, BlinkModelB, BlinkModelC):
        for D in (DataA, DataB, DataC, DataD):
            t0 = time.time()
            summary, edges, series = run_reference(M, D, nburnin, nsamples, seed=12345)
            rec = summary_to_json(summary, edges, range(6))
            rec['batch_means'] = batch_means(series)
            # deterministic quantities computed BY THE REFERENCE on this summary
            Q = M.get_Q_primary()
            pre_Q = np.zeros((6, 6))
            for a, b in Q.edges():
                pre_Q[a, b] = Q[a][b]['weight']
            pd = M.get_primary_distn()
            distn = np.array([pd[i] for i in range(6)])
            on, off = M.get_rate_on(), M.get_rate_off()
            edge_rates = [M.get_edge_to_rate()[e] for e in edges]
            use = [(e, r) for e, r in zip(edges, edge_rates) if r]
            rec['em'] = dict(
                ll_root=float(em.get_ll_root(summary, distn, on, off)),
                ll_dwell=float(em.get_ll_dwell(summary, pre_Q, distn, on, off,
                    [e for e, r in use], [r for e, r in use])),
                ll_trans=float(em.get_ll_trans(summary, pre_Q, distn, on, off,
                    [e for e, r in use], [r for e, r in use])),
                edges=[[repr(e[0]), repr(e[1])] for e, r in use],
                edge_rates=[r for e, r in use])
            six = (summary.root_xon_count, summary.root_off_count,
                    sum(summary.edge_to_off_xon_trans[e] for e in edges),
                    sum(summary.edge_to_xon_off_trans[e] for e in edges),
                    sum(summary.edge_to_off_xon_dwell[e] for e in edges),
                    sum(summary.edge_to_xon_off_dwell[e] for e in edges))
            rec['maxlikelihood'] = dict(
                six=[float(x) for x in six],
                ll_at_model=float(maxlikelihood.ll_blink_contribution(*(six + (on, off)))),
                grad_at_model=[float(x) for x in
                    maxlikelihood.ll_blink_contribution_gradient(*(six + (on, off)))],
                analytic_mle=[float(x) for x in
                    maxlikelihood.get_blink_rate_analytical_mle(*six)],
                numeric_mle=[float(x) for x in maxlikelihood.get_blink_rate_mle(*six)])
            out['%s/%s' % (M.__name__, D.__name__)] = rec
            print(M.__name__, D.__name__, 'done in %.1fs' % (time.time() - t0))
    with open(os.path.join(GOLDEN, 'ref_toy.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


def p53_inputs(n_sites=4, seed=0):
    from nxblink_b200 import p53, packing
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    ref_leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, n_sites, seed=seed, reference_leaf=ref_leaf)
    return m, pm, ls, tt, tf


class MaskData(object):
    """Reference-style data object from packed masks of one site."""

    def __init__(self, pm, ls, tt, tf):
        self.pm, self.ls, self.tt, self.tf = pm, ls, tt, tf

    def get_data(self):
        pm = self.pm
        data = {'PRIMARY': {}}
        for k, name in enumerate(pm.tols):
            data[name] = {}
        for v, node in enumerate(pm.nodes):
            s = self.ls[v]
            data['PRIMARY'][node] = set(pm.states) if s < 0 else {pm.states[s]}
            for k, name in enumerate(pm.tols):
                fset = set()
                if (int(self.tt[v]) >> k) & 1:
                    fset.add(True)
                if (int(self.tf[v]) >> k) & 1:
                    fset.add(False)
                if s >= 0 and pm.state_class[s] == k:
                    fset.discard(False)
                data[name][node] = fset
        return data


def make_p53(site, nsamples=600, nburnin=30):
    m, pm, ls, tt, tf = p53_inputs()
    data = MaskData(pm, ls[site], tt[site], tf[site])
    t0 = time.time()
    summary, edges, series = run_reference(NxAdapter(m), data, nburnin, nsamples, seed=777 + site)
    rec = summary_to_json(summary, edges, pm.states)
    rec['batch_means'] = batch_means(series)
    rec['site'] = site
    rec['seconds'] = time.time() - t0
    rec['leaf_states'] = [int(x) for x in ls[site]]
    rec['tol_true_ok'] = [int(x) for x in tt[site]]
    rec['tol_false_ok'] = [int(x) for x in tf[site]]
    rec['node_order'] = [repr(n) for n in pm.nodes]
    with open(os.path.join(GOLDEN, 'ref_p53_site%d.json' % site), 'w') as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print('p53 site', site, 'done in %.1fs' % rec['seconds'])


def edge_vector(summary, edges):
    """Per-edge cumulative statistics of the running summary (one row per edge)."""
    return np.array([[
        sum(d['weight'] for a, b, d in summary.edge_to_pri_trans[e].edges(data=True)),
        sum(d['weight'] for a, b, d in summary.edge_to_pri_dwell[e].edges(data=True)),
        summary.edge_to_off_xon_trans[e], summary.edge_to_xon_off_trans[e],
        summary.edge_to_off_xon_dwell[e], summary.edge_to_xon_off_dwell[e]] for e in edges],
        dtype=np.float64)


EDGE_STATS = ('pri_trans', 'pri_dwell', 'off_xon_trans', 'xon_off_trans', 'off_xon_dwell',
        'xon_off_dwell')


def make_p53_long(site, nsamples=10000, nburnin=50, nbatches=50):
    """A long run of the reference sampler at the p53 shape with PER-EDGE statistics and
    batch-means standard errors (tests/golden/ref_p53_long_site*.json): the sharp parity gate
    for the sampled quantities at the p53 shape (the 600-sample fixtures only carry 7 scalars)."""
    loader.load()
    from nxblink.raoteh import gen_samples
    from nxblink.summary import Summary
    from nxblink.util import get_node_to_tm
    m, pm, ls, tt, tf = p53_inputs()
    data = MaskData(pm, ls[site], tt[site], tf[site])
    model = NxAdapter(m)
    t0 = time.time()
    np.random.seed(7770 + site)
    T, root = model.get_T_and_root()
    node_to_tm = get_node_to_tm(T, root, model.get_edge_to_blen())
    summary = Summary(T, root, node_to_tm, model.get_primary_to_tol(), model.get_Q_primary())
    edges = [(pm.nodes[pm.parent[v]], pm.nodes[v]) for v in range(1, pm.n_nodes)]
    b = nsamples // nbatches
    marks, roots = [edge_vector(summary, edges)], [(0.0, 0.0)]
    n = 0
    for pri, tols in gen_samples(model, data, nburnin, b * nbatches):
        summary.on_sample(pri, tols)
        n += 1
        if n % b == 0:
            marks.append(edge_vector(summary, edges))
            roots.append((float(summary.root_off_count), float(summary.root_xon_count)))
            print('site', site, n, 'samples, %.0fs' % (time.time() - t0))
            sys.stdout.flush()
    per = np.diff(np.array(marks), axis=0) / b            # [nbatches, E, 6] per-sample means
    rper = np.diff(np.array(roots), axis=0) / b
    rec = dict(site=site, nsamples=b * nbatches, nburnin=nburnin, nbatches=nbatches,
            seconds=time.time() - t0, stats=list(EDGE_STATS),
            edge_mean=per.mean(axis=0).tolist(),
            edge_se=(per.std(axis=0, ddof=1) / np.sqrt(nbatches)).tolist(),
            root_mean=dict(off_root=float(rper[:, 0].mean()), xon_root=float(rper[:, 1].mean())),
            root_se=dict(off_root=float(rper[:, 0].std(ddof=1) / np.sqrt(nbatches)),
                xon_root=float(rper[:, 1].std(ddof=1) / np.sqrt(nbatches))),
            root_pri_to_count=dict((repr(k), int(v)) for k, v in summary.root_pri_to_count.items()),
            leaf_states=[int(x) for x in ls[site]],
            tol_true_ok=[int(x) for x in tt[site]], tol_false_ok=[int(x) for x in tf[site]],
            node_order=[repr(x) for x in pm.nodes])
    with open(os.path.join(GOLDEN, 'ref_p53_long_site%d.json' % site), 'w') as f:
        json.dump(rec, f, sort_keys=True)
    print('p53 long site', site, 'done in %.1fs' % rec['seconds'])


def make_ell(nsamples=6, nburnin=5):
    """Trajectories sampled by the reference together with the reference's own per-sample
    functionals of them (summary.py:247-348 as driven by toymodel-stochastic.py:137-196)."""
    loader.load()
    from nxblink.toymodel import BlinkModelB, BlinkModelC
    from nxblink.toydata import DataA, DataC
    from nxblink.raoteh import gen_samples
    from nxblink.model import get_Q_meta
    from nxblink.util import get_node_to_tm
    from nxblink.navigation import gen_segments
    from nxblink.summary import (get_ell_init_contrib, get_ell_dwell_contrib,
            get_ell_trans_contrib)
    out = {}
    for M, D in ((BlinkModelB, DataC), (BlinkModelC, DataA)):
        np.random.seed(4242)
        T, root = M.get_T_and_root()
        edges = list(T.edges())
        node_to_tm = get_node_to_tm(T, root, M.get_edge_to_blen())
        edge_to_rate = M.get_edge_to_rate()
        primary_to_tol = M.get_primary_to_tol()
        Q_primary = M.get_Q_primary()
        Q_blink = M.get_Q_blink()
        Q_meta = get_Q_meta(Q_primary, primary_to_tol)
        samples = []
        for pri, tols in gen_samples(M, D, nburnin, nsamples):
            tracks = []
            for t in [pri] + tols:
                tracks.append(dict(name=t.name,
                    history=dict((v, int(x)) for v, x in t.history.items()),
                    events=[[list(e), [[float(ev.tm), int(ev.sa), int(ev.sb)] for ev in t.events[e]]]
                        for e in edges]))
            d = get_ell_dwell_contrib(T, root, node_to_tm, edge_to_rate, Q_primary, Q_blink,
                    Q_meta, pri, tols, primary_to_tol)
            tr = get_ell_trans_contrib(T, root, edge_to_rate, Q_primary, Q_blink, pri, tols)
            segs = [[float(a), float(b), sorted((str(k), int(v)) for k, v in st.items())]
                    for a, b, st in gen_segments(edges[0], node_to_tm, [pri] + tols)]
            samples.append(dict(tracks=tracks,
                ell_init=float(get_ell_init_contrib(root, M.get_primary_distn(),
                    M.get_blink_distn(), pri, tols, primary_to_tol)),
                ell_dwell=[[list(e), float(d[e])] for e in edges],
                ell_trans=[[list(e), float(tr[e])] for e in edges],
                segments_edge0=segs))
        out['%s/%s' % (M.__name__, D.__name__)] = dict(
                samples=samples,
                # gen_samples normalised this graph in place (raoteh.py:101-102); record what
                # the functionals above actually saw
                Q_primary=[[int(a), int(b), float(Q_primary[a][b]['weight'])]
                    for a, b in Q_primary.edges()])
    with open(os.path.join(GOLDEN, 'ref_ell_contrib.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


def make_mg94():
    """MG94 tables from the reference's create_mg94.py, for the packing test."""
    loader.load()
    sys.path.insert(0, os.path.join(loader.REFERENCE_ROOT, 'examples', 'p53'))
    import create_mg94 as ref_mg94
    from nxblink_b200 import p53
    gc = p53.universal_genetic_code()
    prm = p53.jeff_params()
    Q, distn, s2r, r2p = ref_mg94.create_mg94(prm['A'], prm['C'], prm['G'], prm['T'],
            prm['kappa'], prm['omega'], gc, target_expected_rate=1.0)
    rec = dict(Q=[[a, b, Q[a][b]['weight']] for a, b in sorted(Q.edges())],
            distn=[distn[i] for i in range(61)],
            residue_to_part=r2p, genetic_code=[[s, r, c] for s, r, c in gc])
    with open(os.path.join(GOLDEN, 'ref_mg94.json'), 'w') as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print('mg94 done: %d rates' % len(rec['Q']))


if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'toy':
        make_toy()
    elif what == 'p53':
        make_p53(int(sys.argv[2]))
    elif what == 'p53long':
        make_p53_long(int(sys.argv[2]), *[int(x) for x in sys.argv[3:]])
    elif what == 'mg94':
        make_mg94()
    elif what == 'ell':
        make_ell()
