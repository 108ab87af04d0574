"""Helpers to read the fixtures written by oracle/refcompat/make_golden.py."""
import ast
import json
import os
from collections import defaultdict

from nxblink_b200.summary import Summary, WeightGraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def summary_from_record(rec):
    """Rebuild a Summary-shaped object from a recorded reference Summary."""
    s = Summary(None, None, None, None, None)
    s.nsamples = rec['nsamples']
    s.root_pri_to_count = defaultdict(int,
            ((ast.literal_eval(k), v) for k, v in rec['root_pri_to_count'].items()))
    s.root_xon_count = rec['root_xon_count']
    s.root_off_count = rec['root_off_count']
    for er in rec['edges']:
        e = (ast.literal_eval(er['edge'][0]), ast.literal_eval(er['edge'][1]))
        gt, gd = WeightGraph(), WeightGraph()
        for a, b, w in er['pri_trans']:
            gt.add_edge(ast.literal_eval(a), ast.literal_eval(b), w)
        for a, b, w in er['pri_dwell']:
            gd.add_edge(ast.literal_eval(a), ast.literal_eval(b), w)
        s.edge_to_pri_trans[e] = gt
        s.edge_to_pri_dwell[e] = gd
        s.edge_to_off_xon_trans[e] = er['off_xon_trans']
        s.edge_to_xon_off_trans[e] = er['xon_off_trans']
        s.edge_to_off_xon_dwell[e] = er['off_xon_dwell']
        s.edge_to_xon_off_dwell[e] = er['xon_off_dwell']
    return s


def scalar_functionals(sm):
    """Per-sample means of the scalar functionals recorded in 'batch_means'."""
    n = float(sm.nsamples)

    def total(graphs):
        t = 0
        for g in graphs.values():
            for a in g:
                for b in g[a]:
                    w = g[a][b]
                    t += w['weight'] if isinstance(w, dict) else w
        return t
    return dict(
        off_root=sm.root_off_count / n,
        xon_root=sm.root_xon_count / n,
        pri_trans=total(sm.edge_to_pri_trans) / n,
        off_xon_trans=sum(sm.edge_to_off_xon_trans.values()) / n,
        xon_off_trans=sum(sm.edge_to_xon_off_trans.values()) / n,
        off_xon_dwell=sum(sm.edge_to_off_xon_dwell.values()) / n,
        xon_off_dwell=sum(sm.edge_to_xon_off_dwell.values()) / n)
