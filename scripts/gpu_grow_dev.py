"""Development check of the grow-and-continue path with statistics accumulated from the first sweep."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import nxblink_b200 as nb
from test_gpu_invariance import _p53_inputs
pm, data = _p53_inputs(24)
def run(nsw=9, **opts):
    s = nb.BlinkSampler(pm, **opts)
    s.set_data(data); s.init_chains(chains=8, seed=5)
    d0 = s.stats()['capacity_doublings']
    s.sweep(nsw, accumulate=True)
    c, d = s.read_summary(); st = s.stats(); L = s.layout; s.close()
    return c, d, st, d0, L
for nsw in (1, 2, 3, 9):
    base = run(nsw)
    L = base[4]
    names = [(L.c_root_pri, 'root_pri'), (L.c_root_on, 'root_on'), (L.c_root_off, 'root_off'), (L.c_root_on_cls, 'root_on_cls'),
             (L.c_pri_trans, 'pri_trans'), (L.c_off_on, 'off_on'), (L.c_on_off, 'on_off'), (L.c_nsamples, 'nsamples')]
    for kw in ({}, {'split_phases': -1}, {'sync': 1}):
        if kw.get('sync'):
            os.environ['NXB_SYNC_SWEEPS'] = '1'; kw = {}
        else:
            os.environ.pop('NXB_SYNC_SWEEPS', None)
        t = run(nsw, cap_scale=0.06, **kw)
        diff = np.nonzero(base[0] != t[0])[0]
        where = {}
        for i in diff:
            nm = [n for o, n in names if o <= i][-1]
            where.setdefault(nm, []).append(int(t[0][i]) - int(base[0][i]))
        print('nsw', nsw, kw, 'sync' if os.environ.get('NXB_SYNC_SWEEPS') else '', 'doublings', t[3], t[2]['capacity_doublings'], 'ndiff', len(diff),
              dict((k, (len(v), sum(v))) for k, v in where.items()), 'dwell diff', np.abs(t[1] - base[1]).max())
