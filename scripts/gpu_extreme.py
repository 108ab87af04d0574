"""Numerical robustness far from the fixtures: toy model B with every edge rate x30 (penalty
exponents in the tens), against the exact posterior from the dense device path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import toymodel, toydata, dense, packing
from gpu_exact_long import flat, flat_exact

class Fast(toymodel.BlinkModelB):
    @classmethod
    def get_edge_to_rate(cls):
        return dict((e, 30.0 * r) for e, r in toymodel.BlinkModelB.get_edge_to_rate().items())

for D in (toydata.DataC, toydata.DataA):
    cp = dense.CompoundProcess(Fast)
    exact = flat_exact(cp.expected_summary(packing.pack_data(cp.pm, [D])), cp.pm)
    runs = []
    for seed in range(8):
        s = nb.BlinkSampler(Fast, validate=True)
        s.set_data([D]); s.init_chains(chains=4096, seed=seed)
        s.sweep(100, accumulate=False); s.sweep(100, accumulate=True)
        runs.append(flat(s.summary(), s.pm)); st = s.stats(); s.close()
    zs = []
    for key, ex in exact.items():
        v = np.array([r[key] for r in runs]); se = v.std(ddof=1) / np.sqrt(len(runs))
        if se > 0: zs.append((abs(v.mean() - ex) / se, key, v.mean(), ex))
    zs.sort(key=lambda x: -x[0]); z = np.array([x[0] for x in zs])
    print(D.__name__, 'events per unit', (st['n_pri_events'] + st['n_blk_events']) / st['n_units'],
            'statistics', len(zs), 'max z %.2f' % z.max(), 'rms z %.2f' % np.sqrt((z ** 2).mean()), 'worst', zs[0][1:])
