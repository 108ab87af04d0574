"""Where the time of one EM iteration goes at the real p53 size (393 columns x 64 chains)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import p53, packing, p53em

m = p53.p53_model()
pm = packing.PackedModel(m)
leaf = pm.node_index[m.name_to_leaf['Has']]
ls, tt, tf = p53.synthetic_alignment(pm, 393, seed=1, reference_leaf=leaf)
data = packing.pack_alignment(pm, ls, tt, tf)
def t(label, f):
    t0 = time.perf_counter(); r = f(); print('%-28s %8.2f ms' % (label, 1e3 * (time.perf_counter() - t0))); return r
for rep in range(2):
    print('--- rep', rep)
    s = t('BlinkSampler (nxb_create)', lambda: nb.BlinkSampler(pm))
    t('set_data', lambda: s.set_data(data))
    t('init_chains(64)', lambda: s.init_chains(chains=64, seed=rep))
    t('10 burn-in sweeps', lambda: s.sweep(10, accumulate=False))
    t('10 sampling sweeps', lambda: s.sweep(10, accumulate=True))
    sm = t('summary()', lambda: s.summary())
    t('close', lambda: s.close())
    prm = p53.jeff_params()
    ms = t('MStep()', lambda: p53em.MStep(pm, sm, m.genetic_code))
    t('maximization_step', lambda: p53em.maximization_step(ms, prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'], 1.5, 0.5, np.maximum(pm.rate[1:], 1e-6)))
