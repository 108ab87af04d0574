"""Prior invariance at the p53 shape with a pure z criterion: forward-simulated trajectories and
the same trajectories after Gibbs sweeps must have the same statistics (no exact answer exists
at this shape; the stationary flux is known in closed form)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import p53, packing
pm = packing.PackedModel(p53.p53_model())
def per_edge(sm):
    f = sm.flat; n = float(sm.nsamples)
    out = {}
    for e in range(pm.n_nodes - 1):
        out['pri_trans', e] = f['pri_trans'][e].sum() / n
        out['off_on', e] = f['off_on'][e] / n; out['on_off', e] = f['on_off'][e] / n
        out['off_dwell', e] = f['off_xon_dwell'][e] / n; out['on_dwell', e] = f['xon_off_dwell'][e] / n
    for s in range(pm.n_states): out['T', s] = f['T'][:, s].sum() / n
    out['root_off'] = sm.root_off_count / n
    return out
fwd, mc = [], []
nseeds, n, sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 12, 65536, 12
base = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for seed in range(nseeds):
    s = nb.BlinkSampler(pm)
    s.forward_sample(n, seed=base + seed); s.reset_summary(); s.summarize_current(); fwd.append(per_edge(s.summary()))
    s.reset_summary(); s.sweep(4, accumulate=False); s.sweep(sweeps, accumulate=True); mc.append(per_edge(s.summary())); s.close()
zs = []
for k in fwd[0]:
    a = np.array([r[k] for r in fwd]); b = np.array([r[k] for r in mc])
    se = np.hypot(a.std(ddof=1), b.std(ddof=1)) / np.sqrt(nseeds)
    if se > 0: zs.append((abs(a.mean() - b.mean()) / se, k, a.mean(), b.mean()))
zs.sort(key=lambda x: -x[0]); z = np.array([x[0] for x in zs])
print('statistics', len(zs), 'max z %.2f' % z.max(), 'frac z>2 %.3f' % (z > 2).mean(), 'rms z %.2f' % np.sqrt((z ** 2).mean()))
print("worst", [(round(x[0], 2), x[1], round(x[2], 5), round(x[3], 5), "%.1e" % ((x[3] - x[2]) / x[2])) for x in zs[:6]]); print("edge rate*len:", dict((e, round(float(pm.blen[e + 1] * pm.rate[e + 1]), 4)) for e in sorted(set(x[1][1] for x in zs[:6] if isinstance(x[1], tuple) and x[1][0] != "T"))))
