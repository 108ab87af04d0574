"""Is the 1e-3 deviation of the no-data grid a burn-in residue or a bias?  Start from exact prior
draws (device forward sampler) and sweep: stationarity must hold to within Monte Carlo error."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import toymodel, toydata, dense, packing
from gpu_exact_long import flat, flat_exact

nseeds, n, sweeps = 16, 65536, 100
for M in (toymodel.BlinkModelA, toymodel.BlinkModelC):
    cp = dense.CompoundProcess(M)
    exact = flat_exact(cp.expected_summary(packing.pack_data(cp.pm, [toydata.DataA])), cp.pm)
    for label, nsw in (('forward only', 0), ('forward + %d sweeps' % sweeps, sweeps)):
        runs = []
        for seed in range(nseeds):
            s = nb.BlinkSampler(M)
            s.forward_sample(n, seed=900 + seed)
            s.reset_summary()
            if nsw == 0:
                s.summarize_current()
            else:
                s.sweep(20, accumulate=False); s.sweep(nsw, accumulate=True)
            runs.append(flat(s.summary(), s.pm)); s.close()
        zs = []
        for key, ex in exact.items():
            v = np.array([r[key] for r in runs]); se = v.std(ddof=1) / np.sqrt(nseeds)
            if se > 0: zs.append((abs(v.mean() - ex) / se, key, v.mean(), ex, se))
        zs.sort(key=lambda x: -x[0]); z = np.array([x[0] for x in zs])
        print(M.__name__, label, 'max z %.2f' % z.max(), 'frac z>2: %.3f' % (z > 2).mean(),
                'rms z %.2f' % np.sqrt((z ** 2).mean()), 'worst', zs[0][1], 'rel err %.2e' % (abs(zs[0][2] - zs[0][3]) / abs(zs[0][3])))
