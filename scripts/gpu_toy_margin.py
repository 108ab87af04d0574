import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import nxblink_b200 as nb
from oracle.dense_oracle import DensePosterior
from nxblink_b200 import toymodel, toydata
from test_gpu_toy_exact import _flatten, _flatten_exact, MODELS, DATA
import time
t0 = time.time()
for nseeds, burn, samp in ((32, 100, 40),):
    allz = []
    for M in MODELS:
        for D in DATA:
            exact, runs = None, []
            for seed in range(nseeds):
                s = nb.BlinkSampler(M, validate=False)
                s.set_data([D]); s.init_chains(chains=1024, seed=1000 + seed)
                s.sweep(burn, accumulate=False); s.sweep(samp, accumulate=True)
                sm = s.summary()
                if exact is None: exact = _flatten_exact(DensePosterior(M, D).expected_summary(), s.pm)
                runs.append(_flatten(sm, s.pm)); s.close()
            zs = []
            for key, ex in exact.items():
                v = np.array([r[key] for r in runs]); se = v.std(ddof=1) / np.sqrt(nseeds)
                zs.append((abs(v.mean() - ex) / (se + 1e-300) if se > 0 else (0.0 if abs(v.mean()-ex) < 1e-12 else 99.0), abs(v.mean()-ex), se, key))
            zs.sort(key=lambda x: -x[0])
            print(M.__name__, D.__name__, 'n', len(zs), 'max z %.2f' % zs[0][0], 'diff %.2e se %.2e' % (zs[0][1], zs[0][2]), zs[0][3], 'rms %.2f' % np.sqrt(np.mean([z[0]**2 for z in zs if z[0] < 90])))
            allz += [z[0] for z in zs]
    print('all', len(allz), 'max', max(allz), 'time', time.time() - t0)
