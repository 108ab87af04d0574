"""Validated soak on shapes away from the fixtures: a 150-node irregular tree (30 states, 8
classes) and the limits K = 32 / P = 64; forward-simulated starts, invariants after every phase."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import nxblink_b200 as nb
from nxblink_b200 import packing
from test_gpu_generality import BigModel
for label, M, n, sweeps in (('150 nodes, P=30, K=8', BigModel(150, 30, 8, seed=1), 32768, 60),
        ('40 nodes, P=64, K=32', BigModel(40, 64, 32, seed=2, deg=4), 16384, 60),
        ('12 nodes, P=5, K=1', BigModel(12, 5, 1, seed=3, deg=2), 65536, 200),
        ('300 nodes, P=12, K=4', BigModel(300, 12, 4, seed=4), 16384, 40)):
    pm = packing.PackedModel(M)
    for kw in (dict(), dict(split_phases=-1)):
        t0 = time.time()
        s = nb.BlinkSampler(pm, validate=True, **kw)
        s.forward_sample(n, seed=11)
        s.sweep(sweeps, accumulate=True)
        ms = s.last_kernel_ms()[0]
        st = s.stats(); s.close()
        print('%-24s %-22s %9d unit-sweeps ok (%.1f s, kernels %.0f ms = %.2e unit-sweeps/s, deferred %d, warps %d)' % (label, kw, n * sweeps,
            time.time() - t0, ms, n * sweeps / (ms * 1e-3), st['deferred_units'], st['warps_per_cta']))
