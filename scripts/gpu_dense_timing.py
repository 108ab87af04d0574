"""Config 5 pieces on the dense path: batched expm (48 edges, 61x61), its Frechet derivative, and
the pruning log-likelihood over 15 625 p53-shaped sites (host-buffer C-ABI calls, wall time)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nxblink_b200 import p53, packing, dense
m = p53.p53_model(); pm = packing.PackedModel(m)
ls, tt, tf = p53.synthetic_alignment(pm, 15625, seed=0)
data = packing.pack_alignment(pm, ls, tt, tf)
P = pm.n_states
Q = np.zeros((P, P)); Q[pm.q_row, pm.q_col] = pm.q_val; Q -= np.diag(Q.sum(axis=1))
t = pm.blen[1:] * pm.rate[1:]
def timeit(label, f, flops=None, reps=5):
    f(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    extra = '' if flops is None else '  %.2f TFLOP/s (FP64 tensor pipe, wall incl. copies)' % (flops / best / 1e12)
    print('%-44s %8.2f ms%s' % (label, 1e3 * best, extra)); return r
Pe = timeit('expm_batched: 48 x (61x61)', lambda: dense.expm_batched(Q, t))
E = np.random.default_rng(0).standard_normal((len(t), P, P))
timeit('expm_frechet_batched: 48 x (61x61)', lambda: dense.expm_frechet_batched(Q, t, E))
S, N = data.pri_ok.shape
flops = 2.0 * 64 * 64 * 32 * (N - 1) * ((S + 31) // 32)
ll = timeit('pruning_loglik: %d sites x %d nodes' % (S, N), lambda: dense.pruning_loglik(pm.parent, Pe, pm.prior, data.pri_ok), flops)
print('log likelihood of the alignment under the primary process:', float(ll.sum()))
many = np.tile(t, 64)
timeit('expm_batched: 3072 x (61x61)', lambda: dense.expm_batched(Q, many), flops=3072 * 2.0 * 64 ** 3 * (13 + 6))
