"""Config 2 of BASELINE.json: codon2x3 toy model B, 10k sites x 64 chains on one B200."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import toymodel, packing
M = toymodel.BlinkModelB
pm = packing.PackedModel(M)
s = nb.BlinkSampler(pm)
node_pri, node_tol = s.forward_sample(10000, seed=0)     # synthetic data from the model itself
s.close()
is_leaf = np.ones(pm.n_nodes, dtype=bool); is_leaf[pm.parent[1:]] = False
ls = np.where(is_leaf[None, :], node_pri, -1)
allk = np.uint32(7)
tt = np.full(ls.shape, allk, np.uint32); tf = np.full(ls.shape, allk, np.uint32)
ref = pm.node_index['N0']            # full tolerance observation at the reference leaf (DataC style)
tt[:, ref] = node_tol[:, ref]; tf[:, ref] = (~node_tol[:, ref]) & allk
data = packing.pack_alignment(pm, ls, tt, tf)
s = nb.BlinkSampler(pm, lockstep_warps=int(os.environ.get('TOY_LOCKSTEP', '0')), split_phases=int(os.environ.get('TOY_SPLIT', '0')))
s.set_data(data); s.init_chains(chains=64, seed=1)
s.sweep(10, accumulate=False)
s.profile(True)
for r in range(3):
    s.sweep(10, accumulate=True); ms, n = s.last_kernel_ms()
    print('toy B 640k units x10 sweeps: %.2f ms -> %.3e site-sweeps/s' % (ms, 640000 * 10 / (ms * 1e-3)))
prof = s.profile(False); tot = sum(prof.values()) or 1
print({k: round(100.0 * v / tot, 1) for k, v in prof.items()})
print(s.stats())
