"""Validated soak at the p53 shape: every trajectory invariant checked after every phase
(validate=1) over ~10^8 unit-sweeps, split and fused forms, with capacity pressure."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nxblink_b200 as nb
from nxblink_b200 import p53, packing
m = p53.p53_model(); pm = packing.PackedModel(m)
leaf = pm.node_index[m.name_to_leaf['Has']]
ls, tt, tf = p53.synthetic_alignment(pm, 1024, seed=5, reference_leaf=leaf)
data = packing.pack_alignment(pm, ls, tt, tf)
total = 0
for label, kw, chains, sweeps in [('split', dict(), 64, 300)] * int(os.environ.get('SOAK_REPS', '2')) + [
        ('fused', dict(split_phases=-1), 32, 150), ('split, tight capacities', dict(cap_scale=0.6), 32, 100),
        ('f64 messages', dict(msg_f64=True), 32, 100)]:
    t0 = time.time()
    s = nb.BlinkSampler(pm, validate=True, **kw)
    s.set_data(data); s.init_chains(chains=chains, seed=int(t0) % 1000)
    s.sweep(sweeps, accumulate=True)
    st = s.stats(); s.close()
    total += 1024 * chains * sweeps
    print('%-26s %9d unit-sweeps ok  (%.1f s, deferred %d)' % (label, 1024 * chains * sweeps, time.time() - t0, st['deferred_units']))
print('total', total)
