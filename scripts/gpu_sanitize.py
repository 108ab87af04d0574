"""Small end-to-end run for compute-sanitizer: split and fused forms, deferral tiers, forward
sampler, summaries, dense kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import p53, packing, toymodel, toydata, dense
m = p53.p53_model(); pm = packing.PackedModel(m)
leaf = pm.node_index[m.name_to_leaf['Has']]
ls, tt, tf = p53.synthetic_alignment(pm, 6, seed=0, reference_leaf=leaf)
data = packing.pack_alignment(pm, ls, tt, tf)
for kw in (dict(), dict(split_phases=-1), dict(cap_scale=0.55), dict(validate=True)):
    s = nb.BlinkSampler(pm, **kw); s.set_data(data); s.init_chains(chains=8, seed=1)
    s.sweep(2, accumulate=False); s.sweep(2, accumulate=True); s.summarize_current(); s.summary()
    s.export_tracks(3); print(kw, s.stats()['deferred_units']); s.close()
s = nb.BlinkSampler(toymodel.BlinkModelB); s.set_data([toydata.DataC] * 5); s.init_chains(chains=16, seed=2)
s.sweep(3, accumulate=True); s.close()
s = nb.BlinkSampler(pm); s.forward_sample(64, seed=3); s.sweep(2, accumulate=True); s.close()
cp = dense.CompoundProcess(toymodel.BlinkModelB)
cp.expected_summary(packing.pack_data(cp.pm, [toydata.DataC])); print('ok')
