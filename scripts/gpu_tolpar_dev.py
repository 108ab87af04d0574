"""Development check of the event-parallel tolerance update: toy model against the exact posterior
(validated trajectories), then the p53 shape with validation."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nxblink_b200 as nb
from nxblink_b200 import toymodel, toydata
from oracle.dense_oracle import DensePosterior

for M, D in ((toymodel.BlinkModelB, toydata.DataC), (toymodel.BlinkModelA, toydata.DataA), (toymodel.BlinkModelC, toydata.DataB)):
    for f64 in (False, True):
        s = nb.BlinkSampler(M, device=0, validate=True, msg_f64=f64)
        s.set_data([D])
        s.init_chains(chains=8192, seed=11)
        s.sweep(40, accumulate=False)
        s.sweep(40, accumulate=True)
        sm = s.summary()
        ex = DensePosterior(M, D).expected_summary()
        n = float(sm.nsamples)
        worst = 0.0
        for e in s.pm.edges:
            for got, exact in ((sm.edge_to_off_xon_trans[e], ex['edge_to_off_xon_trans'][e]),
                    (sm.edge_to_xon_off_trans[e], ex['edge_to_xon_off_trans'][e]),
                    (sm.edge_to_off_xon_dwell[e], ex['edge_to_off_xon_dwell'][e]),
                    (sm.edge_to_xon_off_dwell[e], ex['edge_to_xon_off_dwell'][e])):
                worst = max(worst, abs(got / n - exact))
        for st, pr in ex['root_pri_to_count'].items():
            worst = max(worst, abs(sm.root_pri_to_count.get(st, 0) / n - pr))
        worst = max(worst, abs(sm.root_off_count / n - ex['root_off_count']))
        print(M.__name__, D.__name__, 'f64' if f64 else 'f32', 'max |sampled - exact| = %.5f' % worst, s.stats()['deferred_units'])
        s.close()
