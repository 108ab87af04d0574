"""Long-run bias check: every Summary statistic of a toy grid against the exact posterior with a
pure 5-SE criterion (no absolute slack), ~10^8 samples.  Exact values come from the device dense
path (dense.CompoundProcess.expected_summary), itself checked against the oracle in the tests."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nxblink_b200 as nb
from nxblink_b200 import toymodel, toydata, dense, packing

def flat_exact(ex, pm):
    out = {}
    for s in pm.states: out['root_pri', s] = ex['root_pri_to_count'][s]
    out['root_off'] = ex['root_off_count']; out['root_xon'] = ex['root_xon_intended']
    for e in pm.edges:
        for key in ('off_xon_trans', 'xon_off_trans', 'off_xon_dwell', 'xon_off_dwell'):
            out[key, e] = ex['edge_to_' + key][e]
        for i in range(pm.nnz):
            sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
            out['pri_trans', e, sa, sb] = ex['edge_to_pri_trans'][e][sa][sb]
            out['pri_dwell', e, sa, sb] = ex['edge_to_pri_dwell'][e][sa][sb]
    return out

def flat(sm, pm):
    n = float(sm.nsamples); out = {}
    for s in pm.states: out['root_pri', s] = sm.root_pri_to_count.get(s, 0) / n
    out['root_off'] = sm.root_off_count / n; out['root_xon'] = sm.root_xon_intended / n
    for e in pm.edges:
        out['off_xon_trans', e] = sm.edge_to_off_xon_trans[e] / n
        out['xon_off_trans', e] = sm.edge_to_xon_off_trans[e] / n
        out['off_xon_dwell', e] = sm.edge_to_off_xon_dwell[e] / n
        out['xon_off_dwell', e] = sm.edge_to_xon_off_dwell[e] / n
        for i in range(pm.nnz):
            sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
            out['pri_trans', e, sa, sb] = sm.flat['pri_trans'][pm.edge_index[e], i] / n
            out['pri_dwell', e, sa, sb] = sm.flat['pri_dwell'][pm.edge_index[e], i] / n
    return out

def main(nseeds=16, chains=65536, sweeps=100, grids=None):
  for M, D in ((toymodel.BlinkModelB, toydata.DataC), (toymodel.BlinkModelA, toydata.DataA),
          (toymodel.BlinkModelC, toydata.DataB)):
      cp = dense.CompoundProcess(M)
      exact = flat_exact(cp.expected_summary(packing.pack_data(cp.pm, [D])), cp.pm)
      runs = []
      for seed in range(nseeds):
          s = nb.BlinkSampler(M)
          s.set_data([D]); s.init_chains(chains=chains, seed=7000 + seed)
          s.sweep(50, accumulate=False); s.sweep(sweeps, accumulate=True)
          runs.append(flat(s.summary(), s.pm)); s.close()
      zs = []
      for key, ex in exact.items():
          v = np.array([r[key] for r in runs]); se = v.std(ddof=1) / np.sqrt(nseeds)
          if se > 0: zs.append((abs(v.mean() - ex) / se, key, v.mean(), ex, se))
          elif abs(v.mean() - ex) > 1e-12: zs.append((np.inf, key, v.mean(), ex, se))
      zs.sort(key=lambda x: -x[0])
      z = np.array([x[0] for x in zs])
      print(M.__name__, D.__name__, 'samples %.1e' % (nseeds * chains * sweeps), 'statistics', len(zs),
              'max z %.2f' % z.max(), 'frac z>2: %.3f' % (z > 2).mean(), 'rms z %.2f' % np.sqrt((z ** 2).mean()))
      print('   worst:', zs[0][1:], 'rel err %.2e' % (abs(zs[0][2] - zs[0][3]) / max(abs(zs[0][3]), 1e-300)))


if __name__ == '__main__':
    if len(sys.argv) > 1:      # e.g. 64: more seeds
        main(nseeds=int(sys.argv[1]))
    else:
        main()
