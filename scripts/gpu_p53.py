"""Ad-hoc p53-shaped run used during development."""
import sys, time, os, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nxblink_b200 as nb
from nxblink_b200 import p53, packing

ap = argparse.ArgumentParser()
ap.add_argument('--sites', type=int, default=2048)
ap.add_argument('--chains', type=int, default=16)
ap.add_argument('--burn', type=int, default=4)
ap.add_argument('--sweeps', type=int, default=4)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--validate', type=int, default=0)
ap.add_argument('--f64', type=int, default=0)
ap.add_argument('--cap', type=float, default=1.0)
ap.add_argument('--warps', type=int, default=0)
ap.add_argument('--lockstep', type=int, default=0)
ap.add_argument('--split', type=int, default=0)
a = ap.parse_args()
m = p53.p53_model()
pm = packing.PackedModel(m)
ls, tt, tf = p53.synthetic_alignment(pm, a.sites, seed=0, reference_leaf=pm.node_index[m.name_to_leaf['Has']])
d = packing.pack_alignment(pm, ls, tt, tf)
s = nb.BlinkSampler(pm, validate=bool(a.validate), msg_f64=bool(a.f64), cap_scale=a.cap, max_warps=a.warps, lockstep_warps=a.lockstep, split_phases=a.split)
s.set_data(d)
t0 = time.time(); s.init_chains(chains=a.chains, seed=7); print('init %.3fs kernel %s' % (time.time() - t0, s.last_kernel_ms()))
print(s.stats())
t0 = time.time(); s.sweep(a.burn, accumulate=False); print('burn %.3fs kernel %s' % (time.time() - t0, s.last_kernel_ms()))
s.profile(True)
for r in range(a.reps):
    t0 = time.time(); s.sweep(a.sweeps, accumulate=True); dt = time.time() - t0
    ms, nl = s.last_kernel_ms()
    print('   by kernel:', dict((k, round(v[0], 3)) for k, v in s.last_kernel_ms_by_kind().items() if v[1]))
    n = a.sites * a.chains * a.sweeps
    print('sweep x%d: wall %.3fs kernel %.2f ms (%d launches) -> %.3e unit-sweeps/s (kernel)' % (a.sweeps, dt, ms, nl, n / (ms * 1e-3)))
prof = s.profile(False); tot = sum(prof.values()) or 1
for k, v in prof.items(): print('  %-22s %6.2f%%' % (k, 100.0 * v / tot))
st = s.stats(); print(st)
R = (st['n_pri_events'] + st['n_blk_events']) / st['n_units']
print('mean retained events R = %.1f (pri %.2f blk %.2f); state bytes/unit %.0f' % (R, st['n_pri_events'] / st['n_units'], st['n_blk_events'] / st['n_units'], st['state_bytes_used'] / st['n_units']))
sm = s.summary()
print('nsamples', sm.nsamples, 'root off/sample', sm.root_off_count / sm.nsamples, 'xon', sm.root_xon_count / sm.nsamples)
tot_pri = sum(w['weight'] for e in pm.edges for sa in sm.edge_to_pri_trans[e] for w in sm.edge_to_pri_trans[e][sa].values()) / sm.nsamples
print('pri trans/sample %.3f  blink trans/sample %.3f' % (tot_pri, sum(sm.edge_to_off_xon_trans[e] + sm.edge_to_xon_off_trans[e] for e in pm.edges) / sm.nsamples))
