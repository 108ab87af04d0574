"""Ad-hoc GPU check used during development (not part of the test suite)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nxblink_b200 as nb
from nxblink_b200 import toymodel, toydata
from oracle.dense_oracle import DensePosterior


def compare(name, got, exp, tol, scale=1.0):
    err = abs(got - exp)
    flag = '' if err <= tol * max(scale, abs(exp)) + 1e-12 else '   <<<<<<'
    print('  %-28s got %.5f exact %.5f%s' % (name, got, exp, flag))


def run(M, D, chains=4096, burn=50, sweeps=50, msg_f64=False):
    print('==', M.__name__, D.__name__, 'f64' if msg_f64 else 'f32')
    s = nb.BlinkSampler(M, validate=True, msg_f64=msg_f64)
    s.set_data([D])
    t0 = time.time()
    s.init_chains(chains=chains, seed=1234)
    s.sweep(burn, accumulate=False)
    s.sweep(sweeps, accumulate=True)
    dt = time.time() - t0
    print('  wall %.3fs  kernel ms %s  stats %s' % (dt, s.last_kernel_ms(), s.stats()))
    sm = s.summary()
    n = sm.nsamples
    ex = DensePosterior(M, D).expected_summary()
    pm = s.pm
    for st in pm.states:
        compare('root pri %s' % st, sm.root_pri_to_count.get(st, 0) / n, ex['root_pri_to_count'][st], 0.02, 1.0)
    compare('root off', sm.root_off_count / n, ex['root_off_count'], 0.02, 1.0)
    compare('root xon intended', sm.root_xon_intended / n, ex['root_xon_intended'], 0.02, 1.0)
    for e in pm.edges:
        compare('%s off->on' % (e,), sm.edge_to_off_xon_trans[e] / n, ex['edge_to_off_xon_trans'][e], 0.03, 0.3)
        compare('%s on->off' % (e,), sm.edge_to_xon_off_trans[e] / n, ex['edge_to_xon_off_trans'][e], 0.03, 0.3)
        compare('%s offxon dwell' % (e,), sm.edge_to_off_xon_dwell[e] / n, ex['edge_to_off_xon_dwell'][e], 0.03, 0.3)
        compare('%s xonoff dwell' % (e,), sm.edge_to_xon_off_dwell[e] / n, ex['edge_to_xon_off_dwell'][e], 0.03, 0.3)
        tr = sum(w['weight'] for sa in sm.edge_to_pri_trans[e] for w in sm.edge_to_pri_trans[e][sa].values()) / n
        trx = sum(c for sa in ex['edge_to_pri_trans'][e] for c in ex['edge_to_pri_trans'][e][sa].values())
        compare('%s pri trans total' % (e,), tr, trx, 0.03, 0.3)
        dw = sum(w['weight'] for sa in sm.edge_to_pri_dwell[e] for w in sm.edge_to_pri_dwell[e][sa].values()) / n
        dwx = sum(c for sa in ex['edge_to_pri_dwell'][e] for c in ex['edge_to_pri_dwell'][e][sa].values())
        compare('%s pri dwell total' % (e,), dw, dwx, 0.03, 0.3)
    # fused statistics vs the independent kernel on the final state
    s.reset_summary(); s.sweep(1, accumulate=True)
    c1, d1 = s.read_summary()
    s.reset_summary(); s.summarize_current()
    c2, d2 = s.read_summary()
    print('  fused vs independent: counts equal %s, dwell max abs diff %.3e' % (
        bool((c1 == c2).all()), float(np.abs(d1 - d2).max())))
    s.close()


if __name__ == '__main__':
    import ctypes as C
    lib = nb._lib.load()
    ctr = (C.c_uint32 * 4)(0, 0, 0, 0); key = (C.c_uint32 * 2)(0, 0); out = (C.c_uint32 * 4)()
    lib.nxb_philox4x32_10(ctr, key, out)
    print('philox KAT', [hex(x) for x in out])
    for M in (toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC):
        for D in (toydata.DataA, toydata.DataB, toydata.DataC, toydata.DataD):
            run(M, D)
    run(toymodel.BlinkModelB, toydata.DataC, msg_f64=True)
