"""Driver of the CPU simulation of the tolerance update (oracle/hostsim, test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

from nxblink_b200 import _lib as nlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'oracle', '_ref', 'libtol_hostsim.so')


class HsUnit(C.Structure):
    _fields_ = [('n_pri', C.c_int32), ('n_blk', C.c_int32),
            ('node_pri', C.c_void_p), ('node_tol', C.c_void_p),
            ('ptm', C.c_void_p), ('pcode', C.c_void_p),
            ('btm', C.c_void_p), ('bcode', C.c_void_p), ('cap_blk', C.c_int32),
            ('tt', C.c_void_p), ('tf', C.c_void_p)]


SO_WALK = os.path.join(ROOT, 'oracle', '_ref', 'libtol_hostsim_walk.so')
_hs = {}


def load(walk=False):
    """walk=True: the comparison build whose downward pass is the full walk (TOL_FORCE_WALK)."""
    if walk not in _hs:
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle', 'hostsim'), 'all'])
        lib = C.CDLL(SO_WALK if walk else SO)
        lib.hostsim_tol_update.argtypes = [C.POINTER(nlib.ModelDesc), C.POINTER(HsUnit), C.c_uint64,
                C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _hs[walk] = lib
    return _hs[walk]


def forward_primary_path(pm, rng):
    """A primary trajectory drawn from the unconstrained primary process (all classes ON)."""
    N, P = pm.n_nodes, pm.n_states
    node_pri = np.zeros(N, dtype=np.uint8)
    node_pri[0] = rng.choice(P, p=pm.prior / pm.prior.sum())
    events = []
    tm = pm.node_tm
    for v in range(1, N):
        e = v - 1
        s = int(node_pri[pm.parent[v]])
        t, tend, r = tm[pm.parent[v]], tm[v], pm.rate[v]
        while r > 0 and tend > t:
            lo, hi = pm.q_ptr[s], pm.q_ptr[s + 1]
            tot = pm.q_val[lo:hi].sum() * r
            t += rng.exponential(1.0 / tot)
            if t >= tend:
                break
            sb = int(pm.q_col[lo + rng.choice(hi - lo, p=pm.q_val[lo:hi] / pm.q_val[lo:hi].sum())])
            events.append((e, float(t), s, sb))
            s = sb
        node_pri[v] = s
    events.sort()
    return node_pri, events


class HostTolChain(object):
    """Repeated tolerance updates of one unit with a FIXED primary trajectory."""

    def __init__(self, pm, node_pri, pri_events, tt, tf, msg_f64=False, seed=1, cap=4096, walk=False):
        self.lib = load(walk)
        self.pm = pm
        self.desc = pm.desc()
        N, E, P, K = pm.n_nodes, pm.n_nodes - 1, pm.n_states, pm.n_classes
        self.node_pri = np.ascontiguousarray(node_pri, dtype=np.uint8)
        allk = (1 << K) - 1
        # start: every class ON; where the data force a class OFF at a leaf, one on->off event on the
        # leaf edge after the last stretch in which the primary process is in that class (feasible)
        self.node_tol = np.full(N, allk, dtype=np.uint32)
        init = []
        kids = set(int(x) for x in pm.parent[1:])
        for v in range(1, N):
            if v in kids:
                continue
            for k in range(K):
                if (int(tt[v]) >> k) & 1:
                    continue
                e = v - 1
                t0 = pm.node_tm[pm.parent[v]]
                if pm.state_class[node_pri[pm.parent[v]]] == k:
                    t0 = None
                for ee, t, sa, sb in pri_events:
                    if ee == e:
                        if pm.state_class[sb] == k:
                            t0 = None
                        elif t0 is None:
                            t0 = t
                assert t0 is not None and pm.node_tm[v] > t0
                init.append((k, e, 0.5 * (t0 + pm.node_tm[v])))
                self.node_tol[v] &= ~np.uint32(1 << k)
        init.sort()
        self.ptm = np.array([t for _, t, _, _ in pri_events] + [0.0], dtype=np.float64)
        self.pcode = np.array([e | (sa << 16) | (sb << 24) for e, _, sa, sb in pri_events] + [0],
                dtype=np.uint32)
        self.btm = np.zeros(cap)
        self.bcode = np.zeros(cap, dtype=np.uint32)
        for i, (k, e, t) in enumerate(init):
            self.btm[i] = t
            self.bcode[i] = e | (k << 16)          # new state OFF
        self.tt = np.ascontiguousarray(tt, dtype=np.uint32)
        self.tf = np.ascontiguousarray(tf, dtype=np.uint32)
        self.un = HsUnit(len(pri_events), len(init), self.node_pri.ctypes.data, self.node_tol.ctypes.data,
                self.ptm.ctypes.data, self.pcode.ctypes.data, self.btm.ctypes.data,
                self.bcode.ctypes.data, cap, self.tt.ctypes.data, self.tf.ctypes.data)
        self.doff = np.zeros(E * P * K)
        self.off_on = np.zeros(E, dtype=np.uint64)
        self.on_off = np.zeros(E, dtype=np.uint64)
        self.msg_f64 = 1 if msg_f64 else 0
        self.seed = seed
        self.sweeps = 0
        self.n_live = C.c_int32()

    def sweep(self, accumulate=True):
        rc = self.lib.hostsim_tol_update(C.byref(self.desc), C.byref(self.un), self.seed, 0,
                self.sweeps, self.msg_f64, 1 if accumulate else 0, self.doff.ctypes.data,
                self.off_on.ctypes.data, self.on_off.ctypes.data, C.addressof(self.n_live))
        if rc:
            raise RuntimeError('hostsim_tol_update failed: %d' % rc)
        self.sweeps += 1

    def blink_events(self):
        n = self.un.n_blk
        code = self.bcode[:n]
        return self.btm[:n].copy(), (code & 0xffff).astype(int), ((code >> 16) & 0xff).astype(int), (code >> 31).astype(int)
