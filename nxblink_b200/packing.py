"""Host-side packing of nxblink's duck-typed model / data objects into the
integer-indexed tables the C-ABI takes.

Mirrors the model extraction of nxblink/raoteh.py:59-119 (9 model getters,
normalisation of Q_primary to expected rate 1) and the data handling of
raoteh.py:61,132-134,160-184 (per-track node -> feasible-set maps, pooling over
zero-length / zero-rate branches).  Unlike raoteh.py:101-102 the caller's
``Q_primary`` graph is NOT mutated.
"""
from __future__ import division, print_function

import ctypes as C

import numpy as np

from . import _lib

PRIMARY = 'PRIMARY'


def _weight(w):
    return w['weight'] if isinstance(w, dict) else w


def _children(T, v):
    try:
        return list(T[v])
    except KeyError:
        return []


class PackedModel(object):
    """Integer-indexed model tables plus the maps back to the caller's labels.

    Attributes
    ----------
    nodes : list            node label of internal node index v (root first)
    edges : list            edge label (va, vb) of internal edge index e (edge into node e+1)
    states : list           primary state label of internal state index
    tols : list             tolerance-class (track) name of internal class index
    """

    def __init__(self, model):
        T, root = model.get_T_and_root()
        edge_to_blen = model.get_edge_to_blen()
        edge_to_rate = model.get_edge_to_rate()
        primary_to_tol = dict(model.get_primary_to_tol())
        Q_nx = model.get_Q_primary()
        distn = dict(model.get_primary_distn())

        # ---- tree: depth-first numbering, heaviest subtree numbered last -------
        # (the upward blink pass walks edges in descending index, so the heaviest
        # child is folded first and few accumulators are live at once)
        size = {}
        order = [root]
        for v in order:
            order.extend(_children(T, v))
        for v in reversed(order):
            size[v] = 1 + sum(size[c] for c in _children(T, v))
        nodes, parent = [], []
        stack = [(root, -1)]
        while stack:
            v, par = stack.pop()
            idx = len(nodes)
            nodes.append(v)
            parent.append(par)
            kids = sorted(_children(T, v), key=lambda c: size[c])   # light first
            for c in reversed(kids):          # stack: light child popped first
                stack.append((c, idx))
        self.root = root
        self.T = T
        self.nodes = nodes
        self.node_index = dict((v, i) for i, v in enumerate(nodes))
        self.parent = np.array(parent, dtype=np.int32)
        N = len(nodes)
        self.n_nodes = N
        self.edges = [(nodes[parent[v]], nodes[v]) for v in range(1, N)]
        self.edge_index = dict((e, i) for i, e in enumerate(self.edges))
        self.blen = np.zeros(N)
        self.rate = np.zeros(N)
        for i, e in enumerate(self.edges):
            self.blen[i + 1] = edge_to_blen[e]
            self.rate[i + 1] = edge_to_rate[e]
        self.node_tm = np.zeros(N)
        for v in range(1, N):
            self.node_tm[v] = self.node_tm[parent[v]] + self.blen[v]
        self.node_to_tm = dict((nodes[v], float(self.node_tm[v])) for v in range(N))

        # ---- primary process ----------------------------------------------------
        try:
            self.states = sorted(primary_to_tol)
        except TypeError:
            self.states = sorted(primary_to_tol, key=repr)
        self.state_index = dict((s, i) for i, s in enumerate(self.states))
        names = set(primary_to_tol.values())
        try:
            self.tols = sorted(names)
        except TypeError:
            self.tols = sorted(names, key=repr)
        self.tol_index = dict((t, i) for i, t in enumerate(self.tols))
        self.primary_to_tol = primary_to_tol
        P, K = len(self.states), len(self.tols)
        self.n_states, self.n_classes = P, K
        self.state_class = np.array(
                [self.tol_index[primary_to_tol[s]] for s in self.states], dtype=np.int32)
        self.prior = np.array([distn.get(s, 0.0) for s in self.states], dtype=float)
        rows = [[] for _ in range(P)]
        for sa in Q_nx:
            for sb in Q_nx[sa]:
                rows[self.state_index[sa]].append(
                        (self.state_index[sb], float(_weight(Q_nx[sa][sb]))))
        # raoteh.py:90-102: expected rate sum_a pi_a sum_b q_ab := 1
        self.expected_rate = sum(
                self.prior[a] * q for a in range(P) for b, q in rows[a])
        q_ptr, q_col, q_val = [0], [], []
        for a in range(P):
            for b, q in sorted(rows[a]):
                q_col.append(b)
                q_val.append(q / self.expected_rate)
            q_ptr.append(len(q_col))
        self.q_ptr = np.array(q_ptr, dtype=np.int32)
        self.q_col = np.array(q_col, dtype=np.int32)
        self.q_val = np.array(q_val, dtype=float)
        self.nnz = len(q_col)
        self.q_row = np.repeat(np.arange(P, dtype=np.int32), np.diff(self.q_ptr))
        self.rate_on = float(model.get_rate_on())
        self.rate_off = float(model.get_rate_off())
        self.diameter = self._diameter()

    def _diameter(self):
        """nx.diameter(Q_primary) of raoteh.py:351 (directed, unweighted)."""
        P = self.n_states
        adj = [self.q_col[self.q_ptr[a]:self.q_ptr[a + 1]] for a in range(P)]
        diam = 0
        for src in range(P):
            dist = -np.ones(P, dtype=int)
            dist[src] = 0
            frontier = [src]
            while frontier:
                nxt = []
                for a in frontier:
                    for b in adj[a]:
                        if dist[b] < 0:
                            dist[b] = dist[a] + 1
                            nxt.append(b)
                frontier = nxt
            if (dist < 0).any():
                raise Exception('Q_primary is not strongly connected')
            diam = max(diam, int(dist.max()))
        return diam

    # -- zero-length pooling, raoteh.py:160-184 -------------------------------
    def zero_pools(self):
        """Pool id per node: nodes joined by zero-length or zero-rate edges."""
        pool = np.arange(self.n_nodes)
        for v in range(1, self.n_nodes):
            if not self.blen[v] or not self.rate[v]:
                pool[v] = pool[self.parent[v]]      # parent[v] < v: already final
        return pool

    def desc(self, msg_f64=False, cap_scale=1.0, validate=False, max_warps=0, lockstep_warps=0,
            split_phases=0):
        d = _lib.ModelDesc()
        d.n_nodes, d.n_states, d.n_classes, d.nnz = (
                self.n_nodes, self.n_states, self.n_classes, self.nnz)
        d.parent = self.parent.ctypes.data_as(C.POINTER(C.c_int32))
        d.blen = self.blen.ctypes.data_as(C.POINTER(C.c_double))
        d.rate = self.rate.ctypes.data_as(C.POINTER(C.c_double))
        d.q_ptr = self.q_ptr.ctypes.data_as(C.POINTER(C.c_int32))
        d.q_col = self.q_col.ctypes.data_as(C.POINTER(C.c_int32))
        d.q_val = self.q_val.ctypes.data_as(C.POINTER(C.c_double))
        d.state_class = self.state_class.ctypes.data_as(C.POINTER(C.c_int32))
        d.prior = self.prior.ctypes.data_as(C.POINTER(C.c_double))
        d.rate_on, d.rate_off = self.rate_on, self.rate_off
        d.diameter = self.diameter
        d.msg_f64 = 1 if msg_f64 else 0
        d.cap_scale = float(cap_scale)
        d.validate = 1 if validate else 0
        d.max_warps = int(max_warps)
        d.lockstep_warps = int(lockstep_warps)
        d.split_phases = int(split_phases)
        return d


class PackedData(object):
    """Per-site feasibility masks: pri_ok u64[S, N], tol_true_ok / tol_false_ok u32[S, N]."""

    def __init__(self, pri_ok, tol_true_ok, tol_false_ok):
        self.pri_ok = np.ascontiguousarray(pri_ok, dtype=np.uint64)
        self.tol_true_ok = np.ascontiguousarray(tol_true_ok, dtype=np.uint32)
        self.tol_false_ok = np.ascontiguousarray(tol_false_ok, dtype=np.uint32)
        self.n_sites = self.pri_ok.shape[0]


def pool_masks(pm, pri_ok, tt, tf):
    """raoteh.py:160-184 on bit masks: intersect within zero-length pools."""
    pool = pm.zero_pools()
    if (pool == np.arange(pm.n_nodes)).all():
        return pri_ok, tt, tf
    pri_ok, tt, tf = pri_ok.copy(), tt.copy(), tf.copy()
    for root in np.unique(pool):
        members = np.nonzero(pool == root)[0]
        if len(members) < 2:
            continue
        p = np.bitwise_and.reduce(pri_ok[:, members], axis=1)
        a = np.bitwise_and.reduce(tt[:, members], axis=1)
        b = np.bitwise_and.reduce(tf[:, members], axis=1)
        allk = np.uint32((1 << pm.n_classes) - 1)
        if (p == 0).any() or (((a | b) & allk) != allk).any():
            raise Exception('no data for the node pool consisting of '
                    + str([pm.nodes[m] for m in members]))
        pri_ok[:, members] = p[:, None]
        tt[:, members] = a[:, None]
        tf[:, members] = b[:, None]
    return pri_ok, tt, tf


def pack_data(pm, data_objects):
    """Pack a sequence of reference-style ``data`` objects (one per site).

    Each object exposes ``get_data()`` -> {track name: {node: set of feasible
    states}} (raoteh.py:45-52,61).
    """
    S, N = len(data_objects), pm.n_nodes
    pri_ok = np.zeros((S, N), dtype=np.uint64)
    tt = np.zeros((S, N), dtype=np.uint32)
    tf = np.zeros((S, N), dtype=np.uint32)
    for s, data in enumerate(data_objects):
        track_to_data = data.get_data()
        for v, node in enumerate(pm.nodes):
            m = 0
            for st in track_to_data[PRIMARY][node]:
                m |= 1 << pm.state_index[st]
            pri_ok[s, v] = m
            a = b = 0
            for k, name in enumerate(pm.tols):
                fset = track_to_data[name][node]
                if True in fset:
                    a |= 1 << k
                if False in fset:
                    b |= 1 << k
            tt[s, v], tf[s, v] = a, b
    return PackedData(*pool_masks(pm, pri_ok, tt, tf))


def pack_alignment(pm, leaf_states, tol_true_ok=None, tol_false_ok=None):
    """Vectorised packing for alignment-shaped data (examples/p53/p53data.py:52-94).

    leaf_states : int array [S, N], internal state index observed at node v or -1.
        An observed state also forces its own tolerance class ON at that node
        (p53data.py:76-81).
    tol_true_ok / tol_false_ok : optional u32 [S, N] extra constraints that are
        ANDed in (e.g. disease data at the reference leaf, p53data.py:83-92).
    """
    leaf_states = np.asarray(leaf_states)
    S, N = leaf_states.shape
    P, K = pm.n_states, pm.n_classes
    allp = np.uint64((1 << P) - 1) if P < 64 else np.uint64(0xffffffffffffffff)
    allk = np.uint32((1 << K) - 1) if K < 32 else np.uint32(0xffffffff)
    obs = leaf_states >= 0
    st = np.where(obs, leaf_states, 0).astype(np.uint64)
    pri_ok = np.where(obs, np.uint64(1) << st, allp)
    cls = pm.state_class[np.where(obs, leaf_states, 0)].astype(np.uint32)
    own = np.where(obs, np.uint32(1) << cls, np.uint32(0)).astype(np.uint32)
    tt = np.full((S, N), allk, dtype=np.uint32)
    tf = (allk & ~own).astype(np.uint32)
    if tol_true_ok is not None:
        tt &= np.asarray(tol_true_ok, dtype=np.uint32)
    if tol_false_ok is not None:
        tf &= np.asarray(tol_false_ok, dtype=np.uint32)
    return PackedData(*pool_masks(pm, pri_ok, tt, tf))
