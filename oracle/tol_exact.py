"""TEST INFRASTRUCTURE ONLY: the exact conditional posterior of ONE tolerance track given a fixed
primary trajectory -- the target distribution of the tolerance update of
nxblink/raoteh.py:413-432 (chunking.py:39-155 + poisson.py:176-199).

Given the primary trajectory the K tolerance tracks are independent two-state (OFF, ON) processes
on the tree.  On a stretch of edge e where the primary state is s, track k evolves with generator

    rate_e * [[-on,  on          ],
              [ off, -off - qm_sk]]        (qm_sk = sum of primary rates from s into class k:
                                            the ON state pays them, chunking.py:76-79)

unless s belongs to class k, where OFF is impossible and ON cannot be left (chunking.py:69-73,
poisson.py:80-86: Q_local keeps only transitions into allowed states): weight diag(0, exp(-rate_e qm_sk dt)).
Message passing over the stretches gives node marginals, expected transition counts and OFF dwell
times in closed form up to a Simpson rule inside each stretch.
"""
from __future__ import division, print_function

import numpy as np
from scipy.linalg import expm


def _stretch_ops(G, own, q_pen, dt, nsub):
    """Transfer operators of the nsub equal pieces of one stretch (top -> bottom)."""
    h = dt / nsub
    if own:
        M = np.diag([0.0, np.exp(-q_pen * h)])
    else:
        M = expm(G * h)
    return M, h


def track_posterior(parent, blen, rate, node_pri, pri_events, k, state_class, qm, on, off,
        tt, tf, nsub=32):
    """Exact statistics of track k.

    parent/blen/rate : [N] arrays (edge e = edge into node e+1)
    node_pri         : [N] primary state at every node
    pri_events       : list of (edge, time, sa, sb), times absolute (root at 0)
    qm               : [P, K] dense Q_meta
    tt, tf           : [N] bit masks, bit k set: track may be ON / OFF at the node
    Returns dict(node_on [N], off_on [E], on_off [E], off_dwell {(e, s): t}).
    """
    N = len(parent)
    E = N - 1
    tm = np.zeros(N)
    for v in range(1, N):
        tm[v] = tm[parent[v]] + blen[v]
    children = [[] for _ in range(N)]
    for v in range(1, N):
        children[parent[v]].append(v)
    ev_by_edge = [[] for _ in range(E)]
    for e, t, sa, sb in pri_events:
        ev_by_edge[e].append((t, sa, sb))
    # stretches of every edge, top -> bottom: (s, t0, t1)
    stretches = []
    for e in range(E):
        va = parent[e + 1]
        s, t = node_pri[va], tm[va]
        out = []
        for te, sa, sb in sorted(ev_by_edge[e]):
            assert sa == s
            out.append((s, t, te))
            s, t = sb, te
        assert s == node_pri[e + 1]
        out.append((s, t, tm[e + 1]))
        stretches.append(out)

    def data(v):
        d = np.array([float((int(tf[v]) >> k) & 1), float((int(tt[v]) >> k) & 1)])
        if state_class[node_pri[v]] == k:
            d[0] = 0.0
        return d

    def ops(e):
        res = []
        r = rate[e + 1]
        for s, t0, t1 in stretches[e]:
            own = state_class[s] == k
            q = qm[s][k] * r
            G = r * np.array([[-on, on], [off, -off]]) - np.diag([0.0, q])
            n = nsub if (t1 > t0 and r > 0) else 1
            M, h = _stretch_ops(G, own, q, t1 - t0, n)
            res.append((s, own, t0, M, h, n))
        return res

    edge_ops = [ops(e) for e in range(E)]
    # upward
    beta_node = [None] * N
    beta_top = [None] * E           # message of edge e at its top
    beta_pts = [None] * E           # beta at every grid point of edge e, top -> bottom
    for v in range(N - 1, -1, -1):
        b = data(v)
        for c in children[v]:
            b = b * beta_top[c - 1]
        beta_node[v] = b
        if v > 0:
            e = v - 1
            pts = [b]
            cur = b
            for s, own, t0, M, h, n in reversed(edge_ops[e]):
                for _ in range(n):
                    cur = M.dot(cur)
                    pts.append(cur)
            beta_pts[e] = pts[::-1]
            beta_top[e] = cur
    prior = np.array([off / (on + off), on / (on + off)])
    Z = float(np.dot(prior, beta_node[0]))
    assert Z > 0
    # downward
    alpha_in = [None] * N           # message arriving at node v from above (without data_v and subtree)
    alpha_in[0] = prior
    node_on = np.zeros(N)
    off_on = np.zeros(E)
    on_off = np.zeros(E)
    off_dwell = {}
    for v in range(N):
        m = alpha_in[v] * beta_node[v]
        node_on[v] = m[1] / m.sum()
        for c in children[v]:
            e = c - 1
            a = alpha_in[v] * data(v)
            for c2 in children[v]:
                if c2 != c:
                    a = a * beta_top[c2 - 1]
            r = rate[e + 1]
            idx = 0
            for s, own, t0, M, h, n in edge_ops[e]:
                # Simpson over the n pieces of this stretch (n even or 1)
                f_off, f_up, f_dn = [], [], []
                cur = a
                for j in range(n + 1):
                    bt = beta_pts[e][idx + j]
                    f_off.append(cur[0] * bt[0])
                    f_up.append(0.0 if own else cur[0] * on * r * bt[1])
                    f_dn.append(0.0 if own else cur[1] * off * r * bt[0])
                    if j < n:
                        cur = cur.dot(M)
                if n > 1:
                    w = np.ones(n + 1)
                    w[1:-1:2] = 4.0
                    w[2:-1:2] = 2.0
                    w *= h / 3.0
                    off_t = float(np.dot(w, f_off))
                    off_on[e] += float(np.dot(w, f_up)) / Z
                    on_off[e] += float(np.dot(w, f_dn)) / Z
                else:
                    # no evolution inside the stretch (zero rate or zero length): the time still counts
                    off_t = f_off[0] * (h * n)
                off_dwell[(e, s)] = off_dwell.get((e, s), 0.0) + off_t / Z
                a = cur
                idx += n
            alpha_in[c] = a
    return dict(node_on=node_on, off_on=off_on, on_off=on_off, off_dwell=off_dwell, Z=Z)
