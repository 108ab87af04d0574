"""Exact posterior expectations of the blink-model summary statistics.

TEST INFRASTRUCTURE ONLY (see oracle/blink_oracle.py header for who may
import ``oracle/``).

Restates the reference's deterministic toy cross-check, generalised from its
hard-coded 48-state space to any small (P primary states, K classes) model and
to per-edge rate scaling:

    nxblink/compound.py:23-90        compound_state_is_ok, get_Q_compound
    nxblink/denseutil.py:39-176      compound states, indicator matrices,
                                     compute_edge_expectation / compute_dwell_times
    examples/codon2x3/toymodel-analytic.py:86-226
                                     expm per edge, pruning likelihood,
                                     posterior node / edge distributions,
                                     Frechet-derivative expectations

``scipy.linalg.expm`` / ``expm_frechet`` are the numerical authority exactly as
in the reference; the pruning recursion (``nxmctree.dynamic_fset_lhood`` in the
reference, an un-vendored dependency) is restated densely here.

Only FEASIBLE compound states are enumerated (primary state p with its own
class ON), i.e. P * 2**(K-1) states; the reference enumerates all P * 2**K and
skips the infeasible ones (compound.py:57-64), which is the same process.

What it returns are the per-sample EXPECTED values of every statistic that
``nxblink.summary.Summary`` accumulates, so a sampler's averages can be checked
against exact numbers instead of against another sampler.
"""
from __future__ import division, print_function

import itertools

import numpy as np
import scipy.linalg

from .blink_oracle import build_tables, PRIMARY


class DensePosterior(object):
    """Exact expectations under the compound process conditioned on the data."""

    def __init__(self, model, data):
        tb = build_tables(model)
        self.tb = tb
        self.states = sorted(tb.primary_to_tol, key=repr)
        self.tols = list(tb.tolerance_names)
        K = len(self.tols)
        tol_idx = dict((t, i) for i, t in enumerate(self.tols))
        self.cls = dict((s, tol_idx[tb.primary_to_tol[s]]) for s in self.states)

        # feasible compound states (compound.py:23-40)
        comp = []
        for p in self.states:
            for bits in itertools.product((0, 1), repeat=K):
                if bits[self.cls[p]]:
                    comp.append((p, bits))
        self.comp = comp
        idx = dict((c, i) for i, c in enumerate(comp))
        n = len(comp)

        # unit-rate compound generator (compound.py:43-90); the per-edge rate
        # scales primary AND blink rates (nxblink/poisson.py:85).
        Q = np.zeros((n, n))
        self.pri_ind = {}     # (sa, sb) -> indicator matrix
        I_on = np.zeros((n, n))
        I_off = np.zeros((n, n))
        for i, (p, bits) in enumerate(comp):
            for pb, rate in tb.Q_primary[p].items():
                if bits[self.cls[pb]]:
                    j = idx[(pb, bits)]
                    Q[i, j] = rate
                    self.pri_ind.setdefault((p, pb), np.zeros((n, n)))[i, j] = 1
            for k in range(K):
                flipped = bits[:k] + (1 - bits[k],) + bits[k + 1:]
                if bits[k] == 0:
                    j = idx[(p, flipped)]
                    Q[i, j] = tb.rate_on
                    I_on[i, j] = 1
                elif k != self.cls[p]:
                    j = idx[(p, flipped)]
                    Q[i, j] = tb.rate_off
                    I_off[i, j] = 1
        Q -= np.diag(Q.sum(axis=1))
        self.Q = Q
        self.I_on, self.I_off = I_on, I_off

        # stationary prior (toymodel-analytic.py:92-104)
        prior = np.zeros(n)
        for i, (p, bits) in enumerate(comp):
            w = tb.primary_distn[p]
            for k in range(K):
                if k != self.cls[p]:
                    w *= tb.blink_distn[bool(bits[k])]
            prior[i] = w
        np.testing.assert_allclose(prior.sum(), 1)
        np.testing.assert_allclose(prior.dot(Q), 0, atol=1e-10)
        self.prior = prior

        # data -> per-node 0/1 feasibility vectors
        track_data = data.get_data()
        self.node_ok = {}
        for v in tb.T:
            ok = np.zeros(n)
            for i, (p, bits) in enumerate(comp):
                good = p in track_data[PRIMARY][v]
                for k, t in enumerate(self.tols):
                    good = good and (bool(bits[k]) in track_data[t][v])
                ok[i] = 1.0 if good else 0.0
            self.node_ok[v] = ok

        self._solve()

    # -- pruning + posterior (toymodel-analytic.py:112-157) -----------------
    def _solve(self):
        tb = self.tb
        self.edge_P = {}
        for edge in tb.edges:
            t = tb.edge_to_blen[edge] * tb.edge_to_rate[edge]
            self.edge_P[edge] = scipy.linalg.expm(self.Q * t)
        down = {}    # node -> partial likelihood of the subtree given node state
        order = [tb.root] + [vb for va, vb in tb.edges]
        for v in reversed(order):
            L = self.node_ok[v].copy()
            for c in tb.children.get(v, []):
                L = L * self.edge_P[v, c].dot(down[c])
            down[v] = L
        self.likelihood = float(self.prior.dot(down[tb.root]))
        post = {tb.root: self.prior * down[tb.root] / self.likelihood}
        self.edge_J = {}
        for va, vb in tb.edges:
            # joint posterior of (x_va, x_vb):
            # post[va][a] * P[a,b] * down[vb][b] / (P down[vb])[a]
            P = self.edge_P[va, vb]
            msg = P.dot(down[vb])
            with np.errstate(divide='ignore', invalid='ignore'):
                w = np.where(msg > 0, post[va] / msg, 0.0)
            J = (w[:, None] * P) * down[vb][None, :]
            self.edge_J[va, vb] = J
            post[vb] = J.sum(axis=0)
        self.node_post = post

    # -- Frechet expectations (denseutil.py:148-176) ------------------------
    def _edge_expect(self, edge, E):
        tb = self.tb
        t = tb.edge_to_blen[edge] * tb.edge_to_rate[edge]
        if not t:
            return 0.0
        interact = scipy.linalg.expm_frechet(
                self.Q * t, E * t, compute_expm=False)
        J, P = self.edge_J[edge], self.edge_P[edge]
        mask = J > 0
        return float((J[mask] * interact[mask] / P[mask]).sum())

    def expected_summary(self):
        """Per-sample expectations, keyed like ``Summary``'s attributes."""
        tb = self.tb
        K = len(self.tols)
        out = {}
        root_post = self.node_post[tb.root]
        rp = dict((s, 0.0) for s in self.states)
        on_total = off = xon = 0.0
        for i, (p, bits) in enumerate(self.comp):
            rp[p] += root_post[i]
            on_total += root_post[i] * sum(bits)
            off += root_post[i] * (K - sum(bits))
            xon += root_post[i] * (sum(bits) - 1)
        out['root_pri_to_count'] = rp
        out['root_on_total'] = on_total
        out['root_off_count'] = off
        out['root_xon_intended'] = xon
        out['edge_to_pri_trans'] = {}
        out['edge_to_pri_dwell'] = {}
        out['edge_to_off_xon_trans'] = {}
        out['edge_to_xon_off_trans'] = {}
        out['edge_to_off_xon_dwell'] = {}
        out['edge_to_xon_off_dwell'] = {}
        n = len(self.comp)
        off_cnt = np.array([K - sum(bits) for p, bits in self.comp], float)
        xon_cnt = np.array([sum(bits) - 1 for p, bits in self.comp], float)
        for edge in tb.edges:
            rate = tb.edge_to_rate[edge]
            blen = tb.edge_to_blen[edge]
            # Frechet integrals are in units of the scaled time rate*blen;
            # counts are unit-free, dwell converts back to blen units.
            dwell_scale = (1.0 / rate) if rate else 0.0
            tr, dw = {}, {}
            for (sa, sb), ind in self.pri_ind.items():
                tr.setdefault(sa, {})[sb] = self._edge_expect(edge, self.Q * ind)
                # time in sa with class(sb) ON
                d = np.array([1.0 if (p == sa and bits[self.cls[sb]]) else 0.0
                    for p, bits in self.comp])
                val = self._edge_expect(edge, np.diag(d)) * dwell_scale
                if not rate:
                    # no change along the edge: dwell = blen * P(state at va)
                    val = blen * float(self.node_post[edge[0]].dot(d))
                dw.setdefault(sa, {})[sb] = val
            out['edge_to_pri_trans'][edge] = tr
            out['edge_to_pri_dwell'][edge] = dw
            out['edge_to_off_xon_trans'][edge] = self._edge_expect(
                    edge, self.Q * self.I_on)
            out['edge_to_xon_off_trans'][edge] = self._edge_expect(
                    edge, self.Q * self.I_off)
            if rate:
                out['edge_to_off_xon_dwell'][edge] = self._edge_expect(
                        edge, np.diag(off_cnt)) * dwell_scale
                out['edge_to_xon_off_dwell'][edge] = self._edge_expect(
                        edge, np.diag(xon_cnt)) * dwell_scale
            else:
                pa = self.node_post[edge[0]]
                out['edge_to_off_xon_dwell'][edge] = blen * float(pa.dot(off_cnt))
                out['edge_to_xon_off_dwell'][edge] = blen * float(pa.dot(xon_cnt))
        return out
