"""Deterministic dense path on the device: batched matrix exponentials, the pruning
likelihood, and the explicit compound process for small models.

Mirrors the reference's test-only dense machinery
    nxblink/compound.py:23-90       compound_state_is_ok, get_Q_compound
    nxblink/denseutil.py:39-146     compound states, nx <-> dense matrices
    nxblink/denseutil.py:148-176    Frechet-derivative expectations on an edge
    examples/codon2x3/toymodel-analytic.py:86-226   Q_compound, expm per edge, likelihood,
                                    posterior node / edge distributions, expected statistics
with the numerical kernels (``expm`` and its Frechet derivative per edge, pruning recursion
over sites) running as FP64 tensor-pipe CUDA kernels behind ``nxb_expm_batched`` /
``nxb_expm_frechet_batched`` / ``nxb_pruning_loglik``.
"""
from __future__ import division, print_function

import itertools

import numpy as np

from . import _lib
from ._lib import NxbError
from .packing import PackedModel


def expm_batched(Q, t, device=0):
    """expm(Q * t_b) for every t_b; Q is a dense n x n rate matrix (n <= 64)."""
    lib = _lib.load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    t = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
    n = Q.shape[0]
    out = np.zeros((len(t), n, n))
    rc = lib.nxb_expm_batched(int(device), n, len(t), Q.ctypes.data, t.ctypes.data, out.ctypes.data)
    if rc:
        raise NxbError(rc, lib.nxb_dense_last_error().decode())
    return out


def expm_frechet_batched(Q, t, E, device=0, compute_expm=True):
    """(expm(Q t_b), L(Q t_b, E_b)) with L the Frechet derivative of the matrix exponential,
    L(A, E) = int_0^1 exp(A s) E exp(A (1 - s)) ds -- ``scipy.linalg.expm_frechet(Q t_b, E_b)``
    for every b, on the device."""
    lib = _lib.load()
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    t = np.ascontiguousarray(np.atleast_1d(t), dtype=np.float64)
    n = Q.shape[0]
    E = np.ascontiguousarray(E, dtype=np.float64).reshape(len(t), n, n)
    L = np.zeros((len(t), n, n))
    P = np.zeros((len(t), n, n)) if compute_expm else None
    rc = lib.nxb_expm_frechet_batched(int(device), n, len(t), Q.ctypes.data, t.ctypes.data,
            E.ctypes.data, P.ctypes.data if compute_expm else None, L.ctypes.data)
    if rc:
        raise NxbError(rc, lib.nxb_dense_last_error().decode())
    return (P, L) if compute_expm else L


def pruning_loglik(parent, P_edges, prior, masks, device=0):
    """Per-site log likelihood on a tree (parent[v] < v, root 0).

    P_edges[e] is the transition matrix of the edge into node e+1; masks[s, v] is the
    bit mask of states feasible at node v of site s."""
    lib = _lib.load()
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    P_edges = np.ascontiguousarray(P_edges, dtype=np.float64)
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    n = P_edges.shape[1]
    S, N = masks.shape
    out = np.zeros(S)
    rc = lib.nxb_pruning_loglik(int(device), n, N, parent.ctypes.data, P_edges.ctypes.data,
            prior.ctypes.data, S, masks.ctypes.data, out.ctypes.data)
    if rc:
        raise NxbError(rc, lib.nxb_dense_last_error().decode())
    return out


def primary_loglik(pm, packed_data, device=0):
    """Log likelihood of the PRIMARY process alone (every tolerance class on) per site:
    one expm per edge + pruning over all sites."""
    P, N = pm.n_states, pm.n_nodes
    Q = np.zeros((P, P))
    Q[pm.q_row, pm.q_col] = pm.q_val
    Q -= np.diag(Q.sum(axis=1))
    t = pm.blen[1:] * pm.rate[1:]
    Pe = expm_batched(Q, t, device=device)
    return pruning_loglik(pm.parent, Pe, pm.prior, packed_data.pri_ok, device=device)


class CompoundProcess(object):
    """Explicit compound (primary x tolerance) process of a small blink model.

    Feasible compound states only (the primary state's own class is on): P * 2**(K-1)
    states, at most 64.  ``Q`` is the unit-rate generator; every edge scales it by its
    rate (nxblink/poisson.py:85 scales primary and blink rates alike).
    """

    def __init__(self, model):
        pm = model if isinstance(model, PackedModel) else PackedModel(model)
        self.pm = pm
        P, K = pm.n_states, pm.n_classes
        comp = [(p, bits) for p in range(P) for bits in itertools.product((0, 1), repeat=K)
                if bits[pm.state_class[p]]]
        if len(comp) > 64:
            raise Exception('compound state space too large for the dense path (%d > 64)' % len(comp))
        self.states = comp
        idx = dict((c, i) for i, c in enumerate(comp))
        n = len(comp)
        Q = np.zeros((n, n))
        for i, (p, bits) in enumerate(comp):
            for j in range(pm.q_ptr[p], pm.q_ptr[p + 1]):
                pb = pm.q_col[j]
                if bits[pm.state_class[pb]]:
                    Q[i, idx[(pb, bits)]] = pm.q_val[j]
            for k in range(K):
                flipped = bits[:k] + (1 - bits[k],) + bits[k + 1:]
                if bits[k] == 0:
                    Q[i, idx[(p, flipped)]] = pm.rate_on
                elif k != pm.state_class[p]:
                    Q[i, idx[(p, flipped)]] = pm.rate_off
        self.Q = Q - np.diag(Q.sum(axis=1))
        p_on = pm.rate_on / (pm.rate_on + pm.rate_off)
        prior = np.zeros(n)
        for i, (p, bits) in enumerate(comp):
            w = pm.prior[p]
            for k in range(K):
                if k != pm.state_class[p]:
                    w *= p_on if bits[k] else 1.0 - p_on
            prior[i] = w
        self.prior = prior

    def masks(self, packed_data):
        """Compound feasibility masks u64[S, N] from per-track masks."""
        pm = self.pm
        S, N = packed_data.pri_ok.shape
        out = np.zeros((S, N), dtype=np.uint64)
        for i, (p, bits) in enumerate(self.states):
            ok = (packed_data.pri_ok >> np.uint64(p)) & np.uint64(1)
            ok = ok.astype(bool)
            for k, b in enumerate(bits):
                src = packed_data.tol_true_ok if b else packed_data.tol_false_ok
                ok &= ((src >> np.uint32(k)) & np.uint32(1)).astype(bool)
            out |= ok.astype(np.uint64) << np.uint64(i)
        return out

    def loglik(self, packed_data, device=0):
        """toymodel-analytic.py:112-126 on the device: expm per edge, pruning likelihood."""
        pm = self.pm
        Pe = expm_batched(self.Q, pm.blen[1:] * pm.rate[1:], device=device)
        return pruning_loglik(pm.parent, Pe, self.prior, self.masks(packed_data), device=device)

    def posterior(self, packed_data, site=0, device=0):
        """toymodel-analytic.py:128-157: likelihood, posterior state distribution at every
        node and joint posterior of the endpoint states of every edge, for one site.
        The transition matrices come from the device; the two tree passes over
        N x n numbers are host numpy."""
        pm = self.pm
        N, n = pm.n_nodes, len(self.states)
        t = pm.blen[1:] * pm.rate[1:]
        Pe = expm_batched(self.Q, t, device=device)
        m = self.masks(packed_data)[site]
        ok = ((m[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(float)
        down = ok.copy()                       # partial likelihood of the subtree below v
        for v in range(N - 1, 0, -1):          # parent[v] < v: children are final first
            down[pm.parent[v]] *= Pe[v - 1].dot(down[v])
        lik = float(self.prior.dot(down[0]))
        if not lik > 0:
            raise Exception('the data have zero likelihood')
        post = np.zeros((N, n))
        post[0] = self.prior * down[0] / lik
        weight = np.zeros((N - 1, n))          # post[va] / (P down[vb]): the "outside" factor
        for v in range(1, N):
            msg = Pe[v - 1].dot(down[v])
            with np.errstate(divide='ignore', invalid='ignore'):
                weight[v - 1] = np.where(msg > 0, post[pm.parent[v]] / msg, 0.0)
            post[v] = (weight[v - 1][:, None] * Pe[v - 1]).sum(axis=0) * down[v]
        return dict(likelihood=lik, node_post=post, down=down, weight=weight, P=Pe, t=t)

    def expected_summary(self, packed_data, site=0, device=0):
        """Exact posterior expectations of everything ``Summary`` accumulates, for one sample
        of one site (toymodel-analytic.py:159-226 / denseutil.py:148-176), keyed like the
        ``Summary`` attributes.

        For an edge with A = Q t, endpoint weights C[a, b] = outside[a] * inside[b] and any
        indicator matrix E, sum_ab C[a, b] L(A, E)[a, b] = sum_ij E[i, j] G[i, j] with
        G = L(A^T, C): ONE Frechet derivative per edge (on the device) gives every
        transition and dwell expectation of that edge."""
        pm = self.pm
        K, n = pm.n_classes, len(self.states)
        ps = self.posterior(packed_data, site=site, device=device)
        C = ps['weight'][:, :, None] * ps['down'][1:, None, :]
        G = expm_frechet_batched(self.Q.T.copy(), ps['t'], C, device=device, compute_expm=False)
        pri = np.array([p for p, bits in self.states])
        bits = np.array([b for p, b in self.states])                 # [n, K]
        nbits = bits.sum(axis=1)
        root = ps['node_post'][0]
        out = dict(
            likelihood=ps['likelihood'],
            root_pri_to_count=dict((pm.states[p], float(root[pri == p].sum()))
                for p in range(pm.n_states)),
            root_on_total=float(root.dot(nbits)), root_off_count=float(root.dot(K - nbits)),
            root_xon_intended=float(root.dot(nbits - 1)),
            edge_to_pri_trans={}, edge_to_pri_dwell={}, edge_to_off_xon_trans={},
            edge_to_xon_off_trans={}, edge_to_off_xon_dwell={}, edge_to_xon_off_dwell={})
        Qoff = self.Q - np.diag(np.diag(self.Q))
        same_bits = (bits[:, None, :] == bits[None, :, :]).all(axis=2)
        flips = (bits[:, None, :] != bits[None, :, :]).sum(axis=2)
        turned_on = (pri[:, None] == pri[None, :]) & (flips == 1) & (nbits[None, :] > nbits[:, None])
        turned_off = (pri[:, None] == pri[None, :]) & (flips == 1) & (nbits[None, :] < nbits[:, None])
        for e, edge in enumerate(pm.edges):
            flow = ps['t'][e] * Qoff * G[e]            # expected transitions i -> j
            stay = pm.blen[e + 1] * np.diag(G[e])      # expected time in i (branch-length units)
            tr, dw = {}, {}
            for j in range(pm.nnz):
                a, b = pm.q_row[j], pm.q_col[j]
                sel = (pri[:, None] == a) & (pri[None, :] == b) & same_bits
                tr.setdefault(pm.states[a], {})[pm.states[b]] = float(flow[sel].sum())
                dw.setdefault(pm.states[a], {})[pm.states[b]] = float(
                        stay[(pri == a) & (bits[:, pm.state_class[b]] == 1)].sum())
            out['edge_to_pri_trans'][edge] = tr
            out['edge_to_pri_dwell'][edge] = dw
            out['edge_to_off_xon_trans'][edge] = float(flow[turned_on].sum())
            out['edge_to_xon_off_trans'][edge] = float(flow[turned_off].sum())
            out['edge_to_off_xon_dwell'][edge] = float(stay.dot(K - nbits))
            out['edge_to_xon_off_dwell'][edge] = float(stay.dot(nbits - 1))
        return out
