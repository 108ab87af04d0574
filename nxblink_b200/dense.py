"""Deterministic dense path on the device: batched matrix exponentials, the pruning
likelihood, and the explicit compound process for small models.

Mirrors the reference's test-only dense machinery
    nxblink/compound.py:23-90       compound_state_is_ok, get_Q_compound
    nxblink/denseutil.py:39-146     compound states, nx <-> dense matrices
    examples/codon2x3/toymodel-analytic.py:86-126   Q_compound, expm per edge, likelihood
with the two numerical kernels (``expm`` per edge, pruning recursion over sites) running
as FP64 tensor-pipe CUDA kernels behind ``nxb_expm_batched`` / ``nxb_pruning_loglik``.
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
