"""Expectation maximisation for the codon blinking model (mirror of
examples/p53/p53em.py:58-257 and the EM loop of examples/p53/run-stochastic.py:154-208).

E-step: batched Rao-Teh sweeps on the GPU, statistics reduced on the device
(``BlinkSampler``; across GPUs one all-reduce, ``distributed.allreduce_summary``).
M-step: the reference's objective -- minus (get_ll_root + get_ll_dwell + get_ll_trans)
of nxblink/em.py over log-parameters (kappa, omega, A, C, G, T, blink on, blink off, one
rate per edge; p53em.py:143-185) -- minimised with L-BFGS-B.  The reference differentiates
it with algopy forward-mode AD; here the objective is evaluated vectorised on the flat
device statistics together with its closed-form gradient (``MStep.neg_ll_and_grad``).
"""
from __future__ import division, print_function

import numpy as np

from .p53 import Model
from .sampler import BlinkSampler
from .packing import PackedModel


def mg94_tables(genetic_code):
    """Index tables of the MG94 pre-rate matrix (p53em.py:188-229 get_pre_Q): for every
    single-nucleotide codon pair (a, b): target nucleotide, transition?, nonsynonymous?"""
    nts = 'ACGT'
    transitions = ('AG', 'GA', 'CT', 'TC')
    rows, cols, nt_to, is_ts, is_non = [], [], [], [], []
    for a, (sa, ra, ca) in enumerate(genetic_code):
        for b, (sb, rb, cb) in enumerate(genetic_code):
            diff = [(x, y) for x, y in zip(ca, cb) if x != y]
            if len(diff) != 1:
                continue
            x, y = diff[0]
            rows.append(a); cols.append(b); nt_to.append(nts.index(y))
            is_ts.append(x + y in transitions); is_non.append(ra != rb)
    counts = np.array([[c.count(n) for n in nts] for s, r, c in genetic_code], dtype=float)
    return (np.array(rows), np.array(cols), np.array(nt_to), np.array(is_ts),
            np.array(is_non), counts)


def unpack_params(log_params):
    """p53em.py:143-185: log-parameters -> (kappa, omega, nt[4] normalised, on, off, edge rates)
    and the simplex penalty (computed by the reference but not added to its objective)."""
    prm = np.exp(np.asarray(log_params, dtype=float))
    kappa, omega = prm[0], prm[1]
    nt = prm[2:6]
    total = nt.sum()
    penalty = np.log(total) ** 2
    return kappa, omega, nt / total, prm[6], prm[7], prm[8:], penalty


class MStep(object):
    """Vectorised M-step objective on the flat statistics of a device ``Summary``."""

    def __init__(self, pm, summary, genetic_code):
        self.pm = pm
        f = summary.flat
        self.n = float(summary.nsamples)
        tabs = mg94_tables(genetic_code)
        self.rows, self.cols, self.nt_to, self.is_ts, self.is_non, self.nt_counts = tabs
        # the CSR order of PackedModel is (row, col) sorted = the order of these tables
        assert (self.rows == pm.q_row).all() and (self.cols == pm.q_col).all()
        self.pri_trans = f['pri_trans'].astype(float)         # [E, nnz]
        self.pri_dwell = f['pri_dwell']                        # [E, nnz]
        self.off_on = f['off_on'].astype(float)
        self.on_off = f['on_off'].astype(float)
        self.off_xon_dwell = f['off_xon_dwell']
        self.xon_off_dwell = f['xon_off_dwell']
        self.root_pri = f['root_pri'].astype(float)
        self.root_off = float(summary.root_off_count)
        self.root_xon = float(summary.root_xon_count)
        self.trans_total = self.pri_trans.sum(axis=1) + self.off_on + self.on_off

    def pre_Q(self, kappa, omega, nt):
        q = nt[self.nt_to] * np.where(self.is_ts, kappa, 1.0) * np.where(self.is_non, omega, 1.0)
        return q

    def distn(self, nt):
        w = np.prod(nt[None, :] ** self.nt_counts, axis=1)
        return w / w.sum()

    def neg_ll(self, log_params):
        """-(em.get_ll_root + em.get_ll_dwell + em.get_ll_trans), p53em.py:58-91."""
        kappa, omega, nt, on, off, rates, _ = unpack_params(log_params)
        q = self.pre_Q(kappa, omega, nt)
        distn = self.distn(nt)
        expected_rate = np.dot(np.bincount(self.rows, weights=q, minlength=len(distn)), distn)
        p_off, p_on = off / (on + off), on / (on + off)
        ll_root = (np.dot(self.root_pri, np.log(distn)) + self.root_off * np.log(p_off)
                + self.root_xon * np.log(p_on))
        dwell_e = self.pri_dwell.dot(q) + self.off_xon_dwell * on + self.xon_off_dwell * off
        ll_dwell = -np.dot(dwell_e, rates) / expected_rate
        with np.errstate(divide='ignore'):
            logq = np.log(q)
        ll_trans = (self.pri_trans.dot(logq).sum() + self.off_on.sum() * np.log(on)
                + self.on_off.sum() * np.log(off)
                + np.dot(self.trans_total, np.log(rates / expected_rate)))
        return -(ll_root + ll_dwell + ll_trans) / self.n

    def neg_ll_and_grad(self, log_params):
        """``neg_ll`` and its exact gradient in the log-parameters (closed form; the reference
        gets the same numbers from algopy forward-mode AD, p53em.py:39-56,94-140).

        With q_i = nt[to_i] kappa^[ts_i] omega^[non_i], D_e = sum_i dwell_ei q_i + on*offdwell_e
        + off*ondwell_e, A = sum_e rate_e D_e, TT = sum_e transitions_e and
        ER = sum_i distn[row_i] q_i:
            n*LL = root_pri.log(distn) + root_off log p_off + root_xon log p_on - A/ER
                   + sum_ei trans_ei log q_i + OffOn log on + OnOff log off
                   + sum_e transitions_e (log rate_e - log ER).
        """
        kappa, omega, nt, on, off, rates, _ = unpack_params(log_params)
        q = self.pre_Q(kappa, omega, nt)
        distn = self.distn(nt)
        P = len(distn)
        rowsum = np.bincount(self.rows, weights=q, minlength=P)
        ER = np.dot(rowsum, distn)
        p_off, p_on = off / (on + off), on / (on + off)
        D = self.pri_dwell.dot(q) + self.off_xon_dwell * on + self.xon_off_dwell * off
        A = np.dot(D, rates)
        TT = self.trans_total.sum()
        with np.errstate(divide='ignore'):
            logq = np.log(q)
        ll = (np.dot(self.root_pri, np.log(distn)) + self.root_off * np.log(p_off)
                + self.root_xon * np.log(p_on) - A / ER
                + self.pri_trans.dot(logq).sum() + self.off_on.sum() * np.log(on)
                + self.on_off.sum() * np.log(off)
                + np.dot(self.trans_total, np.log(rates / ER)))
        g = np.zeros(len(log_params))
        dER = A / ER ** 2 - TT / ER                      # d(n LL) / d ER
        # d / d log q_i: direct terms + through ER
        G = -rates.dot(self.pri_dwell) * q / ER + self.pri_trans.sum(axis=0) \
                + dER * distn[self.rows] * q
        g[0] = G[self.is_ts].sum()
        g[1] = G[self.is_non].sum()
        # nucleotide frequencies enter through q and through the codon distribution
        dlogdistn = self.root_pri + dER * rowsum * distn           # d(n LL) / d log distn_s
        cbar = distn.dot(self.nt_counts)
        h = np.bincount(self.nt_to, weights=G, minlength=4) \
                + dlogdistn.dot(self.nt_counts - cbar[None, :])
        g[2:6] = h - nt * h.sum()
        g[6] = (-self.root_off * p_on + self.root_xon * p_off
                - np.dot(rates, self.off_xon_dwell) * on / ER + self.off_on.sum())
        g[7] = (self.root_off * p_on - self.root_xon * p_off
                - np.dot(rates, self.xon_off_dwell) * off / ER + self.on_off.sum())
        g[8:] = -rates * D / ER + self.trans_total
        return -ll / self.n, -g / self.n

    def closed_form_edge_rates(self, kappa, omega, nt, on, off):
        """Stationary point in the edge rates: rate_e = ER * transitions_e / dwell_e."""
        q = self.pre_Q(kappa, omega, nt)
        distn = self.distn(nt)
        expected_rate = np.dot(np.bincount(self.rows, weights=q, minlength=len(distn)), distn)
        dwell_e = self.pri_dwell.dot(q) + self.off_xon_dwell * on + self.xon_off_dwell * off
        return expected_rate * self.trans_total / np.maximum(dwell_e, 1e-300)


def maximization_step(mstep, kappa, omega, A, C, G, T, blink_on, blink_off, edge_rates,
        maxiter=200):
    """p53em.py:94-140: L-BFGS-B over the log-parameters; edges with no sampled event keep
    a tiny positive rate (the reference sets them to zero, p53em.py:25-27)."""
    import scipy.optimize
    rates = np.maximum(np.asarray(edge_rates, dtype=float), 1e-12)
    x0 = np.log(np.concatenate([[kappa, omega, A, C, G, T, blink_on, blink_off], rates]))
    res = scipy.optimize.minimize(mstep.neg_ll_and_grad, x0, method='L-BFGS-B', jac=True,
            options=dict(maxiter=maxiter))
    kappa, omega, nt, on, off, rates, penalty = unpack_params(res.x)
    return dict(kappa=kappa, omega=omega, A=nt[0], C=nt[1], G=nt[2], T=nt[3], rate_on=on,
            rate_off=off, edge_rates=rates, neg_ll=res.fun, penalty=penalty, result=res)


def em_iteration(model, packed_data, params, chains=8, nburnin=10, nsamples=10, seed=0,
        device=0, reduce_fn=None, sampler=None):
    """One EM iteration (run-stochastic.py:154-208): build the model from ``params``, run
    the E-step on the device, maximise.  ``reduce_fn(counts, dwell)`` may all-reduce the
    statistics across ranks (nxblink_b200.distributed).  Pass the ``sampler`` of the previous
    iteration (returned as ``new['sampler']`` when one was given) to reuse its device buffers:
    only the model values are replaced (``BlinkSampler.update_model``)."""
    pm0 = PackedModel(model)
    edge_to_rate = dict(zip(pm0.edges, params['edge_rates']))
    m = Model(params['kappa'], params['omega'], params['A'], params['C'], params['G'], params['T'],
            params['rate_on'], params['rate_off'], model._tree, model._root,
            model.get_edge_to_blen(), edge_to_rate, model.genetic_code)
    pm = PackedModel(m)
    keep = sampler is not None
    if keep:
        sampler.update_model(pm)
        sampler.reset_summary()
    else:
        sampler = BlinkSampler(pm, device=device)
    try:
        if not keep or sampler.data is not packed_data:
            sampler.set_data(packed_data)
        sampler.init_chains(chains=chains, seed=seed)
        sampler.sweep(nburnin, accumulate=False)
        sampler.sweep(nsamples, accumulate=True)
        counts, dwell = sampler.read_summary()
        if reduce_fn is not None:
            counts, dwell = reduce_fn(counts, dwell)
        summary = sampler.summary(counts, dwell)
    finally:
        if not keep:
            sampler.close()
    mstep = MStep(pm, summary, m.genetic_code)
    new = maximization_step(mstep, params['kappa'], params['omega'], params['A'], params['C'],
            params['G'], params['T'], params['rate_on'], params['rate_off'], params['edge_rates'])
    new['summary'] = summary
    new['mstep'] = mstep
    if keep:
        new['sampler'] = sampler
    return new
