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


class DeviceMStep(object):
    """The same objective and minimisation on the DEVICE-resident statistics (``nxb_mstep``:
    csrc/mstep_kernel.cu): objective + closed-form gradient from the flat statistic buffers, L-BFGS
    inside one kernel launch.  ``counts_ptr`` / ``dwell_ptr``: device pointers (default: the
    sampler's own accumulators; after an all-reduce pass the reduced copies)."""

    def __init__(self, sampler, genetic_code, counts_ptr=None, dwell_ptr=None):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        pm = self.pm = sampler.pm
        self.device = getattr(sampler, 'device', 0)
        self.layout = sampler.layout
        if counts_ptr is None:
            counts_ptr, dwell_ptr = sampler.summary_device_ptrs()
        self.counts_ptr, self.dwell_ptr = int(counts_ptr), int(dwell_ptr)
        rows, cols, nt_to, is_ts, is_non, nt_counts = mg94_tables(genetic_code)
        assert (rows == pm.q_row).all() and (cols == pm.q_col).all()
        # summary.py:221: an ON track counts unless the class NAME equals True
        xon_w = np.array([0.0 if (name == True) else 1.0 for name in pm.tols])     # noqa: E712
        self._keep = dict(q_ptr=np.ascontiguousarray(pm.q_ptr, dtype=np.int32),
                q_col=np.ascontiguousarray(pm.q_col, dtype=np.int32),
                cls=np.ascontiguousarray(pm.state_class, dtype=np.int32),
                nt_to=np.ascontiguousarray(nt_to, dtype=np.int32),
                is_ts=np.ascontiguousarray(is_ts, dtype=np.uint8),
                is_non=np.ascontiguousarray(is_non, dtype=np.uint8),
                nt_counts=np.ascontiguousarray(nt_counts, dtype=np.float64),
                xon_w=np.ascontiguousarray(xon_w, dtype=np.float64))
        k = self._keep
        d = self.desc = _lib.MStepDesc()
        d.n_edges, d.n_states, d.n_classes, d.nnz = pm.n_nodes - 1, pm.n_states, pm.n_classes, pm.nnz
        d.q_ptr, d.q_col, d.state_class = k['q_ptr'].ctypes.data, k['q_col'].ctypes.data, k['cls'].ctypes.data
        d.nt_to, d.is_ts, d.is_non = k['nt_to'].ctypes.data, k['is_ts'].ctypes.data, k['is_non'].ctypes.data
        d.nt_counts, d.xon_weight = k['nt_counts'].ctypes.data, k['xon_w'].ctypes.data
        self.n = 8 + d.n_edges
        self._C = C

    def _call(self, x, maxiter, eval_only):
        C = self._C
        x = np.array(x, dtype=np.float64)
        assert x.shape == (self.n,)
        f = C.c_double()
        g = np.zeros(self.n)
        info = np.zeros(3, dtype=np.int32)
        rc = self.lib.nxb_mstep(int(self.device), C.byref(self.desc), C.byref(self.layout),
                self.counts_ptr, self.dwell_ptr, x.ctypes.data, int(maxiter), int(eval_only),
                C.byref(f), g.ctypes.data, info.ctypes.data)
        if rc:
            from ._lib import NxbError
            raise NxbError(rc, self.lib.nxb_mstep_last_error().decode())
        return x, f.value, g, info

    def neg_ll_and_grad(self, log_params):
        _, f, g, _ = self._call(log_params, 0, 1)
        return f, g

    def minimize(self, x0, maxiter=200):
        x, f, g, info = self._call(x0, maxiter, 0)
        return x, f, g, dict(iterations=int(info[0]), evaluations=int(info[1]), status=int(info[2]),
                kernel_ms=float(self.lib.nxb_mstep_last_kernel_ms()))


def m_step(sampler, summary, genetic_code, kappa, omega, A, C, G, T, blink_on, blink_off, edge_rates,
        maxiter=200, where='device', counts_ptr=None, dwell_ptr=None):
    """The M-step of one EM iteration (p53em.py:94-140).  ``where='device'``: on the statistics
    resident on the GPU (``DeviceMStep``); ``'host'``: numpy objective + scipy L-BFGS-B on
    ``summary`` (the check of the device path, tests/test_gpu_em.py)."""
    if where == 'host':
        mstep = MStep(sampler.pm, summary, genetic_code)
        new = maximization_step(mstep, kappa, omega, A, C, G, T, blink_on, blink_off, edge_rates, maxiter)
        new['mstep'] = mstep
        new['where'] = 'host (numpy objective + closed-form gradient, scipy L-BFGS-B)'
        return new
    dm = DeviceMStep(sampler, genetic_code, counts_ptr, dwell_ptr)
    rates = np.maximum(np.asarray(edge_rates, dtype=float), 1e-12)
    x0 = np.log(np.concatenate([[kappa, omega, A, C, G, T, blink_on, blink_off], rates]))
    x, f, g, info = dm.minimize(x0, maxiter)
    kappa, omega, nt, on, off, rates, penalty = unpack_params(x)
    return dict(kappa=kappa, omega=omega, A=nt[0], C=nt[1], G=nt[2], T=nt[3], rate_on=on,
            rate_off=off, edge_rates=rates, neg_ll=f, penalty=penalty, result=info, x=x, grad=g,
            mstep=dm, where='device (nxb_mstep: objective, gradient and L-BFGS in one kernel launch)')


def em_iteration(model, packed_data, params, chains=8, nburnin=10, nsamples=10, seed=0,
        device=0, reduce_fn=None, sampler=None, mstep='device', reduce_device_fn=None):
    """One EM iteration (run-stochastic.py:154-208): build the model from ``params``, run
    the E-step on the device, maximise.  Pass the ``sampler`` of the previous iteration
    (returned as ``new['sampler']`` when one was given) to reuse its device buffers: only the
    model values are replaced (``BlinkSampler.update_model``).

    ``mstep='device'`` (default): the M-step runs on the statistics where they are, on the GPU
    (``DeviceMStep``); with several GPUs ``reduce_device_fn(sampler)`` all-reduces them on the
    device and returns the (counts, dwell) device pointers of the reduced buffers.
    ``mstep='host'``: statistics are copied out (``reduce_fn(counts, dwell)`` may all-reduce host
    copies) and the numpy/scipy M-step runs on them -- the checker of the device path."""
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
    if reduce_fn is not None:
        mstep = 'host'
    args = (params['kappa'], params['omega'], params['A'], params['C'], params['G'], params['T'],
            params['rate_on'], params['rate_off'], params['edge_rates'])
    try:
        if not keep or sampler.data is not packed_data:
            sampler.set_data(packed_data)
        sampler.init_chains(chains=chains, seed=seed)
        sampler.sweep(nburnin, accumulate=False)
        sampler.sweep(nsamples, accumulate=True)
        if mstep == 'device':
            ptrs = reduce_device_fn(sampler) if reduce_device_fn is not None else (None, None)
            new = m_step(sampler, None, m.genetic_code, *args, where='device',
                    counts_ptr=ptrs[0], dwell_ptr=ptrs[1])
            summary = sampler.summary()
        else:
            counts, dwell = sampler.read_summary()
            if reduce_fn is not None:
                counts, dwell = reduce_fn(counts, dwell)
            summary = sampler.summary(counts, dwell)
            new = m_step(sampler, summary, m.genetic_code, *args, where='host')
    finally:
        if not keep:
            sampler.close()
    new['summary'] = summary
    if keep:
        new['sampler'] = sampler
    return new
