"""Multi-GPU sharding of the Rao-Teh sweep: one process per GPU.

Units (site, chain) are independent within an E-step, so alignment sites are
split into contiguous blocks, one per rank, with NO data-path collective.  The
single exchange step is the reduction of the sufficient statistics (the role the
single ``Summary`` object plays across sites in
examples/p53/run-stochastic.py:172-195): one all-reduce(sum) of the u64 count
buffer and one of the f64 dwell buffer, on device copies of the buffers
(NCCL over NVLink; ``gloo`` on host buffers for the CPU tests).

Random streams are keyed by GLOBAL unit index, so integer statistics are
identical for any number of ranks; dwell sums differ only by reduction order.
"""
from __future__ import division, print_function

import os

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [start, stop) of n items owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _DevArray(object):
    """Expose a raw device pointer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(
                shape=(n,), typestr=typestr, data=(int(ptr), False), version=2)


def device_summary_tensors(sampler, device=None):
    """torch views (int64 counts, float64 dwell) of the sampler's device buffers (zero copy).
    ``device``: the CUDA device index the sampler was created on (default: its ``device``
    attribute, else torch's current device)."""
    import torch
    c_ptr, d_ptr = sampler.summary_device_ptrs()
    if device is None:
        device = getattr(sampler, 'device', None)
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
    counts = torch.as_tensor(_DevArray(c_ptr, sampler.layout.n_counts, '<i8'), device=dev)
    dwell = torch.as_tensor(_DevArray(d_ptr, sampler.layout.n_dwell, '<f8'), device=dev)
    return counts, dwell


def allreduce_summary(counts, dwell, group=None):
    """In-place sum over ranks of the two statistic buffers (torch tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(dwell, op=dist.ReduceOp.SUM, group=group)
    return counts, dwell


def allreduce_host_summary(counts, dwell, group=None):
    """Same for host numpy buffers (CPU/gloo path used by the host-logic tests)."""
    import torch
    c = torch.from_numpy(np.ascontiguousarray(counts).view(np.int64).copy())
    d = torch.from_numpy(np.ascontiguousarray(dwell, dtype=np.float64).copy())
    allreduce_summary(c, d, group)
    return c.numpy().view(np.uint64), d.numpy()


def env_rank_world():
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)),
            int(os.environ.get('WORLD_SIZE', 1)))


class ShardedSampler(object):
    """A ``BlinkSampler`` on this rank's GPU holding this rank's block of sites."""

    def __init__(self, model, packed_data, rank, world, device=0, **kwargs):
        from .sampler import BlinkSampler
        from .packing import PackedData
        self.rank, self.world = rank, world
        self.site_start, self.site_stop = shard_range(packed_data.n_sites, rank, world)
        sl = slice(self.site_start, self.site_stop)
        self.sampler = BlinkSampler(model, device=device, **kwargs)
        self.local = PackedData(packed_data.pri_ok[sl], packed_data.tol_true_ok[sl],
                packed_data.tol_false_ok[sl])
        self.sampler.set_data(self.local)

    def init_chains(self, chains, seed):
        self.sampler.init_chains(chains=chains, seed=seed,
                unit_offset=self.site_start * chains)

    def sweep(self, n_sweeps, accumulate=True):
        self.sampler.sweep(n_sweeps, accumulate=accumulate)

    def allreduce(self):
        """NCCL all-reduce of the device statistics; returns the GLOBAL (counts, dwell) as host
        arrays.  The reduction works on copies: the sampler's own accumulators keep this rank's
        statistics, so accumulating further sweeps and reducing again never counts anything twice."""
        import torch
        counts, dwell = device_summary_tensors(self.sampler)
        counts, dwell = counts.clone(), dwell.clone()
        allreduce_summary(counts, dwell)
        return counts.cpu().numpy().view(np.uint64), dwell.cpu().numpy()

    def allreduce_device(self):
        """All-reduce on the device and STAY there: returns the device pointers (counts, dwell) of
        the reduced copies, for the on-device M-step (``p53em.em_iteration(reduce_device_fn=...)``).
        The copies live until the next call."""
        return allreduce_device(self.sampler, self)


def allreduce_device(sampler, keep=None):
    """``reduce_device_fn`` for ``p53em.em_iteration``: device all-reduce of copies of the sampler's
    statistics; returns their device pointers.  ``keep`` (default: the sampler) holds the tensors."""
    import torch
    counts, dwell = device_summary_tensors(sampler)
    counts, dwell = counts.clone(), dwell.clone()
    allreduce_summary(counts, dwell)
    torch.cuda.synchronize()
    (keep if keep is not None else sampler)._reduced = (counts, dwell)
    return counts.data_ptr(), dwell.data_ptr()
