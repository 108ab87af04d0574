"""Rao-Teh CTBN sampling for the blinking model -- reference-compatible facade.

``gen_samples(model, data, nburnin, nsamples)`` has the signature and yields the
same ``(primary_track, tolerance_tracks)`` pairs as nxblink/raoteh.py:25-157, but
every sweep runs on the GPU (one site, one chain per call, as in the reference).
Like the reference it re-yields the same mutable track objects each sweep.
For throughput use ``nxblink_b200.BlinkSampler`` directly: thousands of sites x
chains per launch with the statistics reduced on the device.
"""
from __future__ import division, print_function

import numpy as np

from .sampler import BlinkSampler

__all__ = ['gen_samples']


def gen_samples(model, data, nburnin, nsamples, seed=None, device=0, **kwargs):
    """Sample trajectories from the conditional distribution (raoteh.py:25-157).

    ``seed=None`` draws a fresh seed (the reference uses the global unseeded
    ``numpy.random``).
    """
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(2, np.uint32).astype(np.uint64).dot(
            np.array([1, 1 << 32], dtype=np.uint64)))
    sampler = BlinkSampler(model, device=device, **kwargs)
    try:
        sampler.set_data([data])
        sampler.init_chains(chains=1, seed=seed)
        if nburnin:
            sampler.sweep(nburnin, accumulate=False)
        primary_track, tolerance_tracks = None, None
        for _ in range(nsamples):
            sampler.sweep(1, accumulate=False)
            pri, tols = sampler.export_tracks(0)
            if primary_track is None:
                primary_track, tolerance_tracks = pri, tols
            else:
                # same mutable objects every sweep (raoteh.py:152,435)
                primary_track.history, primary_track.events = pri.history, pri.events
                for ev_list in pri.events.values():
                    for ev in ev_list:
                        ev.track = primary_track
                for old, new in zip(tolerance_tracks, tols):
                    old.history, old.events = new.history, new.events
                    for ev_list in new.events.values():
                        for ev in ev_list:
                            ev.track = old
            yield primary_track, tolerance_tracks
    finally:
        sampler.close()
