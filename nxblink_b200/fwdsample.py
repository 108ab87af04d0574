"""Forward samples of blinking-process trajectories (mirror of nxblink/fwdsample.py:18-262).

``gen_forward_samples(model, seq_of_track_to_root_state)`` keeps the reference's signature
and yields the same ``(pri_track, tol_tracks)`` objects, but all trajectories of the
sequence are simulated in ONE launch of the device Gillespie kernel
(csrc/nxb_api.cu forward_kernel, one thread per trajectory).

Own-class loss: the reference's ``gen_potential_transitions``
(fwdsample.py:57-64) tests ``track_to_state[current_tol_class]`` -- always true -- so it
also lets the tolerance class of the CURRENT primary state switch off, which leaves the
state space of the model (raoteh.py / compound.py:23-40 forbid it, and the stationary prior
puts no mass there).  ``gen_potential_transitions`` below reproduces that behaviour only
with ``reference_quirk=True``; the default, and the device kernel, implement the model.
"""
from __future__ import division, print_function

import numpy as np

from .packing import PackedModel, PRIMARY
from .sampler import BlinkSampler

__all__ = ['gen_potential_transitions', 'gen_forward_samples']


def _w(x):
    return x['weight'] if isinstance(x, dict) else x


def gen_potential_transitions(track_to_state, primary_name, primary_to_tol, Q_primary, Q_blink,
        reference_quirk=False):
    """fwdsample.py:18-64: ``((track name, initial state, final state), rate)`` for every
    transition allowed from the current joint state."""
    pri = track_to_state[primary_name]
    own = primary_to_tol[pri]
    for sb in Q_primary[pri]:
        if track_to_state[primary_to_tol[sb]]:
            yield (primary_name, pri, sb), _w(Q_primary[pri][sb])
    for tol_class in set(primary_to_tol.values()):
        if not track_to_state[tol_class]:
            yield (tol_class, 0, 1), _w(Q_blink[0][1])
        elif reference_quirk:
            if track_to_state[own]:
                yield (tol_class, 1, 0), _w(Q_blink[1][0])
        elif tol_class != own:
            yield (tol_class, 1, 0), _w(Q_blink[1][0])


def gen_forward_samples(model, seq_of_track_to_root_state, seed=None, device=0):
    """fwdsample.py:67-262.  One forward sample per entry of ``seq_of_track_to_root_state``
    (a map from track name -- 'PRIMARY' and every tolerance class -- to its root state)."""
    roots = list(seq_of_track_to_root_state)
    if not roots:
        return
    pm = model if isinstance(model, PackedModel) else PackedModel(model)
    root_pri = np.array([pm.state_index[r[PRIMARY]] for r in roots], dtype=np.int32)
    root_tol = np.zeros(len(roots), dtype=np.uint32)
    for i, r in enumerate(roots):
        for k, name in enumerate(pm.tols):
            if r[name]:
                root_tol[i] |= np.uint32(1 << k)
        if not r[pm.tols[pm.state_class[root_pri[i]]]]:
            raise Exception('the tolerance class of the root primary state must be on')
    if seed is None:
        seed = int(np.random.randint(1 << 30))
    sampler = BlinkSampler(pm, device=device)
    try:
        sampler.forward_sample(len(roots), seed=seed, root_pri=root_pri, root_tol=root_tol)
        for u in range(len(roots)):
            yield sampler.export_tracks(u)
    finally:
        sampler.close()
