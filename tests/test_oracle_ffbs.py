"""Pin the oracle's FFBS (a restatement of the un-vendored
nxmctree.sampling.sample_history) by exact enumeration: its output distribution
is fully determined by its inputs."""
import itertools

import numpy as np
from numpy.testing import assert_allclose

from oracle import blink_oracle as bo


def _random_problem(rng, n_nodes, n_states):
    children = {}
    for v in range(1, n_nodes):
        children.setdefault(int(rng.integers(0, v)), []).append(v)
    edge_to_P = {}
    for a, kids in children.items():
        for b in kids:
            P = {}
            for s in range(n_states):
                row = rng.random(n_states) * (rng.random(n_states) < 0.7)
                row[s] += 0.1
                P[s] = dict((t, float(p)) for t, p in enumerate(row / row.sum()) if p)
            edge_to_P[a, b] = P
    lmaps = {}
    for v in range(n_nodes):
        keep = [s for s in range(n_states) if rng.random() < 0.8] or [0]
        lmaps[v] = dict((s, float(rng.random() + 0.05)) for s in keep)
    prior = rng.random(n_states)
    prior = dict(enumerate(prior / prior.sum()))
    return children, edge_to_P, prior, lmaps


def _exact_joint(n_nodes, n_states, children, edge_to_P, prior, lmaps):
    joint = {}
    for x in itertools.product(range(n_states), repeat=n_nodes):
        w = prior[x[0]]
        for v in range(n_nodes):
            w *= lmaps[v].get(x[v], 0.0)
        for (a, b), P in edge_to_P.items():
            w *= P[x[a]].get(x[b], 0.0)
        if w:
            joint[x] = w
    total = sum(joint.values())
    return dict((k, v / total) for k, v in joint.items()), total


def test_partial_likelihood_equals_brute_force():
    rng = np.random.default_rng(3)
    for _ in range(5):
        prob = _random_problem(rng, 5, 3)
        children, edge_to_P, prior, lmaps = prob
        _, total = _exact_joint(5, 3, *prob)
        pl = bo.partial_likelihoods(children, edge_to_P, 0, lmaps)
        assert_allclose(sum(prior[s] * l for s, l in pl[0].items()), total, rtol=1e-12)


def test_sample_history_distribution():
    rng = np.random.default_rng(11)
    prob = _random_problem(rng, 4, 3)
    children, edge_to_P, prior, lmaps = prob
    joint, _ = _exact_joint(4, 3, *prob)
    n = 40000
    counts = {}
    for _ in range(n):
        x = bo.sample_history(rng, children, edge_to_P, 0, prior, lmaps)
        key = tuple(x[v] for v in range(4))
        counts[key] = counts.get(key, 0) + 1
    assert set(counts) <= set(joint)
    chi2 = sum((counts.get(k, 0) - n * p) ** 2 / (n * p) for k, p in joint.items())
    dof = len(joint) - 1
    assert chi2 < dof + 5 * np.sqrt(2 * dof), (chi2, dof)


def test_sample_history_single_chunk():
    # a chunk tree with zero edges (chunking.py:239,279)
    rng = np.random.default_rng(0)
    x = bo.sample_history(rng, {}, {}, 0, {0: 0.5, 1: 0.5}, {0: {1: 1.0}})
    assert x == {0: 1}
