"""Stand-in for ``nxmctree.sampling.sample_history`` -- the FFBS at the heart
of the reference's sweep (call site: nxblink/chunking.py:279-281).

TEST INFRASTRUCTURE ONLY.  The upstream source is not under /root/reference
(un-vendored, un-pinned: README.md:11-12), so this restates the published
algorithm: draw a joint state assignment with probability proportional to

    prior(x_root) * prod_{edges (a,b)} P_ab[x_a][x_b] * prod_{nodes v} lmap_v[x_v]

by an upward pass of partial likelihoods followed by a root-to-leaves sampling
pass.  Its output *distribution* is fully determined by its inputs, which is
what tests/test_oracle_ffbs.py pins by exact enumeration on small trees.

Contract details taken from the call sites:
  * ``T`` may be an empty DiGraph with ``root`` not a node of it (a chunk tree
    with a single chunk: nxblink/chunking.py:239,279).
  * ``edge_to_P[(a, b)]`` is an nx.DiGraph with 'weight' attributes.
  * ``node_to_data_lmap[v]`` maps feasible states to emission likelihoods.
  * returns {node: state}; returns None if the data are infeasible.
"""
import numpy as np


def dict_random_choice(d):
    keys = list(d)
    w = np.array([d[k] for k in keys], dtype=float)
    total = w.sum()
    if not total > 0:
        return None
    i = np.searchsorted(np.cumsum(w), np.random.uniform(0, total), side='right')
    return keys[min(i, len(keys) - 1)]


def _children(T, v):
    return list(T[v]) if v in T else []


def _postorder(T, root):
    order, stack = [], [root]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(_children(T, v))
    return order[::-1]


def get_partial_likelihoods(T, edge_to_P, root, node_to_data_lmap):
    """node -> {state: likelihood of everything at and below the node}."""
    node_to_plmap = {}
    for v in _postorder(T, root):
        plmap = {}
        for s, lk in node_to_data_lmap[v].items():
            val = lk
            for c in _children(T, v):
                P = edge_to_P[v, c]
                sub = node_to_plmap[c]
                acc = 0.0
                if s in P:
                    for sb, d in P[s].items():
                        if sb in sub:
                            acc += d['weight'] * sub[sb]
                val *= acc
                if not val:
                    break
            if val:
                plmap[s] = val
        node_to_plmap[v] = plmap
    return node_to_plmap


def sample_history(T, edge_to_P, root, root_prior_distn, node_to_data_lmap):
    node_to_plmap = get_partial_likelihoods(
            T, edge_to_P, root, node_to_data_lmap)
    weights = dict((s, root_prior_distn[s] * lk)
            for s, lk in node_to_plmap[root].items()
            if root_prior_distn.get(s, 0))
    root_state = dict_random_choice(weights) if weights else None
    if root_state is None:
        return None
    node_to_state = {root: root_state}
    stack = [root]
    while stack:
        v = stack.pop()
        sa = node_to_state[v]
        for c in _children(T, v):
            P = edge_to_P[v, c]
            sub = node_to_plmap[c]
            w = dict((sb, d['weight'] * sub[sb])
                    for sb, d in P[sa].items() if sb in sub)
            node_to_state[c] = dict_random_choice(w)
            stack.append(c)
    return node_to_state
