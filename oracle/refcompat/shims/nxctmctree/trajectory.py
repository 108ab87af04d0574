"""Stand-ins for ``nxctmctree.trajectory.{Event, LightTrajectory, get_node_to_tm}``.

TEST INFRASTRUCTURE ONLY.  Contracts taken from the reference call sites:

* ``Event(track, tm, sa, sb)`` positional (nxblink/tests/test_navigation.py:20,
  nxblink/fwdsample.py:251) and keyword with ``sa``/``sb`` optional
  (nxblink/raoteh.py:225,281; nxblink/poisson.py:48).  Hashable by identity
  (used as a dict key, nxblink/poisson.py:109,122).
* ``LightTrajectory(name, history, events)`` positional
  (test_navigation.py:16) and keyword (fwdsample.py:152-155);
  ``clear_state_labels()`` nulls node states and event labels,
  ``remove_self_transitions()`` drops events with ``sa == sb``
  (nxblink/raoteh.py:348,361,405,411,426,432).
* ``get_node_to_tm(T, root, edge_to_blen)`` -> cumulative branch length from
  the root (nxblink/util.py:10, nxblink/raoteh.py:79).
"""
import networkx as nx


class Event(object):
    __slots__ = ('track', 'tm', 'sa', 'sb')

    def __init__(self, track=None, tm=None, sa=None, sb=None):
        self.track = track
        self.tm = tm
        self.sa = sa
        self.sb = sb

    def __repr__(self):
        name = getattr(self.track, 'name', None)
        return 'Event(%r, %r, %r, %r)' % (name, self.tm, self.sa, self.sb)

    # ``sorted((ev.tm, ev) ...)`` is used all over the reference; ties in tm
    # have probability zero but give a deterministic order anyway.
    def __lt__(self, other):
        return id(self) < id(other)


class LightTrajectory(object):
    def __init__(self, name=None, history=None, events=None):
        self.name = name
        self.history = history
        self.events = events

    def clear_state_labels(self):
        for v in self.history:
            self.history[v] = None
        for edge, events in self.events.items():
            for ev in events:
                ev.sa = None
                ev.sb = None

    def remove_self_transitions(self):
        for edge in list(self.events):
            self.events[edge] = [
                    ev for ev in self.events[edge] if ev.sa != ev.sb]


def get_node_to_tm(T, root, edge_to_blen):
    node_to_tm = {root: 0}
    for va, vb in nx.bfs_edges(T, root):
        node_to_tm[vb] = node_to_tm[va] + edge_to_blen[va, vb]
    return node_to_tm
