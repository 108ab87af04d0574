"""Validation asserts named in examples/p53/create_mg94.py:55-57.  TEST INFRASTRUCTURE ONLY."""
from numpy.testing import assert_allclose


def assert_distn(distn):
    assert all(p >= 0 for p in distn.values())
    assert_allclose(sum(distn.values()), 1)


assert_nxdistn = assert_distn


def assert_rate_matrix(Q):
    for sa, sb in Q.edges():
        assert sa != sb and Q[sa][sb]['weight'] >= 0


def assert_equilibrium(Q, distn):
    for s in Q:
        out = sum(distn[s] * Q[s][b]['weight'] for b in Q[s])
        inn = sum(distn[a] * Q[a][s]['weight'] for a in Q.predecessors(s))
        assert_allclose(out, inn, atol=1e-12)


def assert_detailed_balance(Q, distn):
    for sa, sb in Q.edges():
        assert_allclose(distn[sa] * Q[sa][sb]['weight'],
                distn[sb] * Q[sb][sa]['weight'], atol=1e-12)
