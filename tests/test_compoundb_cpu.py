"""Host pieces of the compound-state mirror (nxblink/compoundb.py:23-193): state
enumeration and rate matrix against the reference's dense toy enumeration
(nxblink/denseutil.py:39-62: 24 feasible of 48 states) and the restated oracle."""
import numpy as np
from numpy.testing import assert_allclose

from nxblink_b200 import compoundb, toymodel
from oracle.dense_oracle import DensePosterior
from nxblink_b200 import toydata


def _p2t(M):
    return dict((p, int(t[1:])) for p, t in M.get_primary_to_tol().items())


def test_gen_states_counts_and_consistency():
    p2t = _p2t(toymodel.BlinkModelA)
    states = list(compoundb.gen_states(p2t))
    assert len(states) == 6 * 4 and len(set(states)) == 24
    assert all(s.tol[p2t[s.pri]] == 1 for s in states)


def test_get_Q_equals_the_oracle_generator():
    for M in (toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC):
        p2t = _p2t(M)
        Q = compoundb.get_Q(M.get_Q_primary(), M.get_rate_on(), M.get_rate_off(), p2t)
        dp = DensePosterior(M, toydata.DataA)
        # oracle Q is built on the rate-1 normalised primary matrix
        scale = dp.tb.Q_primary[0][1] / M.get_Q_primary()[0][1]['weight']
        idx = dict(((p, tuple(bits)), i) for i, (p, bits) in enumerate(dp.comp))
        dense = np.zeros((24, 24))
        for a, b in Q.edges():
            w = Q[a][b]['weight']
            if a.pri != b.pri:
                w *= scale
            dense[idx[(a.pri, tuple(a.tol))], idx[(b.pri, tuple(b.tol))]] = w
        off = dp.Q - np.diag(np.diag(dp.Q))
        assert_allclose(dense, off, rtol=1e-13)


def test_successors_flip_one_component():
    M = toymodel.BlinkModelB
    p2t = _p2t(M)
    s = compoundb.State(0, [1, 0, 1])
    succ = list(compoundb.gen_successors(M.get_Q_primary(), 2.0, 3.0, p2t, s))
    for sb, rate in succ:
        assert sum(x != y for x, y in zip(s.tolist(), sb.tolist())) == 1
        assert compoundb.state_is_consistent(p2t, sb)
    # own class (T0) cannot switch off; T1 can switch on (rate 2), T2 off (rate 3)
    tol_moves = sorted((tuple(sb.tol), r) for sb, r in succ if sb.pri == 0)
    assert tol_moves == [((1, 0, 0), 3.0), ((1, 1, 1), 2.0)]
