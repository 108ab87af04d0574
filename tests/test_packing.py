"""Host logic: model/data packing (raoteh.py:59-119,160-184 mirror)."""
import json
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from nxblink_b200 import packing, toymodel, toydata, p53
from oracle import blink_oracle as bo


@pytest.mark.parametrize('M', [toymodel.BlinkModelA, toymodel.BlinkModelB, toymodel.BlinkModelC])
def test_packed_model_matches_oracle_tables(M):
    pm = packing.PackedModel(M)
    tb = bo.build_tables(M)
    assert pm.parent[0] == -1 and (pm.parent[1:] < np.arange(1, pm.n_nodes)).all()
    assert pm.nodes[0] == tb.root
    assert_allclose(pm.expected_rate, tb.expected_primary_rate, rtol=1e-14)
    for i in range(pm.nnz):
        sa, sb = pm.states[pm.q_row[i]], pm.states[pm.q_col[i]]
        assert_allclose(pm.q_val[i], tb.Q_primary[sa][sb], rtol=1e-14)
    assert pm.nnz == sum(len(r) for r in tb.Q_primary.values())
    # expected rate of the normalised matrix is 1 (raoteh.py:90-102)
    assert_allclose(sum(pm.prior[pm.q_row[i]] * pm.q_val[i] for i in range(pm.nnz)), 1.0, rtol=1e-14)
    for v, node in enumerate(pm.nodes):
        assert_allclose(pm.node_tm[v], tb.node_to_tm[node], rtol=1e-15)
    assert pm.diameter == bo._diameter(tb.Q_primary)
    assert [pm.primary_to_tol[s] for s in pm.states] == [pm.tols[c] for c in pm.state_class]
    # the caller's graph must not be mutated (unlike raoteh.py:101-102)
    Q = M.get_Q_primary()
    before = dict(((a, b), Q[a][b]['weight']) for a, b in Q.edges())

    class Wrap(M):
        @classmethod
        def get_Q_primary(cls):
            return Q
    packing.PackedModel(Wrap)
    assert before == dict(((a, b), Q[a][b]['weight']) for a, b in Q.edges())


def test_zero_length_pooling_matches_oracle():
    M = toymodel.BlinkModelC          # two zero-rate edges N1-N0, N1-N2
    pm = packing.PackedModel(M)
    pd = packing.pack_data(pm, [toydata.DataC])
    tb = bo.build_tables(M)
    pri, tols = bo.make_tracks(tb, toydata.DataC)
    bo._pool_data_for_zero_blen(tb, [pri] + tols)
    for v, node in enumerate(pm.nodes):
        got = set(pm.states[i] for i in range(pm.n_states) if (int(pd.pri_ok[0, v]) >> i) & 1)
        assert got == pri.data[node]
        for k, t in enumerate(tols):
            fset = set()
            if (int(pd.tol_true_ok[0, v]) >> k) & 1:
                fset.add(True)
            if (int(pd.tol_false_ok[0, v]) >> k) & 1:
                fset.add(False)
            assert fset == t.data[node], (node, t.name)


def test_empty_pool_raises_like_reference():
    # raoteh.py:178-180
    class Bad(toydata.DataB):
        PRIMARY = dict(toydata.DataB.PRIMARY, N1=(3,), N2=(5,))
    pm = packing.PackedModel(toymodel.BlinkModelC)
    with pytest.raises(Exception):
        packing.pack_data(pm, [Bad])


def test_p53_shape_and_mg94_against_reference(golden_dir):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    assert (pm.n_nodes, pm.n_states, pm.n_classes, pm.nnz, pm.diameter) == (49, 61, 20, 526, 3)
    assert_allclose((pm.blen * pm.rate).sum(), 7.2606283938, rtol=1e-9)
    deg = np.diff(pm.q_ptr)
    assert deg.min() == 7 and deg.max() == 9
    with open(os.path.join(golden_dir, 'ref_mg94.json')) as f:
        ref = json.load(f)          # written by the reference's create_mg94.py
    assert [tuple(r) for r in ref['genetic_code']] == [tuple(r) for r in p53.universal_genetic_code()]
    Q = m.get_Q_primary()
    assert len(ref['Q']) == 526
    for a, b, w in ref['Q']:
        assert_allclose(Q[a][b]['weight'], w, rtol=1e-12)
    assert_allclose([m.get_primary_distn()[i] for i in range(61)], ref['distn'], rtol=1e-12)
    assert dict(ref['residue_to_part']) == m.residue_to_part


def test_pack_alignment_equals_object_path():
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, 3, seed=5, reference_leaf=leaf)
    fast = packing.pack_alignment(pm, ls, tt, tf)
    from oracle.refcompat.make_golden import MaskData
    slow = packing.pack_data(pm, [MaskData(pm, ls[i], tt[i], tf[i]) for i in range(3)])
    assert (fast.pri_ok == slow.pri_ok).all()
    assert (fast.tol_true_ok == slow.tol_true_ok).all()
    assert (fast.tol_false_ok == slow.tol_false_ok).all()
    # every site keeps a feasible ON/OFF choice and the observed class is forced ON
    allk = (1 << pm.n_classes) - 1
    assert (((fast.tol_true_ok | fast.tol_false_ok) & allk) == allk).all()
