"""fwdsample.gen_potential_transitions (nxblink/fwdsample.py:18-64): the moves out of a joint
state, with the reference's own-class-loss behaviour available behind a flag."""
from nxblink_b200 import fwdsample, toymodel
from nxblink_b200.model import get_Q_blink


def _moves(state, quirk):
    M = toymodel.BlinkModelB
    Qb = get_Q_blink(rate_on=2.0, rate_off=3.0)
    Qb = {0: {1: Qb[False][True]}, 1: {0: Qb[True][False]}}
    return sorted(fwdsample.gen_potential_transitions(state, 'PRIMARY', M.get_primary_to_tol(),
        M.get_Q_primary(), Qb, reference_quirk=quirk), key=repr)


def test_model_semantics():
    # primary 0 (class T0); T1 off blocks 0 -> 2; T2 on
    got = _moves({'PRIMARY': 0, 'T0': 1, 'T1': 0, 'T2': 1}, False)
    assert got == sorted([(('PRIMARY', 0, 1), 2.0), (('T1', 0, 1), 2.0), (('T2', 1, 0), 3.0)], key=repr)


def test_reference_quirk_lets_the_own_class_switch_off():
    got = _moves({'PRIMARY': 0, 'T0': 1, 'T1': 0, 'T2': 1}, True)
    assert (('T0', 1, 0), 3.0) in got and len(got) == 4
