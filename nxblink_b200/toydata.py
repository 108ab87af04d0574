"""Four observation levels for the toy models (the fixtures of nxblink/toydata.py).

DataA: nothing observed; DataB: alignment only; DataC: alignment plus disease
data at the reference leaf N0; DataD: tolerance fully observed at every leaf.
"""
from __future__ import division, print_function

__all__ = ['DataA', 'DataB', 'DataC', 'DataD']

_NODES = ('N0', 'N1', 'N2', 'N3', 'N4', 'N5')
_ANY = (False, True)
_ALL = (0, 1, 2, 3, 4, 5)


def _tol(**overrides):
    """{node: set} with both states feasible except where overridden."""
    return dict((v, set(overrides.get(v, _ANY))) for v in _NODES)


class _Data(object):
    PRIMARY = None
    TOL = None

    @classmethod
    def get_primary_data(cls):
        return dict((v, set(cls.PRIMARY.get(v, _ALL))) for v in _NODES)

    @classmethod
    def get_tolerance_data(cls):
        return dict((name, _tol(**spec)) for name, spec in cls.TOL.items())

    @classmethod
    def get_data(cls):
        data = {'PRIMARY': cls.get_primary_data()}
        data.update(cls.get_tolerance_data())
        return data


_ALIGNMENT = {'N0': (0,), 'N3': (4,), 'N4': (5,), 'N5': (1,)}
_ON, _OFF = (True,), (False,)


class DataA(_Data):
    """No data (toydata.py:23-70)."""
    PRIMARY = {}
    TOL = {'T0': {}, 'T1': {}, 'T2': {}}


class DataB(_Data):
    """Alignment only: an observed state switches its own class on (toydata.py:73-120)."""
    PRIMARY = _ALIGNMENT
    TOL = {'T0': dict(N0=_ON, N5=_ON), 'T1': {}, 'T2': dict(N3=_ON, N4=_ON)}


class DataC(DataB):
    """Alignment and disease data at N0 (toydata.py:123-159)."""
    TOL = {'T0': dict(N0=_ON, N5=_ON), 'T1': dict(N0=_OFF),
           'T2': dict(N0=_ON, N3=_ON, N4=_ON)}


class DataD(DataB):
    """Alignment and fully observed tolerance at the leaves (toydata.py:162-193)."""
    TOL = {'T0': dict(N0=_ON, N3=_ON, N4=_ON, N5=_ON),
           'T1': dict(N0=_OFF, N3=_ON, N4=_ON, N5=_ON),
           'T2': dict(N0=_ON, N3=_ON, N4=_ON, N5=_ON)}
