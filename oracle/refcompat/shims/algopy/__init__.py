"""Stand-in for ``algopy`` with plain numpy values (no derivative tracking).

TEST INFRASTRUCTURE ONLY.  The reference's ``nxblink/em.py`` is written against
algopy so that AD types can flow through (em.py:9-10,16,52,75,...); for *values*
numpy is a drop-in.  Only the symbols em.py touches are provided.
"""
import numpy as _np

log = _np.log
exp = _np.exp
square = _np.square
dot = _np.dot


def zeros(shape, dtype=None):
    # algopy uses ``dtype=<array>`` to mean "same exotic type as this array".
    return _np.zeros(shape, dtype=float)


def ones(shape, dtype=None):
    return _np.ones(shape, dtype=float)


def zeros_like(a):
    return _np.zeros_like(_np.asarray(a, dtype=float))
