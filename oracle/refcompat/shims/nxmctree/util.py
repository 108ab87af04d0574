"""Stand-ins for ``nxmctree.util.{prod, dict_distn}`` (examples/p53/create_mg94.py:52)."""
import operator
from functools import reduce


def prod(seq):
    return reduce(operator.mul, seq, 1)


def dict_distn(weights):
    total = sum(weights.values())
    return dict((k, v / total) for k, v in weights.items())
