"""Maximum likelihood estimation of the blink rates from six statistics.

Mirror of nxblink/maxlikelihood.py:19-163: same function names, argument order
and optimiser settings (L-BFGS-B from (1, 1), lower bounds ``low_rate``).
Golden vectors: nxblink/tests/test_maxlikelihood.py:12-17,42-47.
"""
from __future__ import division, print_function

import numpy as np


def ll_blink_contribution(xon_root_count, off_root_count, off_xon_count,
        xon_off_count, off_xon_dwell, xon_off_dwell, rate_on, rate_off):
    """maxlikelihood.py:19-42."""
    root = (xon_root_count * np.log(rate_on) + off_root_count * np.log(rate_off)
            - (xon_root_count + off_root_count) * np.log(rate_on + rate_off))
    trans = off_xon_count * np.log(rate_on) + xon_off_count * np.log(rate_off)
    dwell = -off_xon_dwell * rate_on - xon_off_dwell * rate_off
    return root + trans + dwell


def ll_blink_contribution_gradient(xon_root_count, off_root_count, off_xon_count,
        xon_off_count, off_xon_dwell, xon_off_dwell, rate_on, rate_off):
    """maxlikelihood.py:45-88."""
    total = rate_on + rate_off
    n_root = xon_root_count + off_root_count
    d_on = (xon_root_count + off_xon_count) / rate_on - n_root / total - off_xon_dwell
    d_off = (off_root_count + xon_off_count) / rate_off - n_root / total - xon_off_dwell
    return d_on, d_off


def get_blink_rate_analytical_mle(xon_root_count, off_root_count, off_xon_count,
        xon_off_count, off_xon_dwell, xon_off_dwell):
    """Closed form of maxlikelihood.py:91-119."""
    u = -(xon_root_count + off_root_count)

    def root_of(a, b, h, f):
        z = a * f - 2 * a * h - b * h + f * u - h * u
        disc = z * z - 4 * a * h * (h - f) * (a + b + u)
        return (np.sqrt(disc) + z) / (2 * h * (f - h))
    a = xon_root_count + off_xon_count
    b = off_root_count + xon_off_count
    return (root_of(a, b, off_xon_dwell, xon_off_dwell),
            root_of(b, a, xon_off_dwell, off_xon_dwell))


def get_blink_rate_mle(xon_root_count, off_root_count, off_xon_count,
        xon_off_count, off_xon_dwell, xon_off_dwell, low_rate=1e-4):
    """maxlikelihood.py:122-163."""
    import scipy.optimize
    stats = (xon_root_count, off_root_count, off_xon_count, xon_off_count,
            off_xon_dwell, xon_off_dwell)

    def neg_ll(rates):
        return -ll_blink_contribution(*(stats + tuple(rates)))

    def fprime(rates):
        return -np.array(ll_blink_contribution_gradient(*(stats + tuple(rates))))
    result = scipy.optimize.fmin_l_bfgs_b(neg_ll, np.array([1.0, 1.0]), fprime,
            bounds=((low_rate, None), (low_rate, None)))
    return result[0][0], result[0][1]
