"""The C-ABI library loads and exports every symbol include/*.h declares; the
host-side pieces of it that need no GPU behave (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def _declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'include', 'nxblink_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nxb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(nb):
    lib = nb._lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    assert sorted(nb._lib.EXPORTS) == names


def test_philox_known_answers(nb):
    # Random123 known-answer vectors for philox4x32-10
    lib = nb._lib.load()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        out = (C.c_uint32 * 4)()
        lib.nxb_philox4x32_10(c, k, out)
        assert tuple(out) == exp


def test_bad_arguments_are_rejected_before_touching_the_gpu(nb):
    from nxblink_b200 import packing, toymodel
    lib = nb._lib.load()
    pm = packing.PackedModel(toymodel.BlinkModelA)
    h = C.c_void_p()
    d = pm.desc()
    d.rate_on = -1.0
    assert lib.nxb_create(C.byref(d), 0, C.byref(h)) == 5       # NXB_ERR_BAD_ARG
    assert b'positive' in lib.nxb_last_error(None)
    d = pm.desc()
    bad_parent = pm.parent.copy()
    bad_parent[2] = 4
    d.parent = bad_parent.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.nxb_create(C.byref(d), 0, C.byref(h)) == 5
    assert lib.nxb_destroy(None) == 0


def test_no_gpu_fails_loudly(nb):
    """There is no CPU fallback: without a device the product path raises."""
    lib = nb._lib.load()
    if lib.nxb_device_count() > 0:
        pytest.skip('a CUDA device is present')
    from nxblink_b200 import toymodel
    with pytest.raises(nb.NxbError) as err:
        nb.BlinkSampler(toymodel.BlinkModelA)
    assert err.value.code == 6                                   # NXB_ERR_CUDA
    assert 'no CPU fallback' in str(err.value)
