"""Error behaviour through the C-ABI (the reference raises bare ``Exception``; here every entry
point returns a status and the Python layer raises ``NxbError(Exception)`` with the message of
``nxb_last_error``): bad models are refused at creation, calls out of order are refused, and a
failed call leaves the handle usable."""
import ctypes as C

import numpy as np
import pytest

from nxblink_b200 import _lib, packing, toymodel, toydata

pytestmark = pytest.mark.gpu


def _create(pm, mutate):
    lib = _lib.load()
    keep = []                       # keep the mutated arrays alive for the call
    desc = pm.desc()
    mutate(desc, keep)
    h = C.c_void_p()
    rc = lib.nxb_create(C.byref(desc), 0, C.byref(h))
    msg = lib.nxb_last_error(None).decode()
    if rc == 0:
        lib.nxb_destroy(h)
    return rc, msg


def test_bad_models_are_refused(nb):
    pm = packing.PackedModel(toymodel.BlinkModelB)

    def bad_parent(d, keep):
        a = pm.parent.copy(); a[1] = 3; keep.append(a)
        d.parent = a.ctypes.data_as(C.POINTER(C.c_int32))

    def negative_rate(d, keep):
        a = pm.rate.copy(); a[2] = -1.0; keep.append(a)
        d.rate = a.ctypes.data_as(C.POINTER(C.c_double))

    def zero_blink_rate(d, keep):
        d.rate_on = 0.0

    def self_transition(d, keep):
        a = pm.q_col.copy(); a[0] = 0; keep.append(a)
        d.q_col = a.ctypes.data_as(C.POINTER(C.c_int32))

    def class_out_of_range(d, keep):
        a = pm.state_class.copy(); a[0] = 7; keep.append(a)
        d.state_class = a.ctypes.data_as(C.POINTER(C.c_int32))

    def too_many_states(d, keep):
        d.n_states = 65

    for mutate in (bad_parent, negative_rate, zero_blink_rate, self_transition, class_out_of_range,
            too_many_states):
        rc, msg = _create(pm, mutate)
        assert rc == 5 and msg, (mutate.__name__, rc, msg)      # NXB_ERR_BAD_ARG
    rc, msg = _create(pm, lambda d, keep: None)
    assert rc == 0


def test_calls_out_of_order_are_refused_and_the_handle_survives(nb):
    s = nb.BlinkSampler(toymodel.BlinkModelB)
    with pytest.raises(nb.NxbError) as e:
        s.sweep(1)                                  # no chains yet
    assert e.value.code == 5
    with pytest.raises(nb.NxbError):
        s.init_chains(chains=4, seed=0)             # no data yet
    s.set_data([toydata.DataC])
    with pytest.raises(nb.NxbError):
        s.init_chains(chains=0, seed=0)
    s.init_chains(chains=4, seed=0)
    s.sweep(2, accumulate=True)
    assert s.summary().nsamples == 8
    with pytest.raises(nb.NxbError):
        s.export_unit(99)                           # unit out of range
    s.sweep(1, accumulate=True)
    assert s.summary().nsamples == 12
    s.close()


def test_infeasible_data_is_reported_with_the_unit(nb):
    # primary state 0 (class T0) observed at a leaf whose T0 tolerance is observed OFF
    pm = packing.PackedModel(toymodel.BlinkModelB)
    good = packing.pack_data(pm, [toydata.DataC, toydata.DataC])
    pri, tt, tf = good.pri_ok.copy(), good.tol_true_ok.copy(), good.tol_false_ok.copy()
    leaf = pm.node_index['N0']
    pri[1, leaf] = np.uint64(1)                     # only state 0
    tt[1, leaf] = np.uint32(0b110)                  # T0 cannot be on
    s = nb.BlinkSampler(pm)
    s.set_data(packing.PackedData(pri, tt, tf))
    with pytest.raises(nb.NxbError) as e:
        s.init_chains(chains=3, seed=1)
    assert e.value.code == 1 and 'unit' in str(e.value)        # NXB_ERR_INFEASIBLE
    s.close()


def test_mstep_without_samples_is_refused(nb):
    """nxb_mstep on empty statistics (no accumulated sweep) raises instead of dividing by zero."""
    from nxblink_b200 import p53, packing, p53em
    import numpy as np
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    s = nb.BlinkSampler(pm)
    dm = p53em.DeviceMStep(s, m.genetic_code)
    with pytest.raises(Exception) as ei:
        dm.minimize(np.zeros(8 + pm.n_nodes - 1))
    assert 'no samples' in str(ei.value)
    s.close()


def test_new_site_count_drops_the_old_chains(nb):
    """nxb_set_data with another number of sites: the chains belong to the old sites, so a sweep
    without a new nxb_init_chains is refused (round-1 advisor finding: it used to read past the new
    data buffers); init_chains right after forward_sample (no caller data) is refused with a message."""
    M = toymodel.BlinkModelB
    s = nb.BlinkSampler(M)
    s.set_data([toydata.DataC] * 4)
    s.init_chains(chains=8, seed=1)
    s.sweep(2, accumulate=True)
    s.set_data([toydata.DataC] * 2)
    with pytest.raises(Exception) as ei:
        s.sweep(1, accumulate=True)
    assert 'init_chains' in str(ei.value)
    s.init_chains(chains=8, seed=1)
    s.sweep(1, accumulate=False)                 # usable again
    s.forward_sample(64, seed=2)
    with pytest.raises(Exception) as ei:
        s.init_chains(chains=2, seed=3)
    assert 'set_data' in str(ei.value)
    s.close()
