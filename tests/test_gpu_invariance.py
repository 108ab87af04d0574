"""Determinism and partition invariance of the CUDA sweep.

Random streams are keyed by (seed, global unit, sweep, stream, block), so integer
statistics must be bit-identical regardless of launch geometry, shared-memory
capacity tier (deferral path), message precision of the blink path, or how units
are sharded over handles.  Dwell sums are fp64
atomics: equal up to summation order (rtol 1e-11)."""
import numpy as np
import pytest

from nxblink_b200 import toymodel, toydata, p53, packing

pytestmark = pytest.mark.gpu


def _run(nb, model, data, chains=64, seed=5, sweeps=6, unit_offset=0, **kw):
    s = nb.BlinkSampler(model, **kw)
    s.set_data(data)
    s.init_chains(chains=chains, seed=seed, unit_offset=unit_offset)
    s.sweep(3, accumulate=False)
    s.sweep(sweeps, accumulate=True)
    c, d = s.read_summary()
    st = s.stats()
    s.close()
    return c, d, st


def _p53_inputs(n_sites):
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    leaf = pm.node_index[m.name_to_leaf['Has']]
    ls, tt, tf = p53.synthetic_alignment(pm, n_sites, seed=0, reference_leaf=leaf)
    return pm, packing.pack_alignment(pm, ls, tt, tf)


def test_same_seed_same_result(nb):
    a = _run(nb, toymodel.BlinkModelB, [toydata.DataC] * 4)
    b = _run(nb, toymodel.BlinkModelB, [toydata.DataC] * 4)
    assert (a[0] == b[0]).all()
    np.testing.assert_allclose(a[1], b[1], rtol=1e-11, atol=1e-9)
    c = _run(nb, toymodel.BlinkModelB, [toydata.DataC] * 4, seed=6)
    assert (a[0] != c[0]).any()


def test_capacity_tiers_and_geometry_do_not_change_results(nb):
    pm, data = _p53_inputs(24)
    base = _run(nb, pm, data, chains=8, validate=True)
    assert base[0][-1] == 24 * 8 * 6                     # nsamples
    small = _run(nb, pm, data, chains=8, cap_scale=0.55)  # forces the deferral path
    assert small[2]['deferred_units'] > 0
    few = _run(nb, pm, data, chains=8, max_warps=3)
    # one fused kernel instead of the alternating primary / tolerance kernels, with and
    # without lockstep groups, and the fused kernel under capacity pressure
    fused = _run(nb, pm, data, chains=8, split_phases=-1, validate=True)
    fused_free = _run(nb, pm, data, chains=8, split_phases=-1, lockstep_warps=-1)
    fused_small = _run(nb, pm, data, chains=8, split_phases=-1, cap_scale=0.55)
    assert fused_small[2]['deferred_units'] > 0
    assert fused[2]['kernel_launches'] < base[2]['kernel_launches']
    for other in (small, few, fused, fused_free, fused_small):
        assert (base[0] == other[0]).all()
        np.testing.assert_allclose(base[1], other[1], rtol=1e-11, atol=1e-9)
        assert base[2]['n_blk_events'] == other[2]['n_blk_events']
        assert base[2]['n_pri_events'] == other[2]['n_pri_events']


def test_warp_groups_and_lean_primary_kernel_do_not_change_results(nb, monkeypatch):
    """blink_kernel as independent groups of 5 warps against one group per CTA (NXB_BLINK_GROUPS=0),
    the LEAN primary-only kernel with 24 warps against 20 and 7 (NXB_LEAN_WARPS), all against the
    validated path (non-LEAN kernel, one host wait per sweep): identical integer statistics."""
    pm, data = _p53_inputs(40)
    base = _run(nb, pm, data, chains=16, validate=True)
    dflt = _run(nb, pm, data, chains=16)
    monkeypatch.setenv('NXB_BLINK_GROUPS', '0')
    one_group = _run(nb, pm, data, chains=16)
    monkeypatch.delenv('NXB_BLINK_GROUPS')
    monkeypatch.setenv('NXB_LEAN_WARPS', '20')
    w20 = _run(nb, pm, data, chains=16)
    monkeypatch.setenv('NXB_LEAN_WARPS', '7')
    w7 = _run(nb, pm, data, chains=16)
    for other in (dflt, one_group, w20, w7):
        assert (base[0] == other[0]).all()
        np.testing.assert_allclose(base[1], other[1], rtol=1e-11, atol=1e-9)


def test_sharding_over_handles_is_invisible(nb):
    """Two handles with half the sites each (unit_offset = first global unit) give
    the same integer statistics as one handle with all sites."""
    pm, data = _p53_inputs(16)
    chains = 4
    whole = _run(nb, pm, data, chains=chains)
    lo = packing.PackedData(data.pri_ok[:8], data.tol_true_ok[:8], data.tol_false_ok[:8])
    hi = packing.PackedData(data.pri_ok[8:], data.tol_true_ok[8:], data.tol_false_ok[8:])
    a = _run(nb, pm, lo, chains=chains, unit_offset=0)
    b = _run(nb, pm, hi, chains=chains, unit_offset=8 * chains)
    assert (whole[0] == a[0] + b[0]).all()
    np.testing.assert_allclose(whole[1], a[1] + b[1], rtol=1e-11, atol=1e-9)


def test_conservation_properties_at_scale(nb):
    """Size-independent properties on a large batch: total dwell per edge equals
    branch length x samples; root ON + OFF = K x samples; blink on/off transition
    counts balance against the endpoint states."""
    pm, data = _p53_inputs(512)
    s = nb.BlinkSampler(pm)
    s.set_data(data)
    s.init_chains(chains=32, seed=9)
    s.sweep(2, accumulate=False)
    s.sweep(3, accumulate=True)
    sm = s.summary()
    n = sm.nsamples
    assert n == 512 * 32 * 3
    E, K = pm.n_nodes - 1, pm.n_classes
    np.testing.assert_allclose(sm.flat['T'].sum(axis=1), pm.blen[1:] * n, rtol=1e-10)
    assert sm.root_on_total + sm.root_off_count == K * n
    assert sum(sm.root_pri_to_count.values()) == n
    tot = sm.flat['off_xon_dwell'] + sm.flat['xon_off_dwell']
    np.testing.assert_allclose(tot, (K - 1) * pm.blen[1:] * n, rtol=1e-10)
    assert (sm.flat['Doff'] >= 0).all() and (sm.flat['pri_dwell'] >= -1e-9).all()
    st = s.stats()
    R = (st['n_pri_events'] + st['n_blk_events']) / st['n_units']
    assert 60 < R < 160          # ~84-121 retained events at the p53 shape
    s.close()


def test_invariants_hold_over_a_million_unit_sweeps(nb):
    """Every trajectory invariant (sorted strictly increasing event times, compatible
    transitions, own class ON, data constraints; navigation.py:39-44,121-135) checked after
    every phase of 1.3 million unit-sweeps at the p53 shape."""
    pm, data = _p53_inputs(4096)
    s = nb.BlinkSampler(pm, validate=True)
    s.set_data(data)
    s.init_chains(chains=16, seed=3)
    s.sweep(20, accumulate=True)
    assert s.summary().nsamples == 4096 * 16 * 20
    s.close()


def test_hard_constraints_survive_a_long_soak(nb):
    """1.6e8 validated unit-sweeps on the toy grid with tolerance data.  Regression: the message
    ratio kept per live tolerance event once came from an approximate division, so that l1 / l1
    could be 1 - 2^-24 and the downward pass switched the own class of the primary state off once
    in ~3e8 unit-sweeps (found by scripts/gpu_exact_long.py).  Hard constraints must be exact."""
    for seed in range(16):
        s = nb.BlinkSampler(toymodel.BlinkModelB, validate=True)
        s.set_data([toydata.DataC])
        s.init_chains(chains=65536, seed=7000 + seed)
        s.sweep(150, accumulate=False)          # raises on the first violated invariant
        s.close()


def test_event_parallel_kernel_geometry_invariance(nb, monkeypatch):
    """NXB_TOLPAR=1 (tolpar_kernel): its random streams are keyed by (seed, global unit, sweep,
    slice | retained index | track), so one or two warps per unit, the capacity tier (deferral) and
    sharding leave the integer statistics bit-identical.  (It is a different exact sampler from
    the lane-per-track walk, so the two are compared statistically, not bitwise:
    tests/test_gpu_precision.py.)"""
    pm, data = _p53_inputs(24)
    monkeypatch.setenv('NXB_TOLPAR', '1')
    monkeypatch.setenv('NXB_TP_WPU', '2')
    base = _run(nb, pm, data, chains=8, validate=True)
    assert base[0][-1] == 24 * 8 * 6
    small = _run(nb, pm, data, chains=8, cap_scale=0.55)
    assert small[2]['deferred_units'] > 0
    few = _run(nb, pm, data, chains=8, max_warps=4)
    monkeypatch.setenv('NXB_TP_WPU', '1')
    one = _run(nb, pm, data, chains=8)
    lo = packing.PackedData(data.pri_ok[:8], data.tol_true_ok[:8], data.tol_false_ok[:8])
    hi = packing.PackedData(data.pri_ok[8:], data.tol_true_ok[8:], data.tol_false_ok[8:])
    a = _run(nb, pm, lo, chains=8, unit_offset=0)
    b = _run(nb, pm, hi, chains=8, unit_offset=8 * 8)
    for other in (small, few, one, (a[0] + b[0], a[1] + b[1], None)):
        assert (base[0] == other[0]).all()
        np.testing.assert_allclose(base[1], other[1], rtol=1e-11, atol=1e-9)


@pytest.mark.parametrize('kw', [{}, {'split_phases': -1}, {'validate': True}], ids=['split', 'fused', 'validated'])
@pytest.mark.parametrize('cap_scale', [0.06, 0.02])
def test_slab_overflow_grows_the_slabs_and_continues_exactly(nb, cap_scale, kw):
    """HBM slabs far too small for the trajectories: while the chains move from their initial
    trajectories to the stationary event load, units outgrow their slabs (once at cap_scale 0.06,
    three times at 0.02).  The engine doubles every capacity, moves the unit states to the larger
    slabs and runs the units that were left behind again from where they stood, skipping the
    statistics the failed attempt had booked -- here the statistics are accumulated from the very
    first sweep, so that path is exercised.  The result is bit-identical (integers) to the run
    whose slabs were large enough from the start."""
    pm, data = _p53_inputs(24)

    def run(**opts):
        s = nb.BlinkSampler(pm, **opts)
        s.set_data(data)
        s.init_chains(chains=8, seed=5)
        d0 = s.stats()['capacity_doublings']
        s.sweep(9, accumulate=True)
        c, d = s.read_summary()
        st = s.stats()
        s.close()
        return c, d, st, d0
    base = run()
    assert base[2]['capacity_doublings'] == 0
    tiny = run(cap_scale=cap_scale, **kw)
    assert tiny[2]['capacity_doublings'] > tiny[3]          # grown during the accumulating sweeps
    assert (base[0] == tiny[0]).all()
    np.testing.assert_allclose(base[1], tiny[1], rtol=1e-11, atol=1e-9)
    assert base[2]['n_blk_events'] == tiny[2]['n_blk_events']
    assert base[2]['n_pri_events'] == tiny[2]['n_pri_events']


def test_update_model_with_much_higher_rates_resizes_the_handle(nb):
    """An EM step may multiply the rates (run-stochastic.py:154-208 builds a new model per
    iteration): nxb_update_model derives the capacities again, so the handle created for the slow
    model gives exactly what a handle created for the fast model gives."""
    m = p53.p53_model()
    pm = packing.PackedModel(m)
    prm = p53.jeff_params()
    fast_rates = dict((e, 5.0 * r) for e, r in m.get_edge_to_rate().items())
    m5 = p53.Model(prm['kappa'], prm['omega'], prm['A'], prm['C'], prm['G'], prm['T'], 3.0, 1.0,
            m._tree, m._root, m.get_edge_to_blen(), fast_rates, m.genetic_code)
    pm5 = packing.PackedModel(m5)
    _, data = _p53_inputs(8)

    def run(s):
        s.set_data(data)
        s.init_chains(chains=4, seed=21)
        s.sweep(2, accumulate=False)
        s.reset_summary()
        s.sweep(3, accumulate=True)
        return s.read_summary()
    fresh = nb.BlinkSampler(pm5)
    c0, d0 = run(fresh)
    fresh.close()
    s = nb.BlinkSampler(pm)
    run(s)
    s.update_model(pm5)
    c1, d1 = run(s)
    assert (c0 == c1).all()
    np.testing.assert_allclose(d0, d1, rtol=1e-11, atol=1e-9)
    s.close()
