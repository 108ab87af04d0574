"""ctypes binding of libnxblink_b200.so (include/nxblink_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no CPU fallback: if the shared object is missing or no CUDA device is
usable, everything here raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libnxblink_b200.so')


class NxbError(Exception):
    """Non-zero status from the C-ABI (the reference raises bare ``Exception``)."""

    def __init__(self, code, message):
        Exception.__init__(self, '%s (nxb status %d)' % (message, code))
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [
        ('n_nodes', C.c_int32), ('n_states', C.c_int32),
        ('n_classes', C.c_int32), ('nnz', C.c_int32),
        ('parent', C.POINTER(C.c_int32)),
        ('blen', C.POINTER(C.c_double)), ('rate', C.POINTER(C.c_double)),
        ('q_ptr', C.POINTER(C.c_int32)), ('q_col', C.POINTER(C.c_int32)),
        ('q_val', C.POINTER(C.c_double)),
        ('state_class', C.POINTER(C.c_int32)), ('prior', C.POINTER(C.c_double)),
        ('rate_on', C.c_double), ('rate_off', C.c_double),
        ('diameter', C.c_int32), ('msg_f64', C.c_int32),
        ('cap_scale', C.c_double),
        ('validate', C.c_int32), ('max_warps', C.c_int32),
        ('lockstep_warps', C.c_int32),
        ('split_phases', C.c_int32),
    ]


class SummaryLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'n_counts', 'n_dwell', 'c_root_pri', 'c_root_on', 'c_root_off',
        'c_root_on_cls', 'c_pri_trans', 'c_off_on', 'c_on_off', 'c_nsamples',
        'd_T', 'd_off')]


class Stats(C.Structure):
    _fields_ = [
        ('n_units', C.c_int64), ('state_bytes_used', C.c_int64),
        ('n_pri_events', C.c_int64), ('n_blk_events', C.c_int64),
        ('data_bytes', C.c_int64),
        ('warps_per_cta', C.c_int32), ('n_ctas', C.c_int32),
        ('smem_bytes', C.c_int32), ('n_tiers', C.c_int32),
        ('capP', C.c_int32), ('capB', C.c_int32),
        ('capF', C.c_int32), ('capC', C.c_int32),
        ('deferred_units', C.c_int64), ('kernel_launches', C.c_int64),
        ('capacity_doublings', C.c_int64),
    ]


class MStepDesc(C.Structure):
    _fields_ = [
        ('n_edges', C.c_int32), ('n_states', C.c_int32), ('n_classes', C.c_int32), ('nnz', C.c_int32),
        ('q_ptr', C.c_void_p), ('q_col', C.c_void_p), ('state_class', C.c_void_p),
        ('nt_to', C.c_void_p), ('is_ts', C.c_void_p), ('is_non', C.c_void_p),
        ('nt_counts', C.c_void_p), ('xon_weight', C.c_void_p),
    ]


# every symbol include/nxblink_b200.h declares
EXPORTS = (
    'nxb_device_count', 'nxb_create', 'nxb_destroy', 'nxb_update_model', 'nxb_last_error',
    'nxb_set_data', 'nxb_init_chains', 'nxb_sweep', 'nxb_summary_layout_get',
    'nxb_summary_reset', 'nxb_summary_read', 'nxb_summary_device_ptrs',
    'nxb_summarize_current', 'nxb_export_unit', 'nxb_get_stats',
    'nxb_last_kernel_ms', 'nxb_last_kernel_ms_by_kind', 'nxb_profile', 'nxb_philox4x32_10',
    'nxb_expm_batched', 'nxb_expm_frechet_batched', 'nxb_pruning_loglik', 'nxb_dense_last_error',
    'nxb_forward_sample', 'nxb_forward_sample_rooted', 'nxb_read_node_states',
    'nxb_dense_last_kernel_ms', 'nxb_set_stream',
    'nxb_mstep', 'nxb_mstep_last_error', 'nxb_mstep_last_kernel_ms',
)

_lib = None


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
                'libnxblink_b200.so has not been built; run '
                '`python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.nxb_device_count.restype = C.c_int
    lib.nxb_create.argtypes = [C.POINTER(ModelDesc), C.c_int, C.POINTER(H)]
    lib.nxb_destroy.argtypes = [H]
    lib.nxb_last_error.argtypes = [H]
    lib.nxb_last_error.restype = C.c_char_p
    lib.nxb_set_data.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nxb_init_chains.argtypes = [H, C.c_int32, C.c_uint64, C.c_int64]
    lib.nxb_sweep.argtypes = [H, C.c_int32, C.c_int32]
    lib.nxb_summary_layout_get.argtypes = [H, C.POINTER(SummaryLayout)]
    lib.nxb_summary_reset.argtypes = [H]
    lib.nxb_summary_read.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.nxb_summary_device_ptrs.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.nxb_summarize_current.argtypes = [H]
    lib.nxb_export_unit.argtypes = [H, C.c_int64] + [C.c_void_p] * 2 + [C.c_int32] + [C.c_void_p] * 10
    lib.nxb_get_stats.argtypes = [H, C.POINTER(Stats)]
    lib.nxb_last_kernel_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.nxb_last_kernel_ms_by_kind.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.nxb_profile.argtypes = [H, C.c_int32, C.c_void_p]
    lib.nxb_update_model.argtypes = [H, C.POINTER(ModelDesc)]
    lib.nxb_expm_batched.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nxb_expm_frechet_batched.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nxb_pruning_loglik.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.nxb_dense_last_error.restype = C.c_char_p
    lib.nxb_dense_last_kernel_ms.restype = C.c_float
    lib.nxb_set_stream.argtypes = [H, C.c_uint64]
    lib.nxb_mstep.argtypes = [C.c_int, C.POINTER(MStepDesc), C.POINTER(SummaryLayout), C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    lib.nxb_mstep_last_error.restype = C.c_char_p
    lib.nxb_mstep_last_kernel_ms.restype = C.c_float
    lib.nxb_forward_sample.argtypes = [H, C.c_int64, C.c_uint64, C.c_int64]
    lib.nxb_forward_sample_rooted.argtypes = [H, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.nxb_read_node_states.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.nxb_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nxb_philox4x32_10.restype = None
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib
