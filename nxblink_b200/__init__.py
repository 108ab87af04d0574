"""nxblink_b200 -- B200-native Rao-Teh sweep for nxblink's blinking-tolerance CTBN.

Keeps the reference's call surface for the hot path
(``raoteh.gen_samples``, ``summary.Summary``, ``em.get_ll_*``,
``maxlikelihood.get_blink_rate_mle``) and adds the batched ``BlinkSampler``.
The compute runs in hand-written sm_100a CUDA behind include/nxblink_b200.h.
"""
from . import model, em, maxlikelihood, summary, raoteh, toymodel, toydata  # noqa: F401
from .sampler import BlinkSampler, NxbError  # noqa: F401
from .packing import PackedModel, PackedData, pack_data, pack_alignment  # noqa: F401

__all__ = ['BlinkSampler', 'NxbError', 'PackedModel', 'PackedData',
        'pack_data', 'pack_alignment', 'model', 'em', 'maxlikelihood',
        'summary', 'raoteh', 'toymodel', 'toydata']
