"""ghm_b200 -- B200-native JGHM sampler + exact belief propagation.

``ghm_b200.ops``               torch-tensor wrappers over the C ABI (include/ghm_b200.h)
``ghm_b200.data_random_GHM``   call-compatible mirror of the reference module
                               ``ghmclip.data.data_random_GHM`` (GHMTree, *Sampler, PPCLIPLoss ...)
``ghm_b200.sweeps``            device-resident p_flip / sigma sweeps, reference-layout JSON + checkpoint writers
``ghm_b200.feed``              asynchronous training-loop feed (BatchPrefetcher over get_batch(async_=True))
"""
from . import _lib  # noqa: F401

__all__ = ["ops", "data_random_GHM", "sweeps", "feed"]
