# -*- coding: utf-8 -*-
"""
xicsrt_b200 -- B200-native photon-raytrace hot path behind XICSRT's
``raytrace(config)`` API.

    import xicsrt_b200 as xicsrt
    results = xicsrt.raytrace(config)

The per-iteration pipeline (source ray generation -> optic train -> detector
binning) runs as hand-written FP64 CUDA kernels in ``libxrt.so`` (C ABI in
``include/xrt.h``, built by ``python -m xicsrt_b200.build``).  There is no CPU
fallback: without the library or without a GPU the calls raise.
"""
from .config import get_config, __version__  # noqa: F401


def raytrace(config):
    from . import _driver as _rt
    return _rt.raytrace(config)


def raytrace_single(config, _internal=False):
    from . import _driver as _rt
    return _rt.raytrace_single(config, _internal=_internal)


def raytrace_mp(config, processes=None):
    from . import _driver as _rt
    return _rt.raytrace_mp(config, processes=processes)


def combine_raytrace(input_list, keep_images=True, components=None):
    from . import _driver as _rt
    return _rt.combine_raytrace(input_list, keep_images=keep_images, components=components)
