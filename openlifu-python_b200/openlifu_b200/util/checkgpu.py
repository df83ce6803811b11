"""Is a CUDA device visible?  Same contract as the reference probe (/root/reference/src/openlifu/util/checkgpu.py:
``gpu_available() -> bool``, never raises): the NVML device count, asked without creating a CUDA context so that the
planner can decide before any solver handle exists.  Where NVML cannot answer (pynvml is not installed in every image)
the solver library itself is asked (``lifu_device_count``), i.e. the component that will do the work."""
from __future__ import annotations

import functools
import logging

log = logging.getLogger(__name__)


def _nvml_device_count() -> int:
    import pynvml
    pynvml.nvmlInit()
    try:
        return int(pynvml.nvmlDeviceGetCount())
    finally:
        pynvml.nvmlShutdown()


def _library_device_count() -> int:
    from .. import _lib
    return _lib.device_count()


@functools.lru_cache(maxsize=1)
def _count() -> int:
    try:
        return _nvml_device_count()
    except Exception as e:  # noqa: BLE001 - no pynvml / no driver / no permission: ask the solver library
        log.debug("NVML probe failed (%s); asking liblifusim", e)
    try:
        return _library_device_count()
    except Exception as e:  # noqa: BLE001 - library not built: nothing can run the solver here
        log.debug("liblifusim probe failed (%s)", e)
        return 0


def gpu_available() -> bool:
    return _count() > 0
