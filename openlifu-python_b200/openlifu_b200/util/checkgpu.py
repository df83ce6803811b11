"""Is a CUDA device visible?  Same contract as the reference probe (/root/reference/src/openlifu/util/checkgpu.py:
``gpu_available() -> bool``, never raises): the NVML device count, asked without creating a CUDA context so that the
planner can decide before any solver handle exists."""
from __future__ import annotations

import functools


def _nvml_device_count() -> int:
    import pynvml
    pynvml.nvmlInit()
    try:
        return int(pynvml.nvmlDeviceGetCount())
    finally:
        pynvml.nvmlShutdown()


@functools.lru_cache(maxsize=1)
def _count() -> int:
    try:
        return _nvml_device_count()
    except Exception:  # noqa: BLE001 - no driver, no NVML, no permission: all mean "no GPU"
        return 0


def gpu_available() -> bool:
    return _count() > 0
