"""GPU probe (mirrors /root/reference/src/openlifu/util/checkgpu.py:6-14, NVML device count)."""
from __future__ import annotations


def gpu_available() -> bool:
    try:
        from pynvml import nvmlDeviceGetCount, nvmlInit, nvmlShutdown
        nvmlInit()
        n = nvmlDeviceGetCount()
        nvmlShutdown()
        return n > 0
    except Exception:  # noqa: BLE001 - driver problems mean "no GPU"
        return False
