"""Content identity of numpy arrays (medium cache keys, label-volume provenance)."""
from __future__ import annotations

import zlib

import numpy as np


def content_key(a: np.ndarray):
    """Identity of an array's CONTENT for the medium cache: every byte takes part (an in-place edit of a single voxel
    between two calls changes the key; the reference rebuilds the medium on every call, kwave_if.py:113).  Dense
    arrays are summed as 64-bit words with wrap-around on all host threads (~5 ms for a float64 216^3 map) next to an
    Adler-32 of a strided sample, which is sensitive to position; anything else is hashed byte by byte."""
    a = np.asarray(a)
    flat = None
    if a.flags.c_contiguous or a.flags.f_contiguous:
        flat = a.reshape(-1, order="A")
    if flat is not None and flat.nbytes % 8 == 0 and flat.nbytes > 0:
        words = flat.view(np.int64)
        try:
            import torch
            total = int(torch.from_numpy(words).sum().item())
        except Exception:  # noqa: BLE001 - torch missing or a read-only buffer it refuses
            total = int(words.sum(dtype=np.int64))
        step = max(1, flat.size // 4096)
        return (a.shape, a.dtype.str, total, zlib.adler32(np.ascontiguousarray(flat[::step]).tobytes()))
    return (a.shape, a.dtype.str, zlib.adler32(np.ascontiguousarray(a).tobytes()))


