"""TEST INFRASTRUCTURE (oracle) -- slab-decomposed 3-D real FFT, restated on the CPU.

SURVEY.md 8e row 2 / BASELINE.json config C5: a grid too large for one GPU is split into z slabs
(Nz/G planes per rank); a 3-D transform becomes local 2-D (x, y) transforms, one all-to-all that
trades the z split for a ky split, and local 1-D transforms along z.  The CUDA path
(`csrc/slab.cuh`) does the exchange with peer stores over NVLink / NCCL; here the same block
layout goes through `torch.distributed.all_to_all_single` on gloo so the index logic is checked
on CPU.  There is no reference counterpart: k-Wave's binaries are single-device
(/root/reference/src/openlifu/sim/kwave_if.py:117-129 runs one process on one device).

Layouts (x fastest):  real slab R[zl][y][x];  2-D spectrum H[zl][ky][kx], kx = 0..Nx/2;
transposed T[z][kyl][kx] = block q of the exchange stacked along z.
"""
from __future__ import annotations

import numpy as np


def _all_to_all(send: np.ndarray) -> np.ndarray:
    """send[q] goes to rank q; returns recv with recv[q] = what rank q sent to this rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    assert send.shape[0] == world
    if world == 1:
        return send.copy()
    s = torch.from_numpy(np.ascontiguousarray(send).view(np.float64))
    r = torch.empty_like(s)
    dist.all_to_all_single(r, s)
    return r.numpy().view(send.dtype).reshape(send.shape)


def exchange_forward(H: np.ndarray, world: int) -> np.ndarray:
    """H[zl][ky][kx] (this rank's planes, all ky) -> T[z][kyl][kx] (all planes, this rank's ky rows)."""
    nzl, Ny, Nxh = H.shape
    assert Ny % world == 0
    nyl = Ny // world
    send = np.stack([H[:, q * nyl:(q + 1) * nyl, :] for q in range(world)])      # [q][zl][kyl][kx]
    recv = _all_to_all(send)                                                       # [src][zl][kyl][kx]
    return recv.reshape(world * nzl, nyl, Nxh)


def exchange_backward(T: np.ndarray, world: int) -> np.ndarray:
    """Inverse of exchange_forward."""
    Nz, nyl, Nxh = T.shape
    assert Nz % world == 0
    nzl = Nz // world
    send = T.reshape(world, nzl, nyl, Nxh)                                         # block q = planes of rank q
    recv = _all_to_all(send)                                                       # [src][zl][kyl][kx]
    return np.concatenate([recv[q] for q in range(world)], axis=1)                # ky = src*nyl + kyl


def forward(R: np.ndarray, world: int) -> np.ndarray:
    """Real slab -> this rank's rows of the full 3-D half spectrum, T[z][kyl][kx]."""
    H = np.fft.rfft2(R, axes=(1, 2))
    T = exchange_forward(H, world)
    return np.fft.fft(T, axis=0)


def inverse(T: np.ndarray, world: int, Nx: int) -> np.ndarray:
    """T[z][kyl][kx] -> real slab (normalised like numpy's irfftn)."""
    H = exchange_backward(np.fft.ifft(T, axis=0), world)
    return np.fft.irfft2(H, s=(H.shape[1], Nx), axes=(1, 2))
