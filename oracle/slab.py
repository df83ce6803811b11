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


# ------------------------------------------------------------------------------------------------
# The same decomposition with the exchange done the way the fused passes do it (csrc/fft_wide.cuh, wide.cu): there is no
# pack kernel -- the thread that holds a spectrum value after its line transform stores it straight into the owning rank's
# buffer.  A line of N = A x B points is held by B threads; after the transform thread u < A has X[u + A kb] in register kb.
# With G | B that value's row ky = u + A kb belongs to rank kb // (B / G) at local row u + A (kb % (B / G)); with G | A the
# plane z = t + B i (thread t < B, register i < A, real side) belongs to rank i // (A / G) at local plane t + B (i % (A / G)).
# The functions below route element by element with exactly that arithmetic (no ky // nyl division) and move the blocks with
# gloo instead of NVLink stores.

def routed_forward(H: np.ndarray, world: int, ab_y) -> np.ndarray:
    """H[zl][ky][kx] -> T[z][kyl][kx] through the per-thread routing of kw_y_fwd (Q.G > 0)."""
    A, B = ab_y
    nzl, Ny, Nxh = H.shape
    assert Ny == A * B and B % world == 0
    bg, nyl = B // world, Ny // world
    send = np.zeros((world, nzl, nyl, Nxh), dtype=H.dtype)
    for u in range(A):                       # thread
        for kb in range(B):                  # register
            q, rem = divmod(kb, bg)
            send[q, :, u + A * rem, :] = H[:, u + A * kb, :]
    recv = _all_to_all(send)                 # [src][zl][kyl][kx]: rank src's planes
    return recv.reshape(world * nzl, nyl, Nxh)


def routed_backward(T: np.ndarray, world: int, ab_z, rank: int) -> np.ndarray:
    """T[z][kyl][kx] -> H[zl][ky][kx] through the per-thread routing of kw_z (Q.G > 0): plane blocks by register index,
    this rank's rows land at ky0 + kyl in the owner's buffer."""
    A, B = ab_z
    Nz, nyl, Nxh = T.shape
    assert Nz == A * B and A % world == 0
    ag, nzl = A // world, Nz // world
    send = np.zeros((world, nzl, nyl, Nxh), dtype=T.dtype)
    for t in range(B):                       # thread (real side)
        for i in range(A):                   # register
            q, rem = divmod(i, ag)
            send[q, t + B * rem, :, :] = T[t + B * i, :, :]
    recv = _all_to_all(send)                 # [src][zl][kyl][kx]: rank src's ky rows
    return np.concatenate([recv[q] for q in range(world)], axis=1)
