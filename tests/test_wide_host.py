"""CPU restatement of the index arithmetic of pipeline "wide" (csrc/fft_wide.cuh): the N = A x B line transform the
kernels run with B threads and one exchange, the row-pair layout of the y passes and the block arithmetic of the routed
stores of a slab decomposition.  numpy only -- the GPU parity of the kernels themselves is tests/test_gpu_wide.py."""
from __future__ import annotations

import numpy as np
import pytest

WIDE = {64: (8, 8), 128: (8, 16), 256: (16, 16), 512: (16, 32), 768: (24, 32), 1024: (32, 32)}


def line_fft(x, A, B, inverse=False):
    """Thread t < B holds x[t + B i] (i < A): DFT-A over i, twiddle w_N^(t ka), exchange, DFT-B over t on threads u < A;
    X[u + A kb] sits in (thread u, register kb).  The inverse runs the same steps backwards (unnormalised)."""
    N = A * B
    sign = 1.0 if inverse else -1.0
    w = lambda n, m: np.exp(sign * 2j * np.pi * m / n)  # noqa: E731
    if not inverse:
        regs = x.reshape(A, B).T                                  # regs[t, i] = x[t + B i]
        Y = np.stack([[sum(regs[t, i] * w(A, i * ka) for i in range(A)) * w(N, t * ka) for ka in range(A)] for t in range(B)])
        X = np.empty(N, dtype=complex)
        for u in range(A):                                        # threads u >= A idle
            for kb in range(B):
                X[u + A * kb] = sum(Y[t, u] * w(B, t * kb) for t in range(B))
        return X
    Y = np.empty((B, A), dtype=complex)
    for u in range(A):
        for t in range(B):
            Y[t, u] = sum(x[u + A * kb] * w(B, t * kb) for kb in range(B)) * w(N, t * u)
    out = np.empty(N, dtype=complex)
    for t in range(B):
        for i in range(A):
            out[t + B * i] = sum(Y[t, ka] * w(A, i * ka) for ka in range(A))
    return out


@pytest.mark.parametrize("N", [64, 128, 768])
def test_two_stage_line_transform_is_the_dft(N):
    A, B = WIDE[N]
    rng = np.random.default_rng(147)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    X = line_fft(x, A, B)
    assert np.allclose(X, np.fft.fft(x), rtol=1e-10, atol=1e-9)
    assert np.allclose(line_fft(X, A, B, inverse=True) / N, x, rtol=1e-10, atol=1e-9)


@pytest.mark.parametrize("N", sorted(WIDE))
def test_row_pair_layout_covers_every_row_once(N):
    """Rows ylo and ylo + B form packed pair m = q B + t (q < A / 2): the two rows of a pair are adjacent registers
    (i = 2q, 2q + 1) of thread t of the y pass (fft_wide.cuh::wpair_rows)."""
    A, B = WIDE[N]
    sh = B.bit_length() - 1
    seen = []
    for m in range(N // 2):
        ylo = ((m >> sh) << (sh + 1)) | (m & (B - 1))
        q, t = divmod(m, B)
        assert ylo == t + B * (2 * q)
        seen += [ylo, ylo + B]
    assert sorted(seen) == list(range(N))


@pytest.mark.parametrize("N", sorted(WIDE))
@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_slab_routing_blocks(N, G):
    """Routed stores of a slab decomposition (kw_y_fwd / kw_z with Q.G > 0): ky = t + A kb (t < A) lies in the rows of rank
    kb // (B / G) at local row t + A (kb % (B / G)); z = t + B i (t < B) in the planes of rank i // (A / G)."""
    A, B = WIDE[N]
    if B % G or A % G:
        pytest.skip("lifusim.cu::slab_wide_ok sends this decomposition through the library-FFT slab path")
    nl = N // G
    bg, ag = B // G, A // G
    for t in range(A):
        for kb in range(B):
            ky = t + A * kb
            assert ky // nl == kb // bg and ky % nl == t + A * (kb % bg)
    for t in range(B):
        for i in range(A):
            z = t + B * i
            assert z // nl == i // ag and z % nl == t + B * (i % ag)


@pytest.mark.parametrize("A,B", sorted(set(WIDE.values())))
def test_rotation_swizzle_of_the_x_pass_exchange(A, B):
    """wline_fft stores Y[t][ka] at t A + rot(ka + t mod A): a bijection onto the N slots, and the 16 lanes of a half warp
    hit a shared-memory bank pair at most twice."""
    rot = lambda c: c - A if c >= A else c  # noqa: E731
    slots = {t * A + rot(ka + (t % A)) for t in range(B) for ka in range(A)}
    assert slots == set(range(A * B))
    for ka in range(A):
        for h in range(0, B, 16):
            banks = [(t * A + rot(ka + (t % A))) % 16 for t in range(h, min(h + 16, B))]
            assert max(banks.count(b) for b in banks) <= 2
