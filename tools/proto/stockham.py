"""numpy model of the generic Stockham line transform used by csrc/fft_gen.cuh (index logic only)."""
import numpy as np

RADICES = (16, 9, 8, 7, 5, 4, 3, 2)


def factor(n):
    out = []
    for r in RADICES:
        while n % r == 0 and n > 1:
            # prefer leaving a factor >= 4 over a tail of 2: 32 -> 8,4 instead of 16,2
            if r == 16 and (n // 16) == 2:
                break
            out.append(r)
            n //= r
    return out if n == 1 else None


def fft_stockham(x, inverse=False):
    N = x.size
    fac = factor(N)
    assert fac is not None
    tw = np.exp((2j if inverse else -2j) * np.pi * np.arange(N) / N)
    cur = x.astype(np.complex128).copy()
    Ns = 1
    for r in fac:
        nb = N // r
        nxt = np.empty_like(cur)
        step = N // (Ns * r)
        for j in range(nb):
            k = j % Ns
            v = np.array([cur[j + t * nb] * tw[(t * k * step) % N] for t in range(r)])
            # DFT-r
            w = np.exp((2j if inverse else -2j) * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
            v = w @ v
            base = (j - k) * r + k
            for t in range(r):
                nxt[base + t * Ns] = v[t]
        cur = nxt
        Ns *= r
    return cur


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (64, 81, 96, 100, 125, 128, 256, 512, 768, 120, 135, 45, 54, 32, 1000):
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        f = fft_stockham(x)
        b = fft_stockham(f, inverse=True) / N
        print(N, factor(N), np.abs(f - np.fft.fft(x)).max(), np.abs(b - x).max())
