"""Oracle: k-Wave grid, time axis, PML and k-space operators (TEST INFRASTRUCTURE).

Restates what ``get_kgrid`` (/root/reference/src/openlifu/sim/kwave_if.py:13-27) asks
``kwave.kgrid.kWaveGrid`` for, and what ``kspaceFirstOrder3D`` (called at
kwave_if.py:124-129 with ``pml_auto=True, pml_inside=False``, kwave_if.py:117-122)
derives from it.  k-wave-python 0.4.0 is not vendored in the reference; formulas
follow the k-Wave manual / Treeby & Cox 2010.  Ledger ids (A1..A11) refer to
SURVEY.md section 8c.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------------------- grid
def n_vec(N: int) -> np.ndarray:
    """kWaveGrid.makeDim index vector / N (A2): even N -> (-N/2..N/2-1)/N,
    odd N -> (-(N-1)/2..(N-1)/2)/N, middle entry forced to exactly 0."""
    if N % 2 == 0:
        n = np.arange(-N // 2, N // 2, dtype=np.float64) / N
    else:
        n = np.arange(-(N - 1) // 2, (N - 1) // 2 + 1, dtype=np.float64) / N
    n[N // 2] = 0.0
    return n


def x_vec(N: int, d: float) -> np.ndarray:
    """Grid-point positions (m): x_size * n_vec with x_size = N*d (A2).
    Index i sits at d*(i - floor(N/2))."""
    return (N * d) * n_vec(N)


def k_vec(N: int, d: float) -> np.ndarray:
    """Wavenumbers in *centred* (fftshifted) order: (2*pi/d) * n_vec (A2)."""
    return (2.0 * np.pi / d) * n_vec(N)


def k_vec_fft(N: int, d: float) -> np.ndarray:
    """Wavenumbers in FFT (ifftshifted) order; for even N the Nyquist bin is negative."""
    return np.fft.ifftshift(k_vec(N, d))


# ----------------------------------------------------------------------------- time
def make_time(N, d, c_ref: float = 1500.0, cfl: float = 0.5):
    """kWaveGrid.makeTime(c, cfl) as called at kwave_if.py:22-23 (A1).

    dt = cfl*min(d)/c_max, t_end = |N*d|_2 / c_min, Nt = floor(t_end/dt)+1 (+1 more
    when t_end/dt is not an integer in floating point but rem(t_end, dt)==0).
    ``c_ref`` is the hard-coded 1500 default of get_kgrid (kwave_if.py:13), NOT the medium.
    """
    N = np.asarray(N, dtype=np.float64)
    d = np.asarray(d, dtype=np.float64)
    t_end = float(np.sqrt(np.sum((N * d) ** 2)) / c_ref)
    dt = float(cfl * np.min(d) / c_ref)
    q = t_end / dt
    Nt = int(math.floor(q)) + 1
    if math.floor(q) != math.ceil(q) and math.fmod(t_end, dt) == 0.0:
        Nt += 1
    return Nt, dt


def set_time(t_end: float, dt: float):
    """kwave_if.py:25-26: Nt = round(t_end/dt) (python round), kgrid.setTime(Nt, dt)."""
    return int(round(t_end / dt)), float(dt)


# ----------------------------------------------------------------------------- PML
def largest_prime_factor(n: int) -> int:
    n = int(n)
    best = 1
    p = 2
    while p * p <= n:
        while n % p == 0:
            best = p
            n //= p
        p += 1
    if n > 1:
        best = n
    return best


def optimal_pml_size(N, pml_range=(10, 40)):
    """pml_auto=True (kwave_if.py:118): per axis the PML thickness in [10,40] that
    minimises the largest prime factor of N+2*PML; first minimum wins (A3)."""
    out = []
    for n in N:
        facs = [largest_prime_factor(int(n) + 2 * p) for p in range(pml_range[0], pml_range[1] + 1)]
        out.append(pml_range[0] + int(np.argmin(facs)))
    return tuple(out)


def pml_profile(N: int, d: float, dt: float, c_ref: float, pml_size: int,
                pml_alpha: float = 2.0, staggered: bool = False) -> np.ndarray:
    """k-Wave getPML (A4): exp(-alpha*(c/d)*(x/PML)^4 * dt/2), staggered variant shifted +1/2."""
    x = np.arange(1, pml_size + 1, dtype=np.float64)
    if staggered:
        left = pml_alpha * (c_ref / d) * (((x + 0.5) - pml_size - 1.0) / (0.0 - pml_size)) ** 4
        right = pml_alpha * (c_ref / d) * ((x + 0.5) / pml_size) ** 4
    else:
        left = pml_alpha * (c_ref / d) * ((x - pml_size - 1.0) / (0.0 - pml_size)) ** 4
        right = pml_alpha * (c_ref / d) * (x / pml_size) ** 4
    pml = np.ones(N, dtype=np.float64)
    if pml_size > 0:
        pml[:pml_size] = np.exp(-left * dt / 2.0)
        pml[N - pml_size:] = np.exp(-right * dt / 2.0)
    return pml


# ----------------------------------------------------------------------------- medium
def expand_edge(a: np.ndarray, pml) -> np.ndarray:
    """pml_inside=False: medium maps grow by the PML with edge replication (A11)."""
    a = np.asarray(a)
    if a.ndim == 0:
        return a
    return np.pad(a, [(p, p) for p in pml], mode="edge")


def staggered_density(rho0: np.ndarray, axis: int) -> np.ndarray:
    """rho0 at +d/2 along ``axis`` by linear interpolation; last plane keeps rho0 (A9)."""
    rho0 = np.asarray(rho0, dtype=np.float64)
    out = rho0.copy()
    sl_lo = [slice(None)] * 3
    sl_hi = [slice(None)] * 3
    sl_lo[axis] = slice(0, -1)
    sl_hi[axis] = slice(1, None)
    out[tuple(sl_lo)] = 0.5 * (rho0[tuple(sl_lo)] + rho0[tuple(sl_hi)])
    return out


def db2neper(alpha_db, y: float):
    """dB/(MHz^y cm) -> Np/((rad/s)^y m)."""
    return 100.0 * np.asarray(alpha_db, dtype=np.float64) * (1e-6 / (2.0 * np.pi)) ** y / (20.0 * np.log10(np.e))


# ----------------------------------------------------------------------------- k-space operators
def sinc(x):
    """k-Wave sinc: sin(x)/x (unnormalised), 1 at x==0."""
    x = np.asarray(x, dtype=np.float64)
    out = np.ones_like(x)
    nz = x != 0
    out[nz] = np.sin(x[nz]) / x[nz]
    return out


class KOps:
    """All k-space multipliers on the half spectrum used by rfftn over axes (x,y,z)
    of a C-ordered (Nx,Ny,Nz) array, i.e. the *last* axis (z) is halved.

    (The CUDA path halves x instead -- it stores x fastest; the multipliers are the
    same functions of (kx,ky,kz) so the two are interchangeable.)
    """

    def __init__(self, N, d, dt, c_ref, alpha_power=None):
        Nx, Ny, Nz = (int(v) for v in N)
        dx, dy, dz = (float(v) for v in d)
        kx = k_vec_fft(Nx, dx)
        ky = k_vec_fft(Ny, dy)
        kz = k_vec_fft(Nz, dz)[: Nz // 2 + 1]
        # for even Nz rfft's last bin is the Nyquist bin; k-Wave's sign for it is negative and
        # k_vec_fft already carries that sign at index Nz/2.
        self.kx, self.ky, self.kz = kx, ky, kz
        k = np.sqrt(kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2)
        self.k = k
        self.kappa = sinc(c_ref * k * dt / 2.0)
        self.source_kappa = np.cos(c_ref * k * dt / 2.0)
        self.ddx_pos = (1j * kx * np.exp(1j * kx * dx / 2.0))[:, None, None]
        self.ddy_pos = (1j * ky * np.exp(1j * ky * dy / 2.0))[None, :, None]
        self.ddz_pos = (1j * kz * np.exp(1j * kz * dz / 2.0))[None, None, :]
        self.ddx_neg = (1j * kx * np.exp(-1j * kx * dx / 2.0))[:, None, None]
        self.ddy_neg = (1j * ky * np.exp(-1j * ky * dy / 2.0))[None, :, None]
        self.ddz_neg = (1j * kz * np.exp(-1j * kz * dz / 2.0))[None, None, :]
        if alpha_power is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                n1 = k ** (alpha_power - 2.0)
                n2 = k ** (alpha_power - 1.0)
            n1[~np.isfinite(n1)] = 0.0
            n2[~np.isfinite(n2)] = 0.0
            self.nabla1, self.nabla2 = n1, n2
