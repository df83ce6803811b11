"""Known-answer tests for the solver oracle (the part of the path whose parity is unpinned by the
reference): they tie the CPU restatement to physics instead of to golden fields.

  * free-space spherical wave: arrival time r/c and 1/r amplitude decay (Green's function)
  * exactness of the k-space scheme in a homogeneous medium: measured phase speed == c0 at CFL 0.5
  * PML: outgoing pulse is absorbed (late-time energy << peak energy)
  * power-law absorption: amplitude ratio follows exp(-alpha * f^y * d)
  * float32 oracle vs float64 oracle: the noise floor quoted next to the 1e-4 GPU tolerance
  * planar interface between two media (heterogeneous c and rho, staggered density): pressure reflection
    (Z2-Z1)/(Z2+Z1) and transmission 2 Z2/(Z1+Z2) of a normally incident plane wave
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import kgrid as kg
from oracle.solver import Assumptions, SolverInputs, simulate


def point_source_run(N=(40, 40, 40), d=1e-3, c0=1500.0, alpha=0.0, f0=250e3, cycles=2, nt=150, cfl=0.3, dtype=np.float64,
                     src=(8, 20, 20), asm=None):
    dt = cfl * d / c0
    t = np.arange(0, cycles / f0, dt)
    sig = np.sin(2 * np.pi * f0 * t) * np.hanning(t.size)
    idx = np.array([src[0] + N[0] * (src[1] + N[1] * src[2])], dtype=np.int64)
    inp = SolverInputs(N=N, d=(d, d, d), dt=dt, Nt=nt, c0=c0, rho0=1000.0, alpha_db=alpha, src_idx=idx, src_p=sig[None, :])
    trace = []
    out = simulate(inp, dtype=dtype, asm=asm, progress=lambda i, p: trace.append(p.copy()), return_p_final=True)
    return out, np.array(trace), dt, sig


def test_spherical_spreading_and_speed():
    out, tr, dt, sig = point_source_run()
    pml = out["pml"]
    s = np.array([8, 20, 20]) + np.array(pml)
    r1, r2 = 10, 20
    a1 = tr[:, s[0] + r1, s[1], s[2]]
    a2 = tr[:, s[0] + r2, s[1], s[2]]
    # arrival time difference from the cross-correlation peak, sub-sample by parabolic fit
    xc = np.correlate(a2, a1, mode="full")
    k = int(np.argmax(xc))
    y0, y1, y2 = xc[k - 1], xc[k], xc[k + 1]
    lag = (k - (len(a1) - 1)) + 0.5 * (y0 - y2) / (y0 - 2 * y1 + y2)
    c_meas = (r2 - r1) * 1e-3 / (lag * dt)
    assert abs(c_meas - 1500.0) / 1500.0 < 5e-3
    # 1/r decay of the peak amplitude
    ratio = np.abs(a1).max() / np.abs(a2).max()
    assert abs(ratio - r2 / r1) / (r2 / r1) < 0.03


def test_pml_absorbs_outgoing_wave():
    out, tr, dt, sig = point_source_run(nt=260)
    pml = out["pml"]
    inner = tuple(slice(pml[a], pml[a] + 40) for a in range(3))
    energy = np.array([np.sum(p[inner] ** 2) for p in tr])
    assert energy[-1] < 1e-5 * energy.max()


def test_power_law_absorption_decay():
    f0, y, alpha_db = 500e3, 0.9, 3.0
    lossless, tr0, dt, _ = point_source_run(alpha=0.0, f0=f0, cycles=4, nt=170, asm=Assumptions(alpha_power=y))
    lossy, tr1, _, _ = point_source_run(alpha=alpha_db, f0=f0, cycles=4, nt=170, asm=Assumptions(alpha_power=y))
    s = np.array([8, 20, 20]) + np.array(lossless["pml"])
    r1, r2 = 8, 24
    def amp(tr, r):
        return np.abs(tr[:, s[0] + r, s[1], s[2]]).max()
    measured = (amp(tr1, r2) / amp(tr1, r1)) / (amp(tr0, r2) / amp(tr0, r1))
    alpha_np_per_m = alpha_db * (f0 / 1e6) ** y * 100.0 / 8.685889638
    expected = np.exp(-alpha_np_per_m * (r2 - r1) * 1e-3)
    assert abs(measured - expected) / expected < 0.03


def test_float32_noise_floor():
    o64, _, _, _ = point_source_run(dtype=np.float64, nt=120)
    o32, _, _, _ = point_source_run(dtype=np.float32, nt=120)
    err = np.linalg.norm(o32["p_max"].astype(np.float64) - o64["p_max"]) / np.linalg.norm(o64["p_max"])
    assert err < 2e-5


def test_source_scaling_and_sign_convention():
    """Additive source: 2 dt/(3 c0 dx) per split component -> first-step pressure at the node equals
    c0^2 * 3 * scale * s[0] (no k-space correction), i.e. 2 c0 dt / dx * s."""
    N, d, c0 = (24, 24, 24), 1e-3, 1500.0
    dt = 0.3 * d / c0
    idx = np.array([12 + 24 * (12 + 24 * 12)], dtype=np.int64)
    sig = np.array([[1.0, 0.0, 0.0]])
    inp = SolverInputs(N=N, d=(d, d, d), dt=dt, Nt=1, c0=c0, rho0=1000.0, alpha_db=0.0, src_idx=idx, src_p=sig)
    out = simulate(inp, dtype=np.float64, asm=Assumptions(source_kspace_correction=False), return_p_final=True)
    pml = out["pml"]
    val = out["p_final"][12 + pml[0], 12 + pml[1], 12 + pml[2]]
    assert np.isclose(val, 2 * c0 * dt / d, rtol=1e-12)


def test_kgrid_conventions():
    assert np.allclose(kg.x_vec(4, 1.0), [-2, -1, 0, 1])
    assert np.allclose(kg.x_vec(5, 1.0), [-2, -1, 0, 1, 2])
    k = kg.k_vec_fft(4, 1.0)
    assert np.isclose(k[2], -np.pi)                       # negative Nyquist
    p = kg.pml_profile(30, 1e-3, 1e-7, 1500.0, 10)
    assert p[10:20].min() == 1.0 and p[0] < p[9] < 1.0 and p[-1] == p[0]
    sg = kg.pml_profile(30, 1e-3, 1e-7, 1500.0, 10, staggered=True)
    assert sg[-1] < p[-1] and sg[0] > p[0]                 # shifted by +1/2 cell
    assert kg.largest_prime_factor(81) == 3 and kg.largest_prime_factor(125) == 5 and kg.largest_prime_factor(97) == 97


def planar_interface_inputs(nxy=16, nz=384, d=0.5e-3, z_src=40, z_int=200, c=(1500.0, 2500.0), rho=(1000.0, 1800.0),
                            f0=500e3, cycles=4, t_end=100e-6, cfl=0.3):
    """Laterally uniform two-layer medium with a plane source: with no lateral PML the periodic solver makes this an
    exactly one-dimensional problem.  Shared with the GPU known-answer test (tests/test_gpu_parity.py)."""
    N = (nxy, nxy, nz)
    dt = cfl * d / max(c)
    c0 = np.full(N, c[0]); c0[:, :, z_int:] = c[1]
    rho0 = np.full(N, rho[0]); rho0[:, :, z_int:] = rho[1]
    t = np.arange(0, cycles / f0, dt)
    sig = np.sin(2 * np.pi * f0 * t) * np.hanning(t.size)
    ix, iy = np.meshgrid(np.arange(nxy), np.arange(nxy), indexing="ij")
    idx = np.sort((ix + nxy * (iy + nxy * z_src)).ravel().astype(np.int64))
    return dict(N=N, d=d, dt=dt, Nt=int(round(t_end / dt)), c0=c0, rho0=rho0, sig=sig, idx=idx,
                Z=(c[0] * rho[0], c[1] * rho[1]))


def test_planar_interface_reflection_and_transmission():
    k = planar_interface_inputs(nxy=4, nz=768, d=0.25e-3, z_src=80, z_int=400)      # 12 / 20 points per wavelength
    inp = SolverInputs(N=k["N"], d=(k["d"],) * 3, dt=k["dt"], Nt=k["Nt"], c0=k["c0"], rho0=k["rho0"], alpha_db=0.0,
                       src_idx=k["idx"], src_p=np.repeat(k["sig"][None, :], k["idx"].size, axis=0))
    trace = []
    pz = 20
    simulate(inp, dtype=np.float64, asm=Assumptions(pml_size=(0, 0, pz)),
             progress=lambda i, p: trace.append(p[1, 2, :].copy()))
    tr = np.array(trace)                                  # (Nt, Nz_expanded): the field only depends on z
    tt = np.arange(tr.shape[0]) * k["dt"]
    a = tr[:, pz + 240]                                    # medium 1, between source and interface
    b = tr[:, pz + 560]                                    # medium 2
    # amplitude ratios from the pulse energies (insensitive to the residual dispersion in the layer with c != c_ref)
    E_i = np.sum(a[tt < 45e-6] ** 2)
    E_r = np.sum(a[tt > 60e-6] ** 2)
    E_t = np.sum(b ** 2)
    Z1, Z2 = k["Z"]
    R, T = (Z2 - Z1) / (Z2 + Z1), 2 * Z2 / (Z1 + Z2)
    assert abs(np.sqrt(E_r / E_i) - R) < 0.01 * R
    assert abs(np.sqrt(E_t / E_i) - T) < 0.01 * T
    assert abs(E_r / E_i + (E_t / E_i) * Z1 / Z2 - 1.0) < 0.01          # energy flux is conserved
    assert abs(np.abs(b).max() / np.abs(a[tt < 45e-6]).max() - T) < 0.01 * T


def lossy_wavenumber(f, c, alpha_db, y, dispersion=True):
    """Root of k-Wave's lossy dispersion relation w^2 = c^2 k^2 (1 + i w tau k^(y-2) - eta k^(y-1)) for a plane wave
    e^{i(kz - wt)}; Im k is the attenuation the scheme should show, w / Re k the phase speed."""
    w = 2 * np.pi * f
    a0 = kg.db2neper(alpha_db, y)
    tau = -2 * a0 * c ** (y - 1)
    eta = 2 * a0 * c ** y * np.tan(np.pi * y / 2) if dispersion else 0.0

    def F(k):
        return w ** 2 - c ** 2 * k ** 2 * (1 + 1j * w * tau * k ** (y - 2) - eta * k ** (y - 1))

    k = w / c + 0j
    for _ in range(100):
        h = 1e-7 * k
        k = k - F(k) * h / (F(k + h) - F(k))
    return k


@pytest.mark.parametrize("dispersion", [True, False])
def test_plane_wave_attenuation_follows_the_lossy_dispersion_relation(dispersion):
    """Plane wave (no lateral PML -> exactly 1-D), alpha = 3 dB/(MHz^0.9 cm) at 500 kHz: the peak amplitude decays with
    Im k of the dispersion relation -- 18.5 Np/m (= alpha f^y) without the dispersion term, 20.6 Np/m with it."""
    d, z_src, f0, nz, nxy, cycles, y, alpha_db = 0.5e-3, 20, 500e3, 216, 4, 8, 0.9, 3.0
    t_end = (nz - z_src) * d / 1500.0 + cycles / f0
    k = planar_interface_inputs(nxy=nxy, nz=nz, d=d, z_src=z_src, z_int=nz, t_end=t_end, cycles=cycles, f0=f0,
                                c=(1500.0, 1500.0), rho=(1000.0, 1000.0))
    inp = SolverInputs(N=k["N"], d=(d,) * 3, dt=k["dt"], Nt=k["Nt"], c0=1500.0, rho0=1000.0, alpha_db=alpha_db,
                       src_idx=k["idx"], src_p=np.repeat(k["sig"][None, :], k["idx"].size, axis=0))
    out = simulate(inp, dtype=np.float64, asm=Assumptions(pml_size=(0, 0, 20), alpha_power=y, absorb_eta=dispersion))
    line = out["p_max"].reshape(nz, nxy, nxy)[:, 1, 2]
    z = np.arange(z_src + int(0.5 * cycles / f0 * 1500 / d) + 4, nz - 30)
    slope = np.polyfit(z * d, np.log(line[z]), 1)[0]
    kz = lossy_wavenumber(f0, 1500.0, alpha_db, y, dispersion)
    assert abs(-slope - kz.imag) < 0.015 * kz.imag
    nominal = alpha_db * (f0 / 1e6) ** y * 100.0 / 8.685889638
    assert abs(kz.imag / nominal - (1.1113 if dispersion else 1.0)) < 2e-3


@pytest.mark.parametrize("medium", ["water", "phantom"])
def test_threaded_backend_is_the_same_time_loop(medium):
    """bench.py times the oracle's step on torch CPU tensors (all host threads); it must be the checker's arithmetic."""
    from tests import cases
    from oracle import scene as osc
    case = cases.v2_small_case(steps=40)
    if medium == "phantom":
        case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
        case["dt"], case["t_end"] = 1.5e-7, 40 * 1.5e-7
    kw = dict(delays=case["delays"], apod=case["apod"], freq=case["freq"], cycles=case["cycles"], dt=case["dt"], t_end=case["t_end"])
    a = osc.run_simulation(cases.scene_of(case), **kw)
    b = osc.run_simulation(cases.scene_of(case), backend="torch", workers=4, **kw)
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(b[k], a[k]) < 5e-6, k
