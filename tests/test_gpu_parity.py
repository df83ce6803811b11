"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): source index masks / integer geometry bit-exact;
p_max / p_min within 1e-4 relative L2 (float32 CUDA vs float32 oracle).
"""
from __future__ import annotations

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _check_fields(got, want, tol=TOL):
    for k in ("p_max", "p_min"):
        err = cases.rel_l2(got[k], want[k])
        assert err < tol, f"{k}: rel-L2 {err:.3e} >= {tol}"


def test_small_water_bli_and_fields(lifu_lib):
    case = cases.small_water_case()
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert got["Nt"] == want["Nt"]
    assert tuple(got["stats"]["pml"]) == tuple(want["pml"])
    assert tuple(got["stats"]["n_exp"]) == tuple(want["N_exp"])
    assert np.array_equal(got["n_delay"], want["n_delay"])
    assert np.array_equal(got["src_idx"], want["src_idx"])          # bit-exact source mask
    Wg = cases.csr_to_dense(got["src_idx"].size, len(case["pos_m"]), got["row_ptr"], got["col"], got["w"])
    assert np.allclose(Wg, want["W"], rtol=2e-6, atol=1e-7)
    _check_fields(got, want)
    assert got["stats"]["kernel_launches"] > 0


def test_oracle_geometry_through_abi(lifu_lib):
    """lifu_set_source_geometry (explicit CSR upload) gives the same fields as the GPU-built one."""
    case = cases.small_water_case()
    want = cases.run_oracle_case(case)
    W = want["W"]
    rows, cols = np.nonzero(W)
    row_ptr = np.zeros(W.shape[0] + 1, dtype=np.int32)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    got = cases.run_cuda_case(case, geometry=(want["src_idx"], row_ptr, cols.astype(np.int32), W[rows, cols], W.shape[1]))
    _check_fields(got, want)


def test_no_kspace_source_correction(lifu_lib):
    from oracle.solver import Assumptions
    case = cases.small_water_case()
    want = cases.run_oracle_case(case, asm=Assumptions(source_kspace_correction=False))
    got = cases.run_cuda_case(case, source_mode="additive-no-correction")
    _check_fields(got, want)


def test_homogeneous_absorbing(lifu_lib):
    case = cases.small_water_case()
    case["alpha"] = 0.75
    case["c0"] = 1540.0
    case["rho0"] = 1050.0
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert got["stats"]["absorbing"] == 1
    _check_fields(got, want)


def _phantom(N):
    """Layered water / skull-like slab / tissue phantom with per-material c, rho, alpha."""
    c0 = np.full(N, 1500.0)
    rho0 = np.full(N, 1000.0)
    al = np.full(N, 0.0022)
    z = np.arange(N[2])[None, None, :] + 0 * np.arange(N[0])[:, None, None]
    x = np.arange(N[0])[:, None, None]
    slab = (z + (x // 6) >= 12) & (z + (x // 6) < 16)
    tissue = (z + (x // 6)) >= 16
    slab = np.broadcast_to(slab, N)
    tissue = np.broadcast_to(tissue, N)
    c0[slab], rho0[slab], al[slab] = 2800.0, 1900.0, 6.0
    c0[tissue], rho0[tissue], al[tissue] = 1540.0, 1050.0, 0.3
    return c0, rho0, al


@pytest.mark.parametrize("alpha_mode", ["binary", "no_dispersion"])
def test_heterogeneous_absorbing(lifu_lib, alpha_mode):
    from oracle.solver import Assumptions
    case = cases.small_water_case()
    case["c0"], case["rho0"], case["alpha"] = _phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 80 * 1.5e-7
    want = cases.run_oracle_case(case, asm=Assumptions(absorb_eta=alpha_mode == "binary"))
    got = cases.run_cuda_case(case, alpha_mode=alpha_mode)
    assert got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    _check_fields(got, want)


def test_tilted_elements_mask_bit_exact(lifu_lib):
    """Rotated, off-grid elements on an even-sized grid (half-voxel origin quirk, App. B 8)."""
    pos = np.array([[-4.3, 1.1, 0.7], [3.9, -2.2, 1.4], [0.2, 5.1, -0.3]])
    size = np.array([[2.3, 3.1], [2.0, 2.0], [3.3, 1.7]])
    ang = np.array([[0.0, 14.17, 0.0], [-9.0, 0.0, 0.0], [5.0, -7.0, 30.0]])
    case = cases.make_case([(-12, 11.5), (-10, 10), (-3, 20.5)], 0.5, 0, 0, 0, 0, (0, 0, 12), 500e3, 2,
                           elem_pos_mm=pos, elem_size_mm=size, angles_deg=ang, dt=1.2e-7, t_end=50 * 1.2e-7)
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert np.array_equal(got["src_idx"], want["src_idx"])
    Wg = cases.csr_to_dense(got["src_idx"].size, 3, got["row_ptr"], got["col"], got["w"])
    assert np.allclose(Wg, want["W"], rtol=2e-6, atol=1e-7)
    _check_fields(got, want)


def test_c1_full(lifu_lib):
    """SURVEY.md config C1 end to end (81x81x125 expanded grid, 229 steps)."""
    case = cases.c1_case()
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert got["Nt"] == 229 and tuple(got["stats"]["n_exp"]) == (81, 81, 125)
    assert np.array_equal(got["src_idx"], want["src_idx"])
    _check_fields(got, want)


def test_repeatable(lifu_lib):
    case = cases.small_water_case()
    a = cases.run_cuda_case(case)
    b = cases.run_cuda_case(case)
    assert np.array_equal(a["p_max"], b["p_max"]) and np.array_equal(a["p_min"], b["p_min"])


def test_errors(lifu_lib):
    from openlifu_b200 import _lib
    with pytest.raises(ValueError):
        _lib.LifuSim([0, 4, 4], [1e-3] * 3, 1e-7, 10)
    with _lib.LifuSim([16, 16, 16], [1e-3] * 3, 1e-7, 4) as sim:
        with pytest.raises(_lib.LifuError, match="lifu_set_medium first"):
            sim.run()


# ---------------------------------------------------------------------------------------------
# pipeline v2 (hand-written fused FFT passes) -- same oracle, same tolerances
def _is_v2(got):
    return got["stats"]["fft_launches"] == 0


def test_v2_small_water(lifu_lib):
    case = cases.v2_small_case()
    want = cases.run_oracle_case(case)
    assert tuple(want["N_exp"]) == (64, 64, 64)
    got = cases.run_cuda_case(case)
    assert _is_v2(got), "auto selection should pick the fused pipeline on a 64^3 lossless grid"
    assert np.array_equal(got["src_idx"], want["src_idx"])
    _check_fields(got, want)


def test_v2_matches_v1(lifu_lib):
    case = cases.v2_small_case()
    a = cases.run_cuda_case(case, pipeline="v1", fields=(0, 1, 2, 3, 4, 5, 6))
    b = cases.run_cuda_case(case, pipeline="v2", fields=(0, 1, 2, 3, 4, 5, 6))
    assert not _is_v2(a) and _is_v2(b)
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(b[k], a[k]) < 2e-5
    for f in range(7):
        assert cases.rel_l2(b[f"field{f}"], a[f"field{f}"]) < 2e-4, f"state field {f} differs between pipelines"


def test_v2_no_source_correction(lifu_lib):
    from oracle.solver import Assumptions
    case = cases.v2_small_case()
    want = cases.run_oracle_case(case, asm=Assumptions(source_kspace_correction=False))
    got = cases.run_cuda_case(case, source_mode="additive-no-correction")
    assert _is_v2(got)
    _check_fields(got, want)


def test_v2_heterogeneous_lossless(lifu_lib):
    case = cases.v2_small_case()
    c0, rho0, _ = _phantom(tuple(case["N"]))
    case["c0"], case["rho0"], case["alpha"] = c0, rho0, 0.0
    case["dt"], case["t_end"] = 1.5e-7, 100 * 1.5e-7
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert _is_v2(got) and got["stats"]["homogeneous"] == 0
    _check_fields(got, want)


def test_v2_mixed_axes_tilted(lifu_lib):
    """256 x 64 x 64 expanded grid (different radix per axis), rotated off-grid elements."""
    pos = np.array([[-20.3, 1.1, 0.7], [13.9, -2.2, 1.4], [40.2, 3.1, -0.3]])
    size = np.array([[2.3, 3.1], [2.0, 2.0], [3.3, 1.7]])
    ang = np.array([[0.0, 14.17, 0.0], [-9.0, 0.0, 0.0], [5.0, -7.0, 30.0]])
    case = cases.make_case([(-54, 53.5), (-10, 9.5), (-3, 18.5)], 0.5, 0, 0, 0, 0, (0, 0, 12), 500e3, 2,
                           elem_pos_mm=pos, elem_size_mm=size, angles_deg=ang, dt=1.2e-7, t_end=60 * 1.2e-7)
    want = cases.run_oracle_case(case)
    assert tuple(want["N_exp"]) == (256, 64, 64)
    got = cases.run_cuda_case(case)
    assert _is_v2(got)
    assert np.array_equal(got["src_idx"], want["src_idx"])
    _check_fields(got, want)


def test_v2_c2_grid_short(lifu_lib):
    """The headline grid (216^3 -> 256^3, 2x64-element array) for 24 time steps against the oracle."""
    from openlifu_b200 import configs
    arr = configs.openlifu_2x_array()
    half = 53.75
    pos = np.array([el.position for el in arr.elements])
    size = np.array([el.size for el in arr.elements])
    ang = np.array([el.get_angle(units="deg") for el in arr.elements])
    dt = 0.5 * 0.5e-3 / 1500
    case = cases.make_case([(-half, half), (-half, half), (-4, 103.5)], 0.5, 0, 0, 0, 0, (0, 0, 50), 400e3, 20,
                           elem_pos_mm=pos, elem_size_mm=size, angles_deg=ang, sensitivity=None, dt=dt, t_end=24 * dt)
    want = cases.run_oracle_case(case)
    assert tuple(want["N_exp"]) == (256, 256, 256)
    got = cases.run_cuda_case(case)
    assert _is_v2(got)
    assert np.array_equal(got["src_idx"], want["src_idx"])
    _check_fields(got, want)


def test_v2_homogeneous_absorbing(lifu_lib):
    """Fused pipeline with the power-law absorption terms (two extra transform round trips per step)."""
    case = cases.v2_small_case()
    case["alpha"], case["c0"], case["rho0"] = 0.75, 1540.0, 1050.0
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case, fields=(0,))
    assert _is_v2(got) and got["stats"]["absorbing"] == 1
    _check_fields(got, want)
    v1 = cases.run_cuda_case(case, pipeline="v1", fields=(0,))
    assert not _is_v2(v1)
    assert cases.rel_l2(got["field0"], v1["field0"]) < 2e-4          # final pressure field, both pipelines


@pytest.mark.parametrize("alpha_mode", ["binary", "no_dispersion", "no_absorption"])
def test_v2_heterogeneous_absorbing(lifu_lib, alpha_mode):
    from oracle.solver import Assumptions
    case = cases.v2_small_case()
    case["c0"], case["rho0"], case["alpha"] = _phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 100 * 1.5e-7
    asm = Assumptions(absorb_eta=alpha_mode != "no_dispersion", absorb_tau=alpha_mode != "no_absorption")
    want = cases.run_oracle_case(case, asm=asm)
    got = cases.run_cuda_case(case, alpha_mode=alpha_mode)
    assert _is_v2(got) and got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    _check_fields(got, want)
    v1 = cases.run_cuda_case(case, alpha_mode=alpha_mode, pipeline="v1")
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], v1[k]) < 2e-5


from tests.config_cases import c2_case as _c2_case, c3_maps  # noqa: E402


def test_c2_full_size_properties(lifu_lib):
    """BASELINE config C2 at full size (256^3, all 749 steps), size-independent properties:
    linearity (drive x2 -> fields x2, bit-exact: every operation of the step is linear and scaling by two is exact),
    run-to-run determinism, agreement of the two independent FFT implementations (fused passes vs cuFFT) over the
    whole run, and the focus where the geometry puts it."""
    a = cases.run_cuda_case(_c2_case())
    assert _is_v2(a) and a["stats"]["steps"] == 749 and tuple(a["stats"]["n_exp"]) == (256, 256, 256)
    b = cases.run_cuda_case(_c2_case(amplitude=2.0))
    for k in ("p_max", "p_min"):
        assert np.array_equal(b[k], 2.0 * a[k]), k
    again = cases.run_cuda_case(_c2_case())
    assert np.array_equal(again["p_max"], a["p_max"]) and np.array_equal(again["p_min"], a["p_min"])
    v1 = cases.run_cuda_case(_c2_case(), pipeline="v1")
    assert not _is_v2(v1)
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(a[k], v1[k]) < 2e-5, k
    pm = a["p_max"].reshape(216, 216, 216, order="F")
    ix, iy, iz = np.unravel_index(np.argmax(pm[:, :, 40:]), pm[:, :, 40:].shape)
    x = -53.75 + 0.5 * ix
    y = -53.75 + 0.5 * iy
    z = -4.0 + 0.5 * (iz + 40)
    assert abs(x) <= 1.0 and abs(y) <= 1.0 and 40.0 <= z <= 52.0, (x, y, z)   # focal peak just short of the 50 mm geometric focus


def test_c3_full_size_linearity_and_pipelines(lifu_lib):
    """C3 (skull / brain phantom, absorbing) at full size: linearity bit-exact, fused pipeline == cuFFT pipeline."""
    c0, rho0, al = c3_maps()
    a = cases.run_cuda_case(_c2_case(steps=300, c0=c0, rho0=rho0, alpha=al))
    assert _is_v2(a) and a["stats"]["absorbing"] == 1 and a["stats"]["homogeneous"] == 0
    b = cases.run_cuda_case(_c2_case(steps=300, c0=c0, rho0=rho0, alpha=al, amplitude=0.5))
    for k in ("p_max", "p_min"):
        assert np.array_equal(b[k], 0.5 * a[k]), k
    v1 = cases.run_cuda_case(_c2_case(steps=300, c0=c0, rho0=rho0, alpha=al), pipeline="v1")
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(a[k], v1[k]) < 5e-5, k


@pytest.mark.parametrize("pipeline,nxy,nz,d,z_src,z_int,t_end", [
    ("v2", 64, 216, 0.5e-3, 20, 110, 60e-6),        # 64 x 64 x 256 with the z PML: fused-FFT pipeline
    ("v1", 8, 728, 0.25e-3, 60, 380, 100e-6),       # 8 x 8 x 768: cuFFT pipeline, 12 / 20 points per wavelength
])
def test_planar_interface_transmission_known_answer(lifu_lib, pipeline, nxy, nz, d, z_src, z_int, t_end):
    """Oracle-independent anchor of the CUDA path: a plane wave normally incident on a planar interface between two
    media (c, rho = 1500, 1000 | 2500, 1800) is transmitted with pressure amplitude 2 Z2 / (Z1 + Z2) = 1.5.  No lateral
    PML: the periodic solver makes the problem exactly one-dimensional.  Exercises heterogeneous sound speed, the
    staggered density, explicit source geometry and the float64 map import."""
    import os
    from openlifu_b200 import _lib
    from tests.test_oracle_physics import planar_interface_inputs
    k = planar_interface_inputs(nxy=nxy, nz=nz, d=d, z_src=z_src, z_int=z_int, t_end=t_end)
    n_src = k["idx"].size
    os.environ["LIFU_PIPELINE"] = pipeline
    try:
        with _lib.LifuSim(k["N"], (d,) * 3, k["dt"], k["Nt"], pml=(0, 0, 20)) as sim:
            sim.set_medium(k["c0"], k["rho0"], None)
            sim.set_source_geometry(k["idx"], np.arange(n_src + 1), np.zeros(n_src), np.ones(n_src), 1)
            sim.set_drive(k["sig"], [0], [1.0])
            p_max, p_min, stats = sim.run()
    finally:
        os.environ.pop("LIFU_PIPELINE", None)
    assert (stats["fft_launches"] == 0) == (pipeline == "v2") and stats["homogeneous"] == 0
    pm = p_max.reshape(nz, nxy, nxy)                       # x fastest
    assert np.allclose(pm, pm[:, :1, :1], rtol=1e-4, atol=1e-6 * pm.max())        # a plane wave stays a plane wave
    line = pm[:, 3, 5].astype(np.float64)
    half_pulse = int(0.5 * 4 / 500e3 * 1500.0 / d)         # cells covered by half the incident burst
    m1 = line[z_src + half_pulse + 4:z_int - half_pulse - 8]                       # no overlap with the echo here
    m2 = line[z_int + 12:z_int + 12 + int(0.7 * (nz - z_int))]
    assert m1.std() / m1.mean() < 0.01 and m2[: m2.size // 2].std() / m2.mean() < 0.01
    Z1, Z2 = k["Z"]
    T = 2 * Z2 / (Z1 + Z2)
    assert abs(m2[: m2.size // 2].mean() / m1.mean() - T) < 0.015 * T
    # on the far side of the interface the pressure is the transmitted burst from the first cell on: (1 + R) A_i = T A_i
    assert abs(line[z_int:z_int + 4].max() / m1.mean() - T) < 0.03 * T


def _lossy_wavenumber(f, c, alpha_db, y, dispersion=True):
    from tests.test_oracle_physics import lossy_wavenumber
    return lossy_wavenumber(f, c, alpha_db, y, dispersion)


def _plane_wave_run(pipeline, nxy, nz, d, z_src, t_end, medium, cycles=4, fields=(), f0=500e3, alpha_mode="binary"):
    """Plane source in a laterally periodic grid (no x / y PML) through the C ABI; returns the p_max line along z,
    optional final fields, dt and Nt."""
    import os
    from openlifu_b200 import _lib
    from tests.test_oracle_physics import planar_interface_inputs
    k = planar_interface_inputs(nxy=nxy, nz=nz, d=d, z_src=z_src, z_int=nz, t_end=t_end, cycles=cycles, f0=f0,
                                c=(1500.0, 1500.0), rho=(1000.0, 1000.0))
    n_src = k["idx"].size
    os.environ["LIFU_PIPELINE"] = pipeline
    try:
        with _lib.LifuSim(k["N"], (d,) * 3, k["dt"], k["Nt"], pml=(0, 0, 20)) as sim:
            sim.set_medium(*medium, alpha_power=0.9, alpha_mode=alpha_mode)
            sim.set_source_geometry(k["idx"], np.arange(n_src + 1), np.zeros(n_src), np.ones(n_src), 1)
            sim.set_drive(k["sig"], [0], [1.0])
            p_max, p_min, stats = sim.run()
            extra = [sim.get_field(w) for w in fields]
    finally:
        os.environ.pop("LIFU_PIPELINE", None)
    assert (stats["fft_launches"] == 0) == (pipeline == "v2")
    return p_max.reshape(nz, nxy, nxy)[:, 3, 5].astype(np.float64), extra, k["dt"], k["Nt"], stats


@pytest.mark.parametrize("pipeline,nxy,nz", [("v2", 64, 216), ("v1", 8, 364)])
@pytest.mark.parametrize("as_maps,alpha_mode", [(False, "binary"), (True, "binary"), (False, "no_dispersion"), (True, "no_dispersion")])
def test_plane_wave_power_law_absorption_known_answer(lifu_lib, pipeline, nxy, nz, as_maps, alpha_mode):
    """Oracle-independent: a narrow-band plane wave in an absorbing medium (alpha = 3 dB / (MHz^0.9 cm)) decays with the
    imaginary part of the wavenumber that solves the lossy dispersion relation: the nominal alpha f^y = 18.5 Np/m
    without the dispersion term, 20.6 Np/m with it (tan(pi y / 2) = 6.3 at y = 0.9 makes that term an 11 % change of
    c^2).  Scalar medium and per-voxel maps (both absorption code paths), both pipelines."""
    d, z_src, f0, y, alpha_db = 0.5e-3, 20, 500e3, 0.9, 3.0
    if as_maps:
        shape = (nxy, nxy, nz)
        medium = (np.full(shape, 1500.0), np.full(shape, 1000.0), np.full(shape, alpha_db))
        medium[0][0, 0, 0] = 1500.0000001                       # keep the maps from being recognised as uniform
    else:
        medium = (1500.0, 1000.0, alpha_db)
    t_end = (nz - z_src) * d / 1500.0 + 8 / f0
    line, _, dt, nt, stats = _plane_wave_run(pipeline, nxy, nz, d, z_src, t_end, medium, cycles=8, alpha_mode=alpha_mode)
    assert stats["absorbing"] == 1 and stats["homogeneous"] == (0 if as_maps else 1)
    half_pulse = int(0.5 * 8 / f0 * 1500.0 / d)
    z = np.arange(z_src + half_pulse + 4, nz - 30)
    slope = np.polyfit(z * d, np.log(line[z]), 1)[0]
    nominal = alpha_db * (f0 / 1e6) ** y * 100.0 / 8.685889638
    expected = _lossy_wavenumber(f0, 1500.0, alpha_db, y, dispersion=alpha_mode == "binary").imag
    assert abs(expected - (nominal if alpha_mode == "no_dispersion" else 1.1113 * nominal)) < 2e-3 * nominal
    assert abs(-slope - expected) < 0.025 * expected, (slope, expected)     # peak-amplitude fit of an 8-cycle burst


@pytest.mark.parametrize("pipeline,nxy,nz", [("v2", 64, 216), ("v1", 8, 364)])
def test_plane_wave_phase_speed_known_answer(lifu_lib, pipeline, nxy, nz):
    """Oracle-independent: the k-space scheme is exact in a homogeneous medium -- the pulse moves c0 * dt per step
    (two snapshots of the final pressure field, sub-cell shift from the cross-correlation)."""
    d, z_src = 0.5e-3, 20
    snaps = []
    for t_end in (30e-6, 42e-6):
        _, (p,), dt, nt, _ = _plane_wave_run(pipeline, nxy, nz, d, z_src, t_end, (1500.0, 1000.0, 0.0), fields=(0,))
        snaps.append((p[3, 5, :].astype(np.float64), nt))
    (a, n1), (b, n2) = snaps
    xc = np.correlate(b, a, mode="full")
    k = int(np.argmax(xc))
    y0, y1, y2 = xc[k - 1], xc[k], xc[k + 1]
    shift = (k - (a.size - 1)) + 0.5 * (y0 - y2) / (y0 - 2 * y1 + y2)
    c_meas = shift * d / ((n2 - n1) * dt)
    assert abs(c_meas - 1500.0) / 1500.0 < 2e-3, c_meas


@pytest.mark.parametrize("pipeline", ["v1", "v2"])
def test_anisotropic_spacing(lifu_lib, pipeline):
    """dx != dy != dz (the C ABI and get_kgrid carry one spacing per axis, kwave_if.py:20): k vectors, PML profiles,
    staggered shifts and the BLI supports all scale per axis; the source scale keeps k-Wave's dx."""
    case = cases.v2_small_case(steps=70) if pipeline == "v2" else cases.small_water_case()
    n = case["N"]
    sp = (1.0, 0.8, 1.25)
    lo = (-0.5 * (n[0] - 1) * sp[0], -0.5 * (n[1] - 1) * sp[1], -3.0)
    case["coords"] = [lo[a] + sp[a] * np.arange(n[a]) for a in range(3)]
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case, pipeline=pipeline)
    assert _is_v2(got) == (pipeline == "v2")
    assert np.array_equal(got["src_idx"], want["src_idx"])
    _check_fields(got, want)


def test_edge_cases_clipped_elements_long_drive_silent_array(lifu_lib):
    """(a) elements hanging over the edge of the grid: the source mask is the union of the BLI supports clipped to the
    inner grid, bit-exact; (b) a drive signal longer than the run (L > Nt): only the first Nt samples are injected;
    (c) a single time step; (d) zero apodization: the fields stay exactly zero."""
    # (a) + (b): 3 x 3 array of 6 mm pitch on a 25 x 23 grid (outer elements overhang), 40 cycles >> 30 steps
    case = cases.make_case([(-12, 12), (-11, 11), (-3, 27)], 1.0, 3, 3, 11.0, 0.5, (0, 0, 15), 400e3, 40,
                           dt=3e-7, t_end=30 * 3e-7, name="overhang")
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert got["stats"]["steps"] == 30 and got["stats"]["source_steps"] == 30
    assert np.array_equal(got["src_idx"], want["src_idx"])
    n_full = cases.run_oracle_case(cases.make_case([(-30, 30), (-30, 30), (-3, 27)], 1.0, 3, 3, 11.0, 0.5, (0, 0, 15), 400e3, 2,
                                                    dt=3e-7, t_end=3e-7))["src_idx"].size
    assert want["src_idx"].size < n_full                                   # really clipped
    _check_fields(got, want)
    # (c) one time step
    one = dict(case, t_end=3e-7)
    w1, g1 = cases.run_oracle_case(one), cases.run_cuda_case(one)
    assert g1["stats"]["steps"] == 1
    _check_fields(g1, w1)
    # (d) silent array
    silent = dict(cases.small_water_case(), apod=np.zeros(4))
    g0 = cases.run_cuda_case(silent)
    assert not g0["p_max"].any() and not g0["p_min"].any()


# ---------------------------------------------------------------------------------------------
# steady-state source of the fused pipeline: while every element is inside its burst the delayed drive signals have
# rank <= 2 in time, and the k-space source filter is applied once to two spatial basis fields (lifusim.cu: v2_build_steady)
def _steady_case(steps=140):
    case = cases.make_case([(-20, 19), (-22, 21), (-3, 32)], 1.0, 3, 3, 4.0, 0.5, (3, -2, 18), 400e3, 10,
                           dt=3e-7, t_end=steps * 3e-7, name="v2_steady")
    case["apod"] = np.linspace(0.4, 1.0, 9)
    return case


def test_v2_steady_source_matches_generic_path_and_oracle(lifu_lib, monkeypatch):
    case = _steady_case()
    want = cases.run_oracle_case(case)
    assert tuple(want["N_exp"]) == (64, 64, 64) and np.ptp(want["n_delay"]) >= 2       # several distinct delays
    monkeypatch.setenv("LIFU_SOURCE_STEADY", "1")
    on = cases.run_cuda_case(case)
    monkeypatch.setenv("LIFU_SOURCE_STEADY", "0")
    off = cases.run_cuda_case(case)
    monkeypatch.delenv("LIFU_SOURCE_STEADY")
    assert _is_v2(on) and _is_v2(off)
    n_base = int(np.ceil(10 / 400e3 / 3e-7))
    expect = int(want["n_delay"].min()) + n_base - int(want["n_delay"].max())
    assert on["stats"]["steady_source_steps"] == expect > 24 and off["stats"]["steady_source_steps"] == 0
    assert on["stats"]["kernel_launches"] < off["stats"]["kernel_launches"]
    _check_fields(on, want)
    _check_fields(off, want)
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(on[k], off[k]) < 3e-6, k


@pytest.mark.parametrize("medium", ["heterogeneous", "absorbing"])
def test_v2_steady_source_other_media(lifu_lib, medium):
    case = _steady_case(steps=110)
    c0, rho0, al = _phantom(tuple(case["N"]))
    case["c0"], case["rho0"], case["alpha"] = c0, rho0, (al if medium == "absorbing" else 0.0)
    case["dt"], case["t_end"] = 1.5e-7, 150 * 1.5e-7
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case)
    assert _is_v2(got) and got["stats"]["steady_source_steps"] > 24 and got["stats"]["homogeneous"] == 0
    _check_fields(got, want)


def test_v2_steady_source_needs_a_rank_two_drive(lifu_lib):
    """A drive that is not a sinusoidal burst (random samples) on three distinct delays: the rank test rejects it and every
    source step takes the generic path; a chirp likewise.  A constant-frequency burst of any phase is accepted, and so is
    ANY waveform when the array has only two distinct delays (two shifted copies always span a rank-2 space)."""
    from openlifu_b200 import _lib
    rng = np.random.default_rng(5)
    n_base = 90
    t = np.arange(n_base)
    drives = {"random": rng.standard_normal(n_base), "chirp": np.sin(0.02 * t * t),
              "tone": np.cos(0.31 * t + 0.4)}
    got = {}
    for name, base in drives.items():
        with _lib.LifuSim([44, 44, 44], [1e-3] * 3, 3e-7, 130, pml=(10, 10, 10)) as sim:
            sim.set_medium(1500.0, 1000.0, 0.0)
            idx = np.array([22 + 44 * (20 + 44 * 5), 23 + 44 * (24 + 44 * 6), 20 + 44 * (21 + 44 * 7)], dtype=np.int64)
            sim.set_source_geometry(idx, [0, 1, 2, 3], [0, 1, 2], [1.0, 0.7, 0.9], 3)
            sim.set_drive(base, [0, 3, 5], [1.0, 1.0, 0.5])
            _, _, st = sim.run()
            assert st["fft_launches"] == 0
            got[name] = st["steady_source_steps"]
            if name == "random":
                sim.set_drive(base, [0, 5, 5], [1.0, 1.0, 0.5])
                got["random, two delays"] = sim.run()[2]["steady_source_steps"]
    assert got["random"] == 0 and got["chirp"] == 0 and got["tone"] == n_base - 5 and got["random, two delays"] == n_base - 5
