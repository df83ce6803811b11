"""GPU parity of pipeline "wide" (csrc/fft_wide.cuh, wide.cu, wide_x.cu): the fused FFT passes of pipeline v2 for axis
lengths N = A x B with A <= B -- 128 (8 x 16), 512 (16 x 32), 768 (24 x 32), 1024 (32 x 32) next to 64 and 256 --
i.e. the grids pml_auto produces for BASELINE config C5 (768^3) and the slab leg of the bench (512^3).
Same oracle and tolerances as tests/test_gpu_parity.py; every run must report zero library FFTs."""
from __future__ import annotations

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu

TOL = 1e-4
PML = cases.WIDE_PML
wide_case = cases.wide_case


def _oracle(case, **kw):
    from oracle.solver import Assumptions
    return cases.run_oracle_case(case, asm=Assumptions(pml_size=PML, **kw))


def _check(got, want, n_exp, tol=TOL):
    assert tuple(got["stats"]["n_exp"]) == tuple(n_exp)
    assert got["stats"]["fft_launches"] == 0, "library FFTs were used"
    assert np.array_equal(got["src_idx"], want["src_idx"])
    for k in ("p_max", "p_min"):
        err = cases.rel_l2(got[k], want[k])
        assert err < tol, f"{k}: rel-L2 {err:.3e} >= {tol}"


@pytest.mark.parametrize("n_exp", [(128, 64, 64), (64, 128, 64), (64, 64, 128), (128, 128, 128)])
def test_wide_128_axes_match_oracle_and_library_pipeline(lifu_lib, n_exp):
    """8 x 16 factorisation on each axis in turn (half of the lanes idle in the second DFT), then on all three; state
    fields compared with the cuFFT pipeline."""
    case = wide_case(n_exp, steps=70)
    want = _oracle(case)
    got = cases.run_cuda_case(case, pipeline="v2", pml=PML, fields=(0, 1, 2, 3, 4, 5, 6))
    _check(got, want, n_exp)
    v1 = cases.run_cuda_case(case, pipeline="v1", pml=PML, fields=(0, 1, 2, 3, 4, 5, 6))
    assert v1["stats"]["fft_launches"] > 0
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], v1[k]) < 2e-5
    for f in range(7):
        assert cases.rel_l2(got[f"field{f}"], v1[f"field{f}"]) < 2e-4, f"state field {f}"


def test_wide_on_a_square_grid_equals_v2(lifu_lib, monkeypatch):
    """64 x 256 x 64 takes pipeline v2 by default; the wide kernels instantiated for 8 x 8 and 16 x 16 must agree with it
    (LIFU_WIDE_SQUARE=1 routes square grids through them)."""
    case = wide_case((64, 256, 64), steps=50)
    want = _oracle(case)
    v2 = cases.run_cuda_case(case, pipeline="v2", pml=PML)
    monkeypatch.setenv("LIFU_WIDE_SQUARE", "1")
    got = cases.run_cuda_case(case, pipeline="v2", pml=PML)
    _check(got, want, (64, 256, 64))
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], v2[k]) < 2e-5


@pytest.mark.parametrize("alpha_mode", ["binary", "no_dispersion", "no_absorption"])
def test_wide_heterogeneous_absorbing(lifu_lib, alpha_mode):
    n_exp = (64, 128, 128)
    case = wide_case(n_exp, steps=80)
    case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 80 * 1.5e-7
    want = _oracle(case, absorb_eta=alpha_mode != "no_dispersion", absorb_tau=alpha_mode != "no_absorption")
    got = cases.run_cuda_case(case, alpha_mode=alpha_mode, pipeline="v2", pml=PML)
    assert got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    _check(got, want, n_exp)


def test_wide_homogeneous_absorbing(lifu_lib):
    n_exp = (128, 64, 128)
    case = wide_case(n_exp, steps=60)
    case["alpha"], case["c0"], case["rho0"] = 0.75, 1540.0, 1050.0
    _check(cases.run_cuda_case(case, pipeline="v2", pml=PML), _oracle(case), n_exp)


@pytest.mark.parametrize("n_exp", [(512, 64, 64), (64, 768, 64), (64, 64, 768), (768, 64, 64), (64, 64, 512), (1024, 64, 64),
                                   (64, 1024, 64)])
def test_wide_long_axes(lifu_lib, n_exp):
    """512 = 16 x 32, 768 = 24 x 32 (DFT-24 = 3 x 8), 1024 = 32 x 32 on one axis of a thin grid: oracle and cuFFT pipeline."""
    case = wide_case(n_exp, steps=40)
    want = _oracle(case)
    got = cases.run_cuda_case(case, pipeline="v2", pml=PML, fields=(0, 4))
    _check(got, want, n_exp)
    v1 = cases.run_cuda_case(case, pipeline="v1", pml=PML, fields=(0, 4))
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], v1[k]) < 2e-5
    for f in (0, 4):
        assert cases.rel_l2(got[f"field{f}"], v1[f"field{f}"]) < 2e-4, f"state field {f}"


@pytest.mark.parametrize("n_exp", [(768, 64, 64), (512, 128, 64)])
def test_wide_long_x_axis_phantom(lifu_lib, n_exp):
    """Heterogeneous absorbing medium on a long x axis: the 64-thread x kernels with staged medium rows."""
    case = wide_case(n_exp, steps=60)
    case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 60 * 1.5e-7
    want = _oracle(case)
    got = cases.run_cuda_case(case, pipeline="v2", pml=PML)
    assert got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    _check(got, want, n_exp)


def test_wide_is_taken_automatically(lifu_lib):
    """Without LIFU_PIPELINE a 128-point axis goes through the fused passes (LIFU_WIDE_AUTO=0 would give the library FFTs)."""
    case = wide_case((128, 64, 64), steps=20)
    got = cases.run_cuda_case(case, pml=PML)
    assert got["stats"]["fft_launches"] == 0
