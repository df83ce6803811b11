"""The C-ABI library loads on a CPU-only box and exports every symbol include/lifusim.h declares;
its pure-host helpers agree bit-exactly with the oracle.  No compute call is made here."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import kgrid as okg

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "lifusim.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lifu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lifu_lib):
    lib = ctypes.CDLL(str(lifu_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in lifusim.h but not exported"
    assert set(names) == set(lifu_lib.EXPORTED)
    assert lifu_lib.load().lifu_abi_version() == 1


@pytest.mark.parametrize("n", [(61, 61, 75), (216, 216, 216), (728, 728, 728), (21, 21, 13), (25, 23, 31), (48, 47, 40), (100, 37, 64)])
def test_pml_auto_matches_oracle(lifu_lib, n):
    assert lifu_lib.pml_auto(n) == okg.optimal_pml_size(n)


def test_known_pml_sizes(lifu_lib):
    assert lifu_lib.pml_auto((61, 61, 75)) == (10, 10, 25)       # -> 81 x 81 x 125
    assert lifu_lib.pml_auto((216, 216, 216)) == (20, 20, 20)    # -> 256^3 (C2)
    assert lifu_lib.pml_auto((728, 728, 728)) == (20, 20, 20)    # -> 768^3 (C5)


@pytest.mark.parametrize("n,d,cfl", [((61, 61, 75), (1e-3,) * 3, 0.5), ((216,) * 3, (0.5e-3,) * 3, 0.5),
                                     ((728,) * 3, (0.25e-3,) * 3, 0.5), ((21, 21, 13), (1e-3,) * 3, 0.3),
                                     ((50, 60, 70), (0.7e-3, 0.7e-3, 0.7e-3), 0.25)])
def test_make_time_matches_oracle(lifu_lib, n, d, cfl):
    nt, dt = lifu_lib.make_time(n, d, 1500.0, cfl)
    ont, odt = okg.make_time(n, d, 1500.0, cfl)
    assert nt == ont and dt == odt


def test_known_time_axes(lifu_lib):
    assert lifu_lib.make_time((61, 61, 75), (1e-3,) * 3)[0] == 229
    assert lifu_lib.make_time((216,) * 3, (0.5e-3,) * 3)[0] == 749
    assert lifu_lib.make_time((728,) * 3, (0.25e-3,) * 3)[0] == 2522


def test_create_without_gpu_fails_loudly(lifu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lifu_lib.LifuError, match="no CPU fallback"):
        lifu_lib.LifuSim((16, 16, 16), (1e-3,) * 3, 1e-7, 4)


def test_bad_arguments_are_value_errors(lifu_lib):
    with pytest.raises(ValueError):
        lifu_lib.pml_auto((0, 4, 4))
    with pytest.raises(ValueError):
        lifu_lib.make_time((4, 4, 4), (1e-3,) * 3, c_ref=-1.0)


def test_device_analysis_without_gpu_fails_loudly(lifu_lib):
    """Solution.analyze(engine="cuda") has no CPU fallback: without a GPU the C ABI refuses."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ax = [np.linspace(0, 1, 4)] * 3
    with pytest.raises(lifu_lib.LifuError, match="no CPU fallback"):
        lifu_lib.BeamAnalysis(ax, 1)
