"""GPU parity at CONFIG scale: BASELINE.json configs C2 and C3 on the headline grid (216^3 inner -> 256^3), the CUDA
path through the C ABI against the CPU oracle over the WHOLE run -- not against the other CUDA pipeline.

Tolerances (BASELINE.json north_star): source mask / delay samples / grid geometry bit-exact; p_max, p_min and the
intensity within 1e-4 relative L2 (float32 CUDA vs float32 oracle).  The float32-vs-float64 distance of the oracle
itself (the rounding floor) is recorded by tools/parity_floor.py in the committed sub-lattice goldens
tests/golden/c{2,3}_oracle_f64_sub6.npz; the CUDA result is compared with those too.

The oracle's time loop runs on the torch backend (same loop, all host threads): about 0.3 s per time step at 256^3 on
16 cores, i.e. ~4 min for the 749 steps of C2 and ~2 min for the 240 absorbing steps of C3.
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np
import pytest

from tests import cases
from tests.config_cases import c2_case, c3_case

pytestmark = pytest.mark.gpu

TOL = 1e-4
GOLDEN = Path(__file__).resolve().parent / "golden"


def _log(rec):
    """Append the measured distances to $LIFU_PARITY_LOG (a JSON-lines file) when set."""
    path = os.environ.get("LIFU_PARITY_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")
    print(json.dumps(rec))


def _intensity(p_min_flat, case):
    """kwave_if.py:140-141 on flat x-fastest vectors (float32 square and scale, float64 divide)."""
    n = tuple(case["N"])
    z = np.broadcast_to(np.asarray(case["rho0"], dtype=np.float64) * np.asarray(case["c0"], dtype=np.float64), n)
    two_z = (2 * z).flatten("F")
    return (np.float32(1e-4) * p_min_flat ** 2) / two_z


def _compare(name, case, got, want, steps):
    assert got["Nt"] == want["Nt"] == steps
    assert tuple(got["stats"]["n_exp"]) == tuple(want["N_exp"]) == (256, 256, 256)
    assert tuple(got["stats"]["pml"]) == tuple(want["pml"]) == (20, 20, 20)
    assert got["stats"]["fft_launches"] == 0, "the headline grid must run on the fused hand-written passes"
    assert np.array_equal(got["n_delay"], want["n_delay"])
    assert np.array_equal(got["src_idx"], want["src_idx"])                      # bit-exact source mask
    rec = {"workload": name, "steps": steps, "n_src": int(got["src_idx"].size)}
    for k in ("p_max", "p_min"):
        rec[f"{k}_rel_l2_vs_oracle_f32"] = cases.rel_l2(got[k], want[k])
    rec["intensity_rel_l2_vs_oracle_f32"] = cases.rel_l2(_intensity(got["p_min"], case), _intensity(want["p_min"], case))
    gold = GOLDEN / f"{name.lower()}_oracle_f64_sub6.npz"
    if gold.exists():
        g = np.load(gold)
        meta = json.loads(str(g["meta"]))
        if meta["steps"] == steps:
            s = meta["stride"]
            assert meta["n_src"] == got["src_idx"].size and meta["src_idx_sum"] == int(got["src_idx"].sum())
            rec["oracle_f32_vs_f64_floor"] = meta["floor_f32_vs_f64_rel_l2"]
            for k in ("p_max", "p_min"):
                sub = got[k].reshape(tuple(case["N"]), order="F")[::s, ::s, ::s]
                rec[f"{k}_rel_l2_vs_oracle_f64_sublattice"] = cases.rel_l2(sub, g[k])
    _log(rec)
    for k, v in rec.items():
        if k.endswith(("_vs_oracle_f32", "_vs_oracle_f64_sublattice")):
            assert v < TOL, f"{name} {k}: {v:.3e} >= {TOL}"


def test_c2_whole_run_against_oracle(lifu_lib):
    """C2 (2x64-element array, water): all 749 time steps, every inner voxel."""
    case = c2_case()
    got = cases.run_cuda_case(case)
    want = cases.run_oracle_case(case, backend="torch")
    _compare("C2", case, got, want, 749)


def test_c3_phantom_against_oracle(lifu_lib):
    """C3 (skull / brain phantom: c, rho, alpha maps; absorption and dispersion terms): 240 time steps -- the burst has
    crossed the skull shell (8..14 mm in front of the array) and 45 mm of tissue behind it."""
    case = c3_case(240)
    got = cases.run_cuda_case(case)
    assert got["stats"]["absorbing"] == 1 and got["stats"]["homogeneous"] == 0
    want = cases.run_oracle_case(case, backend="torch")
    _compare("C3", case, got, want, 240)
