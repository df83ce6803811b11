"""GPU parity of the z-slab decomposed solve through the C ABI.

One rank (G = 1) exercises the whole decomposed pipeline -- 2-D transforms, exchange kernels, 1-D z
transforms, transposed-layout operators -- on a single GPU; G = 2 runs under torchrun when the box has two."""
from __future__ import annotations

import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
TOL = 1e-4          # north-star tolerance: relative L2 of p_max / p_min


def _one_rank(case, exchange, **kw):
    from openlifu_b200 import _lib
    return cases.run_cuda_case_slab(case, 0, 1, _lib.slab_unique_id(), exchange=exchange, **kw)


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_single_rank_slab_matches_single_gpu_and_oracle(lifu_lib, exchange):
    case = cases.v2_small_case(steps=60)
    got = _one_rank(case, exchange)
    one = cases.run_cuda_case(case, pipeline="v1")
    want = cases.run_oracle_case(case)
    assert got["layout"]["exchange"] == {"nccl": 1, "peer": 2}[exchange]
    assert got["layout"]["sensor_nz"] == case["N"][2]
    # peer stores on a 64^3 grid: the fused passes with the exchange in their store phase (no library transform);
    # NCCL transport: cuFFT 2-D / 1-D transforms around pack / send / unpack
    assert (got["stats"]["fft_launches"] == 0) == (exchange == "peer")
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], one[k]) < 1e-5, k
        assert cases.rel_l2(got[k], want[k]) < TOL, k


def test_single_rank_slab_peer_library_transforms(lifu_lib, monkeypatch):
    """LIFU_SLAB_WIDE=0 keeps the cuFFT-based slab path under the peer-store exchange."""
    monkeypatch.setenv("LIFU_SLAB_WIDE", "0")
    case = cases.v2_small_case(steps=40)
    got = _one_rank(case, "peer")
    one = cases.run_cuda_case(case, pipeline="v1")
    assert got["stats"]["fft_launches"] > 0
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], one[k]) < 1e-5, k


@pytest.mark.parametrize("medium", ["water", "phantom"])
def test_single_rank_slab_fused_passes_on_a_wide_grid(lifu_lib, medium):
    """128 x 64 x 128 expanded grid (8 x 16 factorisations on x and z) through the slab code path of the fused passes:
    routed stores into the (own) transposed buffer, source planes through T4[3], sensor crop of the local planes."""
    case = cases.wide_case((128, 64, 128), steps=60)
    if medium == "phantom":
        case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
        case["dt"], case["t_end"] = 1.5e-7, 60 * 1.5e-7
    got = _one_rank(case, "peer")
    assert got["stats"]["fft_launches"] == 0
    one = cases.run_cuda_case(case, pipeline="v1")
    want = cases.run_oracle_case(case)
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], one[k]) < 2e-5, k
        assert cases.rel_l2(got[k], want[k]) < TOL, k


def test_single_rank_slab_odd_grid_heterogeneous_absorbing(lifu_lib):
    """Odd-sized expanded grid, c/rho/alpha maps given as a plane range, both absorption terms."""
    case = cases.small_water_case()
    case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
    case["dt"], case["t_end"] = 1.5e-7, 80 * 1.5e-7
    got = _one_rank(case, "auto")
    want = cases.run_oracle_case(case)
    assert got["stats"]["homogeneous"] == 0 and got["stats"]["absorbing"] == 1
    for k in ("p_max", "p_min"):
        assert cases.rel_l2(got[k], want[k]) < TOL, k


def test_slab_needs_divisible_grid(lifu_lib):
    from openlifu_b200 import _lib
    case = cases.small_water_case()                      # expanded 45 x 45 x 54-class grid: Ny odd
    with pytest.raises(ValueError):
        cases.run_cuda_case_slab(case, 0, 2, _lib.slab_unique_id())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ranks(tmp_path, world, case, exchange):
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "slab_rank.py"), "--case", case,
           "--exchange", exchange, "--out", str(out)]
    r = subprocess.run(cmd, cwd=str(ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["finite"] and res["world"] == world
    tol_single = 2e-5 if case.startswith("wide") else 1e-5        # fused passes vs the library-FFT pipeline on one GPU
    for k in ("p_max", "p_min"):
        assert res["vs_single"][k] < tol_single, res
        assert res["vs_oracle"][k] < TOL, res
    return res


@pytest.mark.parametrize("case,exchange", [("water", "peer"), ("water", "nccl"), ("phantom", "auto"), ("wide", "peer"),
                                           ("wide_phantom", "peer")])
def test_two_rank_slab(lifu_lib, tmp_path, case, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    _run_ranks(tmp_path, 2, case, exchange)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("case,exchange", [("water", "nccl"), ("phantom", "peer")])
def test_many_rank_slab(lifu_lib, tmp_path, world, case, exchange):
    """64^3 expanded grid over 4 / 8 ranks (8 planes per rank at 8: the first and last ranks hold only PML planes
    and write no sensor data)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    res = _run_ranks(tmp_path, world, case, exchange)
    assert sum(nz for _, nz in res["planes"]) == 36


@pytest.mark.parametrize("case", ["water", "phantom"])
def test_two_rank_run_simulation_slab_mode(lifu_lib, tmp_path, case):
    """openlifu_b200.sim.run_simulation with LIFU_MULTI_GPU=slab: same call on every rank, whole Dataset back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = tmp_path / "res.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "slab_rank.py"), "--case", case,
           "--api", "--out", str(out)]
    r = subprocess.run(cmd, cwd=str(ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert res["finite"] and res["repeat_equal"] and res["shape"] == [40, 44, 36]
    assert max(res["all_ranks_vs_single"]) < 1e-5, res
