"""world_size-2 `gloo` tests of the N>1 host logic (no GPU): foci of a sweep sharded over ranks
(the per-focus loop of /root/reference/src/openlifu/plan/protocol.py:318-339)
and the slab-decomposed 3-D FFT of row 2 (local 2-D transforms + all-to-all + 1-D transform)."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, *args):
    import queue

    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q, args)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        # drain before joining: a child blocks in put() until its (large) result has been read
        for _ in range(world):
            r, v = q.get(timeout=180)
            out[r] = v
    except queue.Empty:
        pass
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    assert len(out) == world and all(p.exitcode == 0 for p in procs), ([p.exitcode for p in procs], sorted(out))
    for r, v in out.items():
        assert not isinstance(v, str) or not v.startswith("ERROR"), v
    return out


def _entry(fn, rank, world, port, q, args):
    for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.pop("LOCAL_RANK", None)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, "ERROR " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def _fake_run_simulation(**kw):
    """Deterministic stand-in for the solver: the field depends on the focus through its delays only."""
    from openlifu_b200 import xa
    c = kw["params"].coords
    X, Y, Z = np.meshgrid(*[c[d].data for d in ("x", "y", "z")], indexing="ij")
    k = 1.0 + 1e6 * float(np.sum(np.abs(kw["delays"])))
    f = (1e5 * k) * np.exp(-((X / 3) ** 2 + (Y / 3) ** 2 + ((Z - 30) / 8) ** 2))
    mk = lambda a, u: xa.DataArray(a, coords=c, dims=("x", "y", "z"), attrs={"units": u})  # noqa: E731
    return xa.Dataset({"p_max": mk(f.astype(np.float32), "Pa"), "p_min": mk(f.astype(np.float32), "Pa"),
                       "intensity": mk(1e-4 * f ** 2 / 3e6, "W/cm^2")}), None


def _sweep(rank, world, n_spokes):
    from openlifu_b200.bf import Pulse, Sequence, focal_patterns
    from openlifu_b200.geo import Point
    from openlifu_b200.plan import Protocol
    from openlifu_b200.plan import protocol as pmod
    from openlifu_b200.sim import SimSetup
    from openlifu_b200.xdc import Transducer
    pmod.run_simulation = _fake_run_simulation
    n_foci = n_spokes + 1
    pr = Protocol(pulse=Pulse(frequency=400e3, duration=25e-6), sequence=Sequence(pulse_interval=0.01, pulse_count=n_foci, pulse_train_interval=0),
                  focal_pattern=focal_patterns.Wheel(center=True, num_spokes=n_spokes, spoke_radius=3, distance_units="mm", target_pressure=0.5, units="MPa"),
                  sim_setup=SimSetup(spacing=2.0, x_extent=(-10, 10), y_extent=(-10, 10), z_extent=(0, 40)))
    arr = Transducer.gen_matrix_array(nx=4, ny=4, pitch=4, kerf=0.5, units="mm", sensitivity=1e4)
    target = Point(position=np.array([0, 0, 30.0]), units="mm", id="tgt")
    calls = []
    orig = pmod.run_simulation

    def counting(**kw):
        calls.append(1)
        return orig(**kw)

    pmod.run_simulation = counting
    sol, agg, analysis = pr.calc_solution(target, arr, simulate=True, scale=True, use_gpu=True)
    res = sol.simulation_result
    return {"calls": len(calls), "p_min": np.asarray(res["p_min"].data), "p_max": np.asarray(res["p_max"].data),
            "intensity": np.asarray(res["intensity"].data), "agg": np.asarray(agg["p_min"].data),
            "pnp": list(analysis.mainlobe_pnp_MPa), "dims": tuple(res["p_min"].dims)}


@pytest.mark.parametrize("n_spokes", [4, 5])
def test_foci_sharded_over_two_ranks_matches_serial(n_spokes):
    """Rank r simulates foci r, r+2, ...; after the gather both ranks hold the serial result."""
    two = _spawn(_sweep, 2, n_spokes)
    one = _spawn(_sweep, 1, n_spokes)[0]
    n_foci = n_spokes + 1
    assert one["calls"] == n_foci
    assert two[0]["calls"] == (n_foci + 1) // 2 and two[1]["calls"] == n_foci // 2      # ragged split when odd
    for r in (0, 1):
        assert two[r]["dims"] == one["dims"] and two[r]["p_min"].shape[0] == n_foci
        for k in ("p_min", "p_max", "intensity", "agg"):
            assert two[r][k].dtype == one[k].dtype
            np.testing.assert_array_equal(two[r][k], one[k])
        np.testing.assert_allclose(two[r]["pnp"], one["pnp"], rtol=0, atol=0)


# ------------------------------------------------------------------------------------------------
def _slab_fft(rank, world, shape, seed):
    """oracle.slab: the decomposition the CUDA slab path uses, with gloo all-to-all in place of NVLink."""
    from oracle import slab as oslab
    rng = np.random.default_rng(seed)
    Nx, Ny, Nz = shape
    full = rng.standard_normal((Nz, Ny, Nx)).astype(np.float64)          # x fastest
    nzl = Nz // world
    mine = full[rank * nzl:(rank + 1) * nzl]
    T = oslab.forward(mine, world)                                         # [Nz][Ny/world][Nxh], this rank's ky rows
    want = np.fft.fftn(full, axes=(0, 1, 2))[:, :, : Nx // 2 + 1]
    nyl = Ny // world
    err_f = float(np.abs(T - want[:, rank * nyl:(rank + 1) * nyl, :]).max())
    back = oslab.inverse(T, world, Nx)
    err_b = float(np.abs(back - mine).max())
    return err_f, err_b


@pytest.mark.parametrize("shape", [(16, 8, 12), (15, 6, 10), (8, 8, 8)])
def test_slab_fft_two_ranks_matches_fftn(shape):
    out = _spawn(_slab_fft, 2, shape, 147)
    for r in (0, 1):
        assert out[r][0] < 1e-10 and out[r][1] < 1e-12, out


def _slab_fft_routed(rank, world, shape, ab_y, ab_z, seed):
    """The exchange of the fused passes (csrc/fft_wide.cuh with Q.G > 0): every spectrum value goes from the thread /
    register that holds it after the line transform straight into the owner's buffer, by block arithmetic."""
    from oracle import slab as oslab
    rng = np.random.default_rng(seed)
    Nx, Ny, Nz = shape
    full = rng.standard_normal((Nz, Ny, Nx)).astype(np.float64)
    nzl, nyl = Nz // world, Ny // world
    mine = full[rank * nzl:(rank + 1) * nzl]
    T = np.fft.fft(oslab.routed_forward(np.fft.rfft2(mine, axes=(1, 2)), world, ab_y), axis=0)
    want = np.fft.fftn(full, axes=(0, 1, 2))[:, :, : Nx // 2 + 1]
    err_f = float(np.abs(T - want[:, rank * nyl:(rank + 1) * nyl, :]).max())
    H = oslab.routed_backward(np.fft.ifft(T, axis=0), world, ab_z, rank)
    back = np.fft.irfft2(H, s=(Ny, Nx), axes=(1, 2))
    return err_f, float(np.abs(back - mine).max())


@pytest.mark.parametrize("shape,ab_y,ab_z", [((8, 64, 128), (8, 8), (8, 16)), ((6, 128, 64), (8, 16), (8, 8))])
def test_slab_fft_routed_stores_two_ranks_matches_fftn(shape, ab_y, ab_z):
    out = _spawn(_slab_fft_routed, 2, shape, ab_y, ab_z, 147)
    for r in (0, 1):
        assert out[r][0] < 1e-9 and out[r][1] < 1e-12, out


# ------------------------------------------------------------------------------------------------
def _candidates(rank, world, n_poses):
    """Protocol.simulate_candidates under torch.distributed: pose i runs on rank i mod world, every rank ends up with
    every pose's fields."""
    from openlifu_b200.bf import Pulse, Sequence
    from openlifu_b200.geo import Point
    from openlifu_b200.plan import Protocol, SolutionAnalysisOptions
    from openlifu_b200.plan import protocol as pmod
    from openlifu_b200.sim import SimSetup
    from openlifu_b200.xdc import Transducer
    calls = []

    def counting(**kw):
        calls.append(float(np.sum(np.abs(kw["delays"]))))
        return _fake_run_simulation(**kw)

    pmod.run_simulation = counting
    pr = Protocol(pulse=Pulse(frequency=400e3, duration=25e-6), sequence=Sequence(pulse_interval=0.01, pulse_count=1, pulse_train_interval=0),
                  sim_setup=SimSetup(spacing=2.0, x_extent=(-10, 10), y_extent=(-10, 10), z_extent=(0, 40)))
    arr = Transducer.gen_matrix_array(nx=4, ny=4, pitch=4, kerf=0.5, units="mm", sensitivity=1e4)
    target = Point(position=np.array([0, 0, 30.0]), units="mm", id="tgt")
    transforms = []
    for i in range(n_poses):
        m = np.eye(4)
        m[0, 3], m[2, 3] = 1.5 * i, -0.5 * i
        transforms.append(m)
    opts = SolutionAnalysisOptions(mainlobe_radius=4.0, beamwidth_radius=8.0, sidelobe_radius=6.0, sidelobe_zmin=1.0, distance_units="mm")
    res = pr.simulate_candidates(target, arr, transforms, analysis_options=opts, use_gpu=True, analyze=True)
    return {"calls": len(calls), "p_min": [np.asarray(s.simulation_result["p_min"].data) for s, _ in res],
            "delays": [np.asarray(s.delays) for s, _ in res], "pnp": [a.mainlobe_pnp_MPa[0] for _, a in res],
            "pos0": [s.transducer.get_positions(units="mm")[0].tolist() for s, _ in res]}


def test_candidate_poses_sharded_over_two_ranks_matches_serial():
    two = _spawn(_candidates, 2, 5)
    one = _spawn(_candidates, 1, 5)[0]
    assert one["calls"] == 5 and two[0]["calls"] == 3 and two[1]["calls"] == 2
    assert len({tuple(p) for p in one["pos0"]}) == 5                       # five different poses
    for r in (0, 1):
        assert two[r]["pos0"] == one["pos0"]
        for i in range(5):
            assert two[r]["p_min"][i].shape[0] == 1
            np.testing.assert_array_equal(two[r]["p_min"][i], one["p_min"][i])
            np.testing.assert_array_equal(two[r]["delays"][i], one["delays"][i])
        np.testing.assert_allclose(two[r]["pnp"], one["pnp"], rtol=0, atol=0)
