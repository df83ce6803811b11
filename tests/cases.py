"""Shared test scenes: the same plain-data description is fed to the oracle and to the C ABI."""
from __future__ import annotations

import numpy as np

from oracle import beamform as obf
from oracle import scene as osc


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make_case(extents_mm, spacing_mm, nx, ny, pitch_mm, kerf_mm, focus_mm, freq, cycles, sensitivity=1e5,
              c0=1500.0, rho0=1000.0, alpha=0.0, dt=0.0, t_end=0.0, angles_deg=None, elem_pos_mm=None,
              elem_size_mm=None, amplitude=1.0, apod=None, name="case"):
    coords = obf.grid_coords(extents_mm, spacing_mm)
    if elem_pos_mm is None:
        pos, size, _ = obf.matrix_array(nx, ny, pitch_mm, kerf_mm)
    else:
        pos, size = np.asarray(elem_pos_mm, dtype=np.float64), np.asarray(elem_size_mm, dtype=np.float64)
    n_el = len(pos)
    ang = np.zeros((n_el, 3)) if angles_deg is None else np.asarray(angles_deg, dtype=np.float64)
    c_ref = float(np.max(c0)) if np.ndim(c0) else float(c0)
    delays = obf.direct_delays(pos * 1e-3, np.asarray(focus_mm, dtype=np.float64) * 1e-3, 1500.0 if np.ndim(c0) else c_ref)
    return {
        "name": name, "coords": coords, "N": [len(c) for c in coords], "pos_m": pos * 1e-3, "size_m": size * 1e-3,
        "angles_deg": ang, "c0": c0, "rho0": rho0, "alpha": alpha, "sensitivity": sensitivity,
        "delays": delays, "apod": np.ones(n_el) if apod is None else np.asarray(apod, dtype=np.float64),
        "freq": freq, "cycles": cycles, "amplitude": amplitude, "dt": dt, "t_end": t_end,
    }


def small_water_case():
    """2x2 array, 1 mm grid 25x23x31 -> PML-expanded 45x45x54-class grid, a few dozen steps."""
    return make_case([(-12, 12), (-11, 11), (-3, 27)], 1.0, 2, 2, 3.0, 0.5, (0, 0, 15), 400e3, 2,
                     dt=3e-7, t_end=60 * 3e-7, name="small_water")


def v2_small_case(steps=80):
    """Inner 40x44x36 at 1 mm -> PML (12,10,14) -> 64^3: the smallest grid the fused-FFT pipeline (v2) takes."""
    return make_case([(-20, 19), (-22, 21), (-3, 32)], 1.0, 2, 2, 3.0, 0.5, (0, 0, 18), 400e3, 2,
                     dt=3e-7, t_end=steps * 3e-7, name="v2_small")


WIDE_PML = (10, 10, 10)


def wide_case(n_exp, steps=60, name="wide"):
    """Water, 1 mm grid whose PML-expanded size (PML 10 per side) is n_exp; 2 x 2 array at z = 0 focused at 18 mm.
    The grids of pipeline "wide" (csrc/fft_wide.cuh): 128 / 512 / 768 / 1024-point axes next to 64 and 256."""
    n = [e - 2 * p for e, p in zip(n_exp, WIDE_PML)]
    ext = [(-(n[0] // 2), n[0] - 1 - n[0] // 2), (-(n[1] // 2), n[1] - 1 - n[1] // 2), (-3, n[2] - 4)]
    case = make_case(ext, 1.0, 2, 2, 3.0, 0.5, (0, 0, 18), 400e3, 2, dt=3e-7, t_end=steps * 3e-7, name=name)
    assert case["N"] == n, (case["N"], n)
    return case


def c1_case():
    """8x8 array, 4 mm pitch, focus 50 mm, water, 1 mm grid."""
    return make_case([(-30, 30), (-30, 30), (-4, 70)], 1.0, 8, 8, 4.0, 0.5, (0, 0, 50), 400e3, 10, name="C1")


def scene_of(case):
    return osc.Scene(coords=case["coords"], coord_scale=1e-3, elem_pos_m=case["pos_m"], elem_size_m=case["size_m"],
                     elem_angles_deg=case["angles_deg"], sound_speed=case["c0"], density=case["rho0"],
                     attenuation=case["alpha"], sensitivity=case["sensitivity"])


def run_oracle_case(case, dtype=np.float32, asm=None, max_steps=None, backend="numpy"):
    """backend "torch": the same time loop on torch CPU tensors (all host threads; equal to the numpy loop to 6e-7,
    tests/test_oracle_physics.py) -- used by the config-scale comparisons, where the numpy loop would take twice as long."""
    out = osc.run_simulation(scene_of(case), delays=case["delays"], apod=case["apod"], freq=case["freq"],
                             cycles=case["cycles"], amplitude=case["amplitude"], dt=case["dt"], t_end=case["t_end"],
                             dtype=dtype, asm=asm, max_steps=max_steps, backend=backend)
    return {"p_max": out["raw"]["p_max"], "p_min": out["raw"]["p_min"], "src_idx": out["src_idx"], "W": out["W"],
            "n_delay": out["n_delay"], "Nt": out["Nt"], "dt": out["dt"], "pml": out["raw"]["pml"],
            "N_exp": out["raw"]["N_exp"]}


def run_cuda_case(case, alpha_mode="binary", source_mode="additive", geometry=None, max_steps=None, device=0,
                  pipeline=None, fields=(), pml=(-1, -1, -1)):
    """Drive the C ABI exactly like openlifu_b200.sim.run_simulation does, from plain data.
    pipeline: None (auto) | "v1" (cuFFT) | "v2" (fused FFT passes); read by lifu_create from LIFU_PIPELINE."""
    import os
    from openlifu_b200 import _lib

    if pipeline is None:
        os.environ.pop("LIFU_PIPELINE", None)
    else:
        os.environ["LIFU_PIPELINE"] = pipeline

    sc = scene_of(case)
    N, d, Nt, dt = osc.time_axis(sc, case["dt"], case["t_end"], 0.5)
    if max_steps is not None:
        Nt = min(Nt, max_steps)
    offset = np.array([-float(np.mean(c)) * 1e-3 for c in case["coords"]])
    t = np.arange(0, case["cycles"] / case["freq"], dt)
    base = case["amplitude"] * np.sin(2 * np.pi * case["freq"] * t)
    if case["sensitivity"] is not None:
        base = base * case["sensitivity"]
    n_delay = np.array([int(dl / dt) for dl in case["delays"]], dtype=np.int32)
    with _lib.LifuSim(N, d, dt, Nt, device=device, pml=pml) as sim:
        sim.set_medium(case["c0"], case["rho0"], case["alpha"], alpha_power=0.9, alpha_mode=alpha_mode)
        if geometry is None:
            sim.set_elements(case["pos_m"] + offset, case["size_m"], case["angles_deg"], 0.05, 5)
        else:
            sim.set_source_geometry(*geometry)
        sim.set_drive(base, n_delay, case["apod"], source_mode=source_mode)
        p_max, p_min, stats = sim.run()
        idx, row_ptr, col, w, n_el = sim.get_source_geometry()
        extra = {f"field{k}": sim.get_field(k) for k in fields}
    os.environ.pop("LIFU_PIPELINE", None)
    return {**extra, "p_max": p_max, "p_min": p_min, "stats": stats, "src_idx": idx, "row_ptr": row_ptr, "col": col, "w": w,
            "n_delay": n_delay, "Nt": Nt, "dt": dt}


def csr_to_dense(n_src, n_el, row_ptr, col, w):
    W = np.zeros((n_src, n_el), dtype=np.float32)
    for i in range(n_src):
        W[i, col[row_ptr[i]:row_ptr[i + 1]]] = w[row_ptr[i]:row_ptr[i + 1]]
    return W


def layered_phantom(N):
    """Layered water / skull-like slab / tissue phantom with per-material c, rho, alpha."""
    c0 = np.full(N, 1500.0)
    rho0 = np.full(N, 1000.0)
    al = np.full(N, 0.0022)
    z = np.arange(N[2])[None, None, :] + 0 * np.arange(N[0])[:, None, None]
    x = np.arange(N[0])[:, None, None]
    slab = np.broadcast_to((z + (x // 6) >= 12) & (z + (x // 6) < 16), N)
    tissue = np.broadcast_to((z + (x // 6)) >= 16, N)
    c0[slab], rho0[slab], al[slab] = 2800.0, 1900.0, 6.0
    c0[tissue], rho0[tissue], al[tissue] = 1540.0, 1050.0, 0.3
    return c0, rho0, al


def run_cuda_case_slab(case, rank, world, nccl_id, exchange="auto", alpha_mode="binary", source_mode="additive",
                       max_steps=None, device=0, fields=()):
    """One rank's share of the same scene on a z-slab decomposed grid (C ABI: lifu_create_slab ...).
    Returns this rank's inner planes of p_max / p_min (x fastest) and the layout."""
    from openlifu_b200 import _lib

    sc = scene_of(case)
    N, d, Nt, dt = osc.time_axis(sc, case["dt"], case["t_end"], 0.5)
    if max_steps is not None:
        Nt = min(Nt, max_steps)
    offset = np.array([-float(np.mean(c)) * 1e-3 for c in case["coords"]])
    t = np.arange(0, case["cycles"] / case["freq"], dt)
    base = case["amplitude"] * np.sin(2 * np.pi * case["freq"] * t)
    if case["sensitivity"] is not None:
        base = base * case["sensitivity"]
    n_delay = np.array([int(dl / dt) for dl in case["delays"]], dtype=np.int32)
    with _lib.LifuSim(N, d, dt, Nt, device=device, slab=(rank, world, nccl_id, exchange)) as sim:
        lay = dict(sim.layout)
        if np.ndim(case["c0"]) == 0 and np.ndim(case["rho0"]) == 0 and np.ndim(case["alpha"]) == 0:
            sim.set_medium(case["c0"], case["rho0"], case["alpha"], alpha_power=0.9, alpha_mode=alpha_mode)
        else:
            lo, n = lay["medium_z0"], lay["medium_nz"]
            maps = [np.broadcast_to(np.asarray(m, dtype=np.float64), tuple(N))[:, :, lo:lo + n]
                    for m in (case["c0"], case["rho0"], case["alpha"])]
            sim.set_medium(*maps, alpha_power=0.9, alpha_mode=alpha_mode, plane0=lo)   # only the planes this rank reads
        sim.set_elements(case["pos_m"] + offset, case["size_m"], case["angles_deg"], 0.05, 5)
        sim.set_drive(base, n_delay, case["apod"], source_mode=source_mode)
        p_max, p_min, stats = sim.run()
        extra = {f"field{k}": sim.get_field(k) for k in fields}
    return {**extra, "p_max": p_max, "p_min": p_min, "stats": stats, "layout": lay, "N": N, "Nt": Nt}
