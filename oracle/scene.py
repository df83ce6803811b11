"""Oracle: ``run_simulation`` on plain numpy data (TEST INFRASTRUCTURE).

Follows /root/reference/src/openlifu/sim/kwave_if.py:80-146 line by line, with the
k-wave-python objects replaced by the restatements in ``oracle.kgrid``, ``oracle.bli``
and ``oracle.solver``.  Inputs are plain arrays so that the oracle does not depend on
the product package.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import beamform as bfm
from . import bli
from . import kgrid as kg
from .solver import Assumptions, SolverInputs, simulate


@dataclass
class Scene:
    coords: list                      # three 1-D coordinate vectors (kwave_if.py:19-20)
    coord_scale: float                # unit -> metres (getunitconversion(units[0], 'm'), :18)
    elem_pos_m: np.ndarray            # (n_el,3) el.get_position(units='m')        (:39)
    elem_size_m: np.ndarray           # (n_el,2) el.get_size(units='m')            (:40)
    elem_angles_deg: np.ndarray       # (n_el,3) el.get_angle(units='deg') = (el, az, roll) (:41)
    sound_speed: object               # scalar or (Nx,Ny,Nz) map                   (:57-59)
    density: object
    attenuation: object
    sensitivity: float | None = None  # Transducer.sensitivity (transducer.py:105-106)
    elem_gain: np.ndarray | None = None
    extras: dict = field(default_factory=dict)


def time_axis(scene: Scene, dt=0.0, t_end=0.0, cfl=0.5):
    """get_kgrid (kwave_if.py:13-27)."""
    N = [len(c) for c in scene.coords]
    d = [float(np.diff(c)[0] * scene.coord_scale) for c in scene.coords]
    if dt == 0 or t_end == 0:
        Nt, dt_ = kg.make_time(N, d, 1500.0, cfl)
    else:
        Nt, dt_ = kg.set_time(t_end, dt)
    return N, d, Nt, dt_


def source_geometry(scene: Scene, bli_tolerance=0.05, upsampling_rate=5):
    """get_karray + get_array_binary_mask (kwave_if.py:108-112, 75)."""
    N = [len(c) for c in scene.coords]
    d = [float(np.diff(c)[0] * scene.coord_scale) for c in scene.coords]
    offset = [-float(np.mean(c)) * scene.coord_scale for c in scene.coords]
    return bli.array_source_geometry(N, d, scene.elem_pos_m, scene.elem_size_m, scene.elem_angles_deg,
                                     offset, bli_tolerance, upsampling_rate, single_precision=True)


def run_simulation(scene: Scene, delays=None, apod=None, freq=1e6, cycles=20, amplitude=1.0, dt=0.0,
                   t_end=0.0, cfl=0.5, bli_tolerance=0.05, upsampling_rate=5, ref_values_only=False,
                   dtype=np.float32, asm: Assumptions | None = None, geometry=None, max_steps=None,
                   workers=-1, backend="numpy"):
    """Returns dict with p_max (PPP), p_min (PNP, sign flipped), intensity on the inner grid in
    (Nx,Ny,Nz) layout plus the raw Fortran-flat solver output and integer geometry."""
    n_el = len(scene.elem_pos_m)
    delays = np.zeros(n_el) if delays is None else np.asarray(delays, dtype=np.float64)
    apod = np.ones(n_el) if apod is None else np.asarray(apod, dtype=np.float64)
    N, d, Nt, dt_ = time_axis(scene, dt, t_end, cfl)
    source_mat, n_delay, _ = bfm.drive_signals(freq, cycles, amplitude, dt_, delays, apod,
                                               scene.sensitivity, scene.elem_gain)
    idx, W = geometry if geometry is not None else source_geometry(scene, bli_tolerance, upsampling_rate)
    src_p = bli.distributed_source_signal(W, source_mat, single_precision=True)
    if ref_values_only:
        c0, rho0, al = (float(np.ravel(scene.extras["ref_values"][k])[0]) for k in ("sound_speed", "density", "attenuation"))
    else:
        c0, rho0, al = scene.sound_speed, scene.density, scene.attenuation
    inp = SolverInputs(N=tuple(N), d=tuple(d), dt=dt_, Nt=Nt, c0=c0, rho0=rho0, alpha_db=al,
                       src_idx=idx, src_p=src_p)
    raw = simulate(inp, dtype=dtype, asm=asm, max_steps=max_steps, workers=workers, backend=backend)
    sz = tuple(N)
    p_max = raw["p_max"].reshape(sz, order="F")
    p_min_raw = raw["p_min"].reshape(sz, order="F")
    Z = np.asarray(scene.density, dtype=np.float64) * np.asarray(scene.sound_speed, dtype=np.float64)
    intensity = 1e-4 * p_min_raw.astype(np.float64) ** 2 / (2 * Z)     # kwave_if.py:140-141 (float64)
    return {"p_max": p_max, "p_min": -1 * p_min_raw, "intensity": intensity,
            "raw": raw, "src_idx": idx, "W": W, "n_delay": n_delay, "Nt": Nt, "dt": dt_,
            "N": tuple(N), "d": tuple(d), "source_mat": source_mat}
