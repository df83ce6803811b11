"""BASELINE.json configs C2 / C3 as plain-data cases (shared by the GPU tests and tools/parity_floor.py)."""
from __future__ import annotations

import numpy as np

from tests import cases


def c2_case(steps=None, c0=1500.0, rho0=1000.0, alpha=0.0, amplitude=1.0):
    """C2: the 2x64-element array of examples/legacy/OpenLIFU_2x_1.json, water, 0.5 mm, 216^3 inner -> 256^3;
    steps=None -> the reference's automatic time axis (Nt = 749)."""
    from openlifu_b200 import configs
    arr = configs.openlifu_2x_array()
    half = 53.75
    pos = np.array([el.position for el in arr.elements])
    size = np.array([el.size for el in arr.elements])
    ang = np.array([el.get_angle(units="deg") for el in arr.elements])
    kw = {}
    if steps is not None:
        dt = 0.5 * 0.5e-3 / 1500
        kw = dict(dt=dt, t_end=steps * dt)
    return cases.make_case([(-half, half), (-half, half), (-4, 103.5)], 0.5, 0, 0, 0, 0, (0, 0, 50), 400e3, 20,
                           elem_pos_mm=pos, elem_size_mm=size, angles_deg=ang, sensitivity=None, c0=c0, rho0=rho0,
                           alpha=alpha, amplitude=amplitude, **kw)


def c3_maps():
    """c, rho, alpha maps of the C3 skull / brain phantom on the 216^3 inner grid."""
    from openlifu_b200 import configs
    cfg = configs.c3(216)
    lab = np.asarray(cfg["volume"].data)
    mats = list(configs.PHANTOM_MATERIALS.values())
    c0 = np.array([m.sound_speed for m in mats])[lab]
    rho0 = np.array([m.density for m in mats])[lab]
    al = np.array([m.attenuation for m in mats])[lab]
    return c0, rho0, al


def c3_case(steps=240, amplitude=1.0):
    c0, rho0, al = c3_maps()
    return c2_case(steps=steps, c0=c0, rho0=rho0, alpha=al, amplitude=amplitude)
