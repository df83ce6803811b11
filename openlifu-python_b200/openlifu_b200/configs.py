"""Benchmark / parity workloads C1..C4 of SURVEY.md section 8d, built through the public API.

C1  8x8 matrix array, focus 50 mm, water, 1 mm grid        -> 81x81x125,  Nt 229
C2  2x64-element OpenLIFU array, water, 0.5 mm grid         -> 256^3,      Nt 749
C3  C2 grid + synthetic skull/brain phantom (c, rho, alpha maps)
C4  C2 + Wheel focal pattern (32 foci), one simulation per GPU
"""
from __future__ import annotations

import numpy as np

from . import xa
from .bf import Pulse, Sequence, apod_methods, delay_methods, focal_patterns
from .geo import Point
from .seg import Material, seg_methods
from .sim import SimSetup
from .xdc import Element, Transducer

# Two 8x8 modules (5 mm pitch, 4.7 mm elements) tilted by -/+0.24735 rad about y, centred at
# (+/-24.4839, 0, 3.0436) mm: the geometry of the reference's examples/legacy/OpenLIFU_2x_1.json.
_MODULE_TILT = 0.2473539555718235
_MODULE_CX = 24.4839310826181
_MODULE_CZ = 3.04363291282334


def openlifu_2x_array() -> Transducer:
    local = (np.arange(8) - 3.5) * 5.0
    elements = []
    idx = 1
    for sign in (+1.0, -1.0):
        az = -sign * _MODULE_TILT
        order = local[::-1] if sign > 0 else -local
        for xl in order:           # columns from the outer edge inwards (module 1), mirrored for module 2
            for y in local:
                x = sign * _MODULE_CX + (xl if sign > 0 else xl) * np.cos(az)
                z = _MODULE_CZ - xl * np.sin(az)
                elements.append(Element(index=idx, pin=idx, position=[x, y, z], orientation=[az, 0.0, 0.0],
                                        size=[4.7, 4.7], impulse_response=[1.0], impulse_dt=1, units="mm"))
                idx += 1
    return Transducer(id="openlifu_2x", name="OpenLIFU 2x", elements=elements, frequency=400.6e3, units="mm")


def c1():
    arr = Transducer.gen_matrix_array(nx=8, ny=8, pitch=4, kerf=0.5, units="mm", sensitivity=1e5)
    setup = SimSetup(spacing=1, x_extent=(-30, 30), y_extent=(-30, 30), z_extent=(-4, 70))
    pulse = Pulse(frequency=400e3, duration=10 / 400e3)
    return {"name": "C1", "arr": arr, "setup": setup, "pulse": pulse, "target": Point(position=(0, 0, 50), units="mm"),
            "seg": seg_methods.UniformWater()}


def c2(n_inner: int = 216):
    """n_inner = 216 is the headline grid (PML 20 -> 256^3); other sizes keep the 0.5 mm spacing."""
    arr = openlifu_2x_array()
    half = (n_inner - 1) * 0.5 / 2.0
    setup = SimSetup(spacing=0.5, x_extent=(-half, half), y_extent=(-half, half), z_extent=(-4, -4 + 2 * half))
    pulse = Pulse(frequency=400e3, duration=20 / 400e3)
    return {"name": "C2", "arr": arr, "setup": setup, "pulse": pulse, "target": Point(position=(0, 0, 50), units="mm"),
            "seg": seg_methods.UniformWater()}


PHANTOM_MATERIALS = {
    "water": Material("water", 1500.0, 1000.0, 0.0022, 4182.0, 0.598),
    "tissue": Material("tissue", 1540.0, 1050.0, 0.3, 3600.0, 0.528),
    "skull": Material("skull", 2800.0, 1900.0, 6.0, 1300.0, 0.4),
}


def skull_phantom_labels(coords, centre_mm=(0.0, 0.0, 70.0), r_in=56.0, r_out=62.0):
    """Spherical skull shell (label 2) around brain tissue (label 1) in water (label 0)."""
    x, y, z = (np.asarray(coords[d].data) for d in ("x", "y", "z"))
    r = np.sqrt((x[:, None, None] - centre_mm[0]) ** 2 + (y[None, :, None] - centre_mm[1]) ** 2
                + (z[None, None, :] - centre_mm[2]) ** 2)
    labels = np.zeros(r.shape, dtype=int)
    labels[r < r_in] = 1
    labels[(r >= r_in) & (r <= r_out)] = 2
    return xa.DataArray(labels, coords=coords, dims=("x", "y", "z"))


def c3(n_inner: int = 216):
    cfg = c2(n_inner)
    cfg["name"] = "C3"
    cfg["seg"] = seg_methods.LabelVolume(materials=dict(PHANTOM_MATERIALS), ref_material="water")
    cfg["volume"] = skull_phantom_labels(cfg["setup"].get_coords())
    return cfg


def c4(n_inner: int = 216, num_spokes: int = 31):
    cfg = c2(n_inner)
    cfg["name"] = "C4"
    cfg["focal_pattern"] = focal_patterns.Wheel(center=True, num_spokes=num_spokes, spoke_radius=5.0, distance_units="mm")
    cfg["sequence"] = Sequence(pulse_interval=0.1, pulse_count=num_spokes + 1, pulse_train_interval=0.1 * (num_spokes + 1))
    return cfg


def prepare(cfg):
    """params Dataset, foci, and per-focus (delays, apod) exactly as Protocol.calc_solution would
    (plan/protocol.py:300-321 of the reference)."""
    params = cfg["setup"].setup_sim_scene(cfg["seg"], volume=cfg.get("volume"))
    pattern = cfg.get("focal_pattern", focal_patterns.SinglePoint())
    foci = pattern.get_targets(cfg["target"])
    dm = cfg.get("delay_method", delay_methods.Direct())
    am = cfg.get("apod_method", apod_methods.Uniform())
    beams = [(dm.calc_delays(cfg["arr"], f, params), am.calc_apodization(cfg["arr"], f, params)) for f in foci]
    cycles = float(np.min([np.round(cfg["pulse"].duration * cfg["pulse"].frequency), 20]))
    return params, foci, beams, cycles
