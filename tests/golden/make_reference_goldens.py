#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference code (/root/reference) in the build
container.  The outputs (tests/golden/*.npz, *.json) are committed; this script only runs where
/root/reference exists and never on the GPU box.

The reference package cannot be imported whole here (xarray, vtk, h5py, kwave, ... are not
installed), so:
  * ``openlifu`` and ``openlifu.plan`` / ``openlifu.db`` are registered as bare namespace packages
    pointing at the reference sources, which skips their import-everything ``__init__``;
  * ``vtk`` is an empty stub (only used inside drawing functions that are never called);
  * ``xarray`` is this repo's labelled-array shim (openlifu_b200.xa) -- so the reference's own
    numpy logic (seg_method._map_params, sim_setup, solution_analysis) runs unmodified on top of it,
    which doubles as an API-coverage test of the shim.
Nothing under k-wave-python can be pinned this way: the solver oracle stays
"parity unpinned".
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
REF_SRC = Path("/root/reference/src")
sys.path.insert(0, str(ROOT / "openlifu-python_b200"))


def install_reference():
    if not REF_SRC.exists():
        raise SystemExit("/root/reference is not available: goldens can only be regenerated in the build container")
    from openlifu_b200 import xa as shim
    xr = types.ModuleType("xarray")
    for n in ("DataArray", "Dataset", "Coordinates", "concat", "open_dataset"):
        setattr(xr, n, getattr(shim, n))
    sys.modules["xarray"] = xr
    sys.modules["vtk"] = types.ModuleType("vtk")
    for pkg in ("openlifu", "openlifu.plan", "openlifu.db"):
        m = types.ModuleType(pkg)
        m.__path__ = [str(REF_SRC / pkg.replace(".", "/"))]
        sys.modules[pkg] = m
    # openlifu.util.json drags in the database layer; the pieces we call only need the encoder name
    js = types.ModuleType("openlifu.util.json")
    js.PYFUSEncoder = json.JSONEncoder
    sys.modules["openlifu.util.json"] = js


def main():
    install_reference()
    from openlifu import bf, geo, seg, xdc
    from openlifu.bf import apod_methods, delay_methods, focal_patterns
    from openlifu.sim.sim_setup import SimSetup
    from openlifu.util import units as ru
    from openlifu.plan import solution_analysis as rsa
    from openlifu_b200 import configs, xa

    out = {}

    # ---- units
    pairs = [("mm", "m"), ("m", "mm"), ("cm", "m"), ("um", "mm"), ("deg", "rad"), ("Pa", "MPa"), ("kPa", "Pa"),
             ("W/cm^2", "W/m^2"), ("mW/cm^2", "W/cm^2"), ("s", "ms"), ("us", "s"), ("MHz", "Hz"), ("kHz", "MHz"),
             ("mm^2", "cm^2"), ("mm^3", "m^3"), ("hour", "s"), ("min", "s")]
    out["unit_pairs"] = np.array([f"{a}>{b}" for a, b in pairs])
    out["unit_scales"] = np.array([ru.getunitconversion(a, b) for a, b in pairs])

    # ---- transducers
    arr1 = xdc.Transducer.gen_matrix_array(nx=8, ny=8, pitch=4, kerf=0.5, units="mm", sensitivity=1e5)
    arr2 = xdc.Transducer.from_file(str(REF_SRC.parent / "examples/legacy/OpenLIFU_2x_1.json"))
    out["c1_positions_mm"] = arr1.get_positions(units="mm")
    out["c2_positions_m"] = arr2.get_positions(units="m")
    out["c2_angles_deg"] = np.array([el.get_angle(units="deg") for el in arr2.elements])
    out["c2_sizes_m"] = np.array([el.get_size(units="m") for el in arr2.elements])
    out["c2_matrix_el5"] = arr2.elements[5].get_matrix(units="mm")
    out["c2_corners_el70_mm"] = arr2.elements[70].get_corners(units="mm")

    # ---- grids / params
    ss1 = SimSetup(spacing=1, x_extent=(-30, 30), y_extent=(-30, 30), z_extent=(-4, 70))
    ss2 = SimSetup(spacing=0.5, x_extent=(-53.75, 53.75), y_extent=(-53.75, 53.75), z_extent=(-4, 103.5))
    ss3 = SimSetup(spacing=0.3, x_extent=(-10, 10.1), y_extent=(-7, 8), z_extent=(0, 20.2))   # needs snapping
    for tag, ss in (("ss1", ss1), ("ss2", ss2), ("ss3", ss3)):
        out[f"{tag}_size"] = np.array(ss.get_size())
        out[f"{tag}_extent"] = np.array(ss.get_extent())
        c = ss.get_coords()
        for d in ("x", "y", "z"):
            out[f"{tag}_coord_{d}"] = np.asarray(c[d].data)
    params1 = seg.seg_methods.UniformWater().ref_params(ss1.get_coords())
    out["params1_c_ref"] = np.array(params1["sound_speed"].attrs["ref_value"])
    out["params1_c_unique"] = np.unique(params1["sound_speed"].data)

    # ---- label volume -> maps through the reference's _map_params (materials of the fixture protocol)
    mats = {k: seg.Material(**v.to_dict()) for k, v in configs.PHANTOM_MATERIALS.items()}
    coords_s = SimSetup(spacing=2, x_extent=(-10, 10), y_extent=(-8, 8), z_extent=(40, 70)).get_coords()
    labels = configs.skull_phantom_labels(coords_s, centre_mm=(0, 0, 70), r_in=18, r_out=24)
    sm = seg.seg_methods.UniformWater(materials=mats)
    pm = sm._map_params(labels, materials=mats)
    out["phantom_labels"] = labels.data
    for k in ("sound_speed", "density", "attenuation", "specific_heat", "thermal_conductivity"):
        out[f"phantom_{k}"] = pm[k].data
        out[f"phantom_{k}_ref"] = np.array(pm[k].attrs["ref_value"])

    # ---- beamforming
    targets = [(0, 0, 50), (5, -3, 42), (-12, 8, 60), (0, 0, 25)]
    d1, d2, a_max, a_pw, ang2 = [], [], [], [], []
    for tpos in targets:
        pt = geo.Point(position=np.array(tpos, dtype=float), units="mm")
        d1.append(delay_methods.Direct().calc_delays(arr1, pt, params1))
        d2.append(delay_methods.Direct(c0=1540).calc_delays(arr2, pt, None))
        a_max.append(apod_methods.MaxAngle(30).calc_apodization(arr2, pt, params1))
        a_pw.append(apod_methods.PiecewiseLinear(zero_angle=40, rolloff_angle=15).calc_apodization(arr2, pt, params1))
        ang2.append([el.angle_to_point(pt.get_position(units="m"), units="m", return_as="deg") for el in arr2.elements])
    out["targets_mm"] = np.array(targets, dtype=float)
    out["direct_delays_c1"] = np.array(d1)
    out["direct_delays_c2_c1540"] = np.array(d2)
    out["maxangle30_c2"] = np.array(a_max)
    out["piecewise_40_15_c2"] = np.array(a_pw)
    out["angles_deg_c2"] = np.array(ang2)
    out["effective_origin_c2"] = arr2.get_effective_origin(np.array(a_pw[1]), units="mm")

    # ---- focal patterns / focus frames
    tgt = geo.Point(position=np.array([3.0, -4.0, 50.0]), units="mm", id="tgt", name="T", radius=2)
    wh = focal_patterns.Wheel(center=True, num_spokes=31, spoke_radius=5.0, distance_units="mm")
    foci = wh.get_targets(tgt)
    out["wheel32_positions"] = np.array([f.position for f in foci])
    out["wheel32_ids"] = np.array([f.id for f in foci])
    out["focus_matrix"] = tgt.get_matrix(center_on_point=True)
    out["focus_matrix_origin0"] = geo.Point(position=np.zeros(3)).get_matrix()
    out["single_positions"] = np.array([f.position for f in focal_patterns.SinglePoint().get_targets(tgt)])

    # ---- drive signals (kwave_if.py:101-103 + Transducer.calc_output)
    for tag, arr, dly, cyc, dt in (("c1", arr1, d1[0], 10, 0.5 * 1e-3 / 1500), ("c2", arr2, d2[1], 20, 0.5 * 0.5e-3 / 1500)):
        t = np.arange(0, cyc / 400e3, dt)
        sig = 1.0 * np.sin(2 * np.pi * 400e3 * t)
        apod = np.linspace(0.5, 1.0, arr.numelements())
        mat = arr.calc_output(sig, dt, dly, apod)
        out[f"drive_{tag}_shape"] = np.array(mat.shape)
        out[f"drive_{tag}_first_nonzero"] = np.array([int(np.flatnonzero(r)[0]) if np.any(r) else -1 for r in mat])
        out[f"drive_{tag}_rows"] = mat[[0, 7, arr.numelements() - 1]]
        out[f"drive_{tag}_delays"] = dly
        out[f"drive_{tag}_dt"] = np.array(dt)

    # ---- solution analysis pieces on a synthetic focal field (next-row component)
    cs = SimSetup(spacing=1, x_extent=(-15, 15), y_extent=(-12, 12), z_extent=(20, 70)).get_coords()
    X, Y, Z = np.meshgrid(cs["x"].data, cs["y"].data, cs["z"].data, indexing="ij")
    field = np.exp(-(((X - 2) / 3.0) ** 2 + ((Y + 1) / 2.5) ** 2 + ((Z - 45) / 9.0) ** 2)) \
        + 0.3 * np.exp(-(((X + 8) / 2.0) ** 2 + (Y / 2.0) ** 2 + ((Z - 50) / 4.0) ** 2))
    da = xa.DataArray(field, coords=cs, dims=("x", "y", "z"), attrs={"units": "MPa"})
    focus = np.array([2.0, -1.0, 45.0])
    out["sa_field"] = field
    out["sa_focus"] = focus
    out["sa_offset_grid"] = rsa.get_offset_grid(da, focus, as_dataset=False)
    out["sa_dist"] = rsa.calc_dist_from_focus(da, focus, aspect_ratio=[1, 1, 5], as_dataarray=False)
    for op in ("<", "<=", ">", ">="):
        out[f"sa_mask_{op}"] = np.asarray(rsa.get_mask(da, focus, distance=6.0, aspect_ratio=[1, 1, 5], operator=op).data)
    out["sa_centroid_half"] = rsa.find_centroid(da, 0.5 * field.max(), units="mm")
    out["sa_focus_matrix"] = rsa.get_focus_matrix(focus, origin=[1.0, 0.5, 0.0])
    bw, bounds = [], []
    for dim in ("x", "y", "z"):
        for frac in (10 ** (-3 / 20), 10 ** (-6 / 20)):
            cutoff = float(field.max()) * frac
            bw.append(rsa.get_beamwidth(da, focus, dim=dim, cutoff=cutoff))
            bounds.append(rsa.get_beam_bounds(da, focus, dim=dim, cutoff=cutoff))
    out["sa_beamwidths"] = np.array(bw)
    out["sa_beam_bounds"] = np.array(bounds)
    line = rsa.interp_transformed_axis(da, focus, "z", min_offset=-10.0, max_offset=12.0)
    out["sa_line_z"] = np.asarray(line.data)
    out["sa_line_z_offsets"] = np.asarray(line.coords["offset_dz"].data)

    # ---- candidate poses: the reference's TransformedTransducer.bake (xdc/transducer.py:412-417)
    from dataclasses import fields as dc_fields
    from openlifu.xdc.transducer import TransformedTransducer
    rng = np.random.default_rng(147)
    poses = []
    for _ in range(3):
        ax, ay, az = rng.uniform(-0.25, 0.25, 3)
        rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        m = np.eye(4)
        m[:3, :3] = rz @ ry @ rx
        m[:3, 3] = rng.uniform(-4, 4, 3)
        poses.append(m)
    out["bake_transforms"] = np.array(poses)
    for tag, arr in (("c1", arr1), ("c2", arr2)):
        pos, ang = [], []
        for m in poses:
            kw = {f.name: getattr(arr.copy(), f.name) for f in dc_fields(xdc.Transducer)}
            baked = TransformedTransducer(transform=m, **kw).bake()
            pos.append(baked.get_positions(units="mm"))
            ang.append([el.get_angle(units="deg") for el in baked.elements])
        out[f"bake_{tag}_positions_mm"] = np.array(pos)
        out[f"bake_{tag}_angles_deg"] = np.array(ang)

    np.savez_compressed(HERE / "ref_beamform.npz", **out)
    print("wrote", HERE / "ref_beamform.npz", "with", len(out), "arrays")


if __name__ == "__main__":
    main()
