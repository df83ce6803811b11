"""Pin the oracle's beamforming restatement AND the product's host-side mirror against vectors
produced by the real reference code (tests/golden/make_reference_goldens.py -> ref_beamform.npz).
Integer results must be bit-exact; float64 results agree to rounding."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

import openlifu_b200 as ol
from openlifu_b200 import configs, xa
from openlifu_b200.util.units import getunitconversion
from oracle import beamform as obf

G = np.load(Path(__file__).parent / "golden" / "ref_beamform.npz", allow_pickle=False)
TOL = dict(rtol=1e-12, atol=1e-15)


def _arrays():
    a1 = ol.Transducer.gen_matrix_array(nx=8, ny=8, pitch=4, kerf=0.5, units="mm", sensitivity=1e5)
    return a1, configs.openlifu_2x_array()


def test_units_table():
    for pair, want in zip(G["unit_pairs"], G["unit_scales"]):
        a, b = str(pair).split(">")
        assert getunitconversion(a, b) == want


def test_transducer_geometry():
    a1, a2 = _arrays()
    assert np.allclose(a1.get_positions(units="mm"), G["c1_positions_mm"], **TOL)
    assert np.allclose(obf.matrix_array(8, 8, 4, 0.5)[0], G["c1_positions_mm"], **TOL)
    assert np.allclose(a2.get_positions(units="m"), G["c2_positions_m"], rtol=0, atol=1e-13)   # generator vs JSON
    assert np.allclose(np.array([el.get_angle(units="deg") for el in a2.elements]), G["c2_angles_deg"], **TOL)
    assert np.allclose(np.array([el.get_size(units="m") for el in a2.elements]), G["c2_sizes_m"], **TOL)
    assert np.allclose(a2.elements[5].get_matrix(units="mm"), G["c2_matrix_el5"], rtol=1e-12, atol=1e-11)
    assert np.allclose(a2.elements[70].get_corners(units="mm"), G["c2_corners_el70_mm"], rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("tag,kw", [("ss1", dict(spacing=1, x_extent=(-30, 30), y_extent=(-30, 30), z_extent=(-4, 70))),
                                    ("ss2", dict(spacing=0.5, x_extent=(-53.75, 53.75), y_extent=(-53.75, 53.75), z_extent=(-4, 103.5))),
                                    ("ss3", dict(spacing=0.3, x_extent=(-10, 10.1), y_extent=(-7, 8), z_extent=(0, 20.2)))])
def test_sim_grid(tag, kw):
    ss = ol.SimSetup(**kw)
    assert np.array_equal(np.array(ss.get_size()), G[f"{tag}_size"])               # integer N bit-exact
    assert np.array_equal(np.array(ss.get_extent()), G[f"{tag}_extent"])
    c = ss.get_coords()
    oc = obf.grid_coords([kw["x_extent"], kw["y_extent"], kw["z_extent"]], kw["spacing"])
    for i, d in enumerate("xyz"):
        assert np.array_equal(c[d].data, G[f"{tag}_coord_{d}"])
        assert np.array_equal(oc[i], G[f"{tag}_coord_{d}"])
        assert c[d].attrs["units"] == "mm"
    assert tuple(c.dims) == ("x", "y", "z")


def test_param_maps_from_labels():
    coords = ol.SimSetup(spacing=2, x_extent=(-10, 10), y_extent=(-8, 8), z_extent=(40, 70)).get_coords()
    labels = configs.skull_phantom_labels(coords, centre_mm=(0, 0, 70), r_in=18, r_out=24)
    assert np.array_equal(labels.data, G["phantom_labels"])
    sm = ol.seg_methods.LabelVolume(materials=dict(configs.PHANTOM_MATERIALS), ref_material="water")
    pm = sm.seg_params(labels)
    for k in ("sound_speed", "density", "attenuation", "specific_heat", "thermal_conductivity"):
        assert np.array_equal(pm[k].data, G[f"phantom_{k}"])
        assert pm[k].attrs["ref_value"] == G[f"phantom_{k}_ref"]
        assert pm[k].data.dtype == np.float64
    water = ol.seg_methods.UniformWater().ref_params(ol.SimSetup(x_extent=(-30, 30), y_extent=(-30, 30), z_extent=(-4, 70)).get_coords())
    assert water["sound_speed"].attrs["ref_value"] == G["params1_c_ref"]
    assert np.array_equal(np.unique(water["sound_speed"].data), G["params1_c_unique"])


def test_delays_and_apodizations():
    a1, a2 = _arrays()
    params = ol.seg_methods.UniformWater().ref_params(ol.SimSetup(x_extent=(-30, 30), y_extent=(-30, 30), z_extent=(-4, 70)).get_coords())
    for i, tpos in enumerate(G["targets_mm"]):
        pt = ol.Point(position=tpos, units="mm")
        assert np.allclose(ol.delay_methods.Direct().calc_delays(a1, pt, params), G["direct_delays_c1"][i], rtol=1e-11, atol=1e-18)
        assert np.allclose(ol.delay_methods.Direct(c0=1540).calc_delays(a2, pt, None), G["direct_delays_c2_c1540"][i], rtol=1e-9, atol=1e-16)
        assert np.allclose(obf.direct_delays(a1.get_positions(units="m"), tpos * 1e-3, 1500.0), G["direct_delays_c1"][i], rtol=1e-11, atol=1e-18)
        assert np.array_equal(ol.apod_methods.MaxAngle(30).calc_apodization(a2, pt, params), G["maxangle30_c2"][i])
        pw = ol.apod_methods.PiecewiseLinear(zero_angle=40, rolloff_angle=15).calc_apodization(a2, pt, params)
        assert np.allclose(pw, G["piecewise_40_15_c2"][i], rtol=1e-9, atol=1e-12)
        ang = [np.degrees(obf.angle_to_point(el.position * 1e-3, el.orientation, tpos * 1e-3)) for el in a2.elements]
        assert np.allclose(ang, G["angles_deg_c2"][i], rtol=1e-9, atol=1e-9)
        assert np.array_equal(obf.apod_maxangle(ang, 30), G["maxangle30_c2"][i])
        assert np.allclose(obf.apod_piecewise_linear(ang, 40, 15), G["piecewise_40_15_c2"][i], rtol=1e-9, atol=1e-12)
    assert np.allclose(a2.get_effective_origin(G["piecewise_40_15_c2"][1], units="mm"), G["effective_origin_c2"], rtol=1e-10, atol=1e-10)


def test_focal_patterns():
    tgt = ol.Point(position=np.array([3.0, -4.0, 50.0]), units="mm", id="tgt", name="T", radius=2)
    foci = ol.focal_patterns.Wheel(center=True, num_spokes=31, spoke_radius=5.0, distance_units="mm").get_targets(tgt)
    assert len(foci) == 32
    assert np.allclose(np.array([f.position for f in foci]), G["wheel32_positions"], rtol=1e-13, atol=1e-13)
    assert [f.id for f in foci] == [str(s) for s in G["wheel32_ids"]]
    assert np.allclose(obf.wheel_targets(tgt.position, True, 31, 5.0), G["wheel32_positions"], rtol=1e-13, atol=1e-13)
    assert np.allclose(tgt.get_matrix(center_on_point=True), G["focus_matrix"], rtol=1e-13, atol=1e-15)
    assert np.allclose(obf.focus_matrix(tgt.position), G["focus_matrix"], rtol=1e-13, atol=1e-15)
    assert np.allclose(ol.Point(position=np.zeros(3)).get_matrix(), G["focus_matrix_origin0"])
    assert np.allclose(ol.focal_patterns.SinglePoint().get_targets(tgt)[0].position, G["single_positions"][0])


@pytest.mark.parametrize("tag,cyc", [("c1", 10), ("c2", 20)])
def test_drive_signals(tag, cyc):
    a1, a2 = _arrays()
    arr = a1 if tag == "c1" else a2
    dt = float(G[f"drive_{tag}_dt"])
    dly = G[f"drive_{tag}_delays"]
    t = np.arange(0, cyc / 400e3, dt)
    sig = 1.0 * np.sin(2 * np.pi * 400e3 * t)
    apod = np.linspace(0.5, 1.0, arr.numelements())
    mat = arr.calc_output(sig, dt, dly, apod)
    assert np.array_equal(np.array(mat.shape), G[f"drive_{tag}_shape"])
    first = np.array([int(np.flatnonzero(r)[0]) if np.any(r) else -1 for r in mat])
    assert np.array_equal(first, G[f"drive_{tag}_first_nonzero"])                 # delay sample counts bit-exact
    assert np.allclose(mat[[0, 7, arr.numelements() - 1]], G[f"drive_{tag}_rows"], rtol=1e-13, atol=1e-9)
    n_delay, gains, base_gain = arr.drive_plan(dt, dly, apod)
    omat, on_delay, _ = obf.drive_signals(400e3, cyc, 1.0, dt, dly, apod, sensitivity=arr.sensitivity,
                                          elem_gain=[el.scalar_gain() for el in arr.elements])
    assert np.array_equal(on_delay, n_delay)
    assert np.array_equal(np.array(omat.shape), G[f"drive_{tag}_shape"])
    assert np.allclose(omat[[0, 7, arr.numelements() - 1]], G[f"drive_{tag}_rows"], rtol=1e-13, atol=1e-9)


def test_candidate_poses_equal_the_reference_bake():
    """candidate_transducer places the array like the reference's TransformedTransducer.bake
    (xdc/transducer.py:412-417, 297-301); golden positions / Euler angles come from the real reference classes."""
    from openlifu_b200 import configs
    from openlifu_b200.plan.protocol import candidate_transducer
    from openlifu_b200.xdc import Transducer
    g = G
    arrays = {"c1": Transducer.gen_matrix_array(nx=8, ny=8, pitch=4, kerf=0.5, units="mm", sensitivity=1e5),
              "c2": configs.openlifu_2x_array()}
    for tag, arr in arrays.items():
        before = arr.get_positions(units="mm").copy()
        for i, m in enumerate(g["bake_transforms"]):
            baked = candidate_transducer(arr, m)
            np.testing.assert_allclose(baked.get_positions(units="mm"), g[f"bake_{tag}_positions_mm"][i], rtol=0, atol=1e-10)
            ang = np.array([el.get_angle(units="deg") for el in baked.elements])
            np.testing.assert_allclose(ang, g[f"bake_{tag}_angles_deg"][i], rtol=0, atol=1e-9)
        assert np.array_equal(arr.get_positions(units="mm"), before)        # the input transducer is not touched
