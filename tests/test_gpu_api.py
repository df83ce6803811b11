"""The public API on one B200: openlifu_b200.sim.run_simulation and Protocol.calc_solution (the calls OpenLIFU
makes, /root/reference/src/openlifu/sim/kwave_if.py:80-146 and plan/protocol.py:242-398) against the CPU oracle
run on the same plain-data scene.  Tolerance: 1e-4 relative L2 on p_max / p_min / intensity (north star); the
Dataset layout (dims, dtypes, attrs, DataArray names) is compared exactly."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import scene as osc
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _scene(arr, params, sensitivity):
    coords = [np.asarray(params.coords[d].data, dtype=np.float64) for d in ("x", "y", "z")]
    pos = np.array([el.get_position(units="m") for el in arr.elements])
    size = np.array([el.get_size(units="m") for el in arr.elements])
    ang = np.array([el.get_angle(units="deg") for el in arr.elements])
    return osc.Scene(coords=coords, coord_scale=1e-3, elem_pos_m=pos, elem_size_m=size, elem_angles_deg=ang,
                     sound_speed=np.asarray(params["sound_speed"].data), density=np.asarray(params["density"].data),
                     attenuation=np.asarray(params["attenuation"].data), sensitivity=sensitivity)


def _setup(spacing=1.0, steps=70):
    from openlifu_b200.seg import seg_methods
    from openlifu_b200.sim import SimSetup
    from openlifu_b200.xdc import Transducer
    arr = Transducer.gen_matrix_array(nx=3, ny=2, pitch=3, kerf=0.5, units="mm", sensitivity=1e5)
    setup = SimSetup(spacing=spacing, x_extent=(-14, 14), y_extent=(-12, 12), z_extent=(-3, 30), dt=2.5e-7, t_end=steps * 2.5e-7)
    params = setup.setup_sim_scene(seg_methods.UniformWater())
    return arr, setup, params


def test_run_simulation_matches_oracle_and_reference_layout(lifu_lib, monkeypatch):
    monkeypatch.setenv("LIFU_PACKAGING", "device")     # exercise lifu_get_packaged; compared with the host packaging below
    from openlifu_b200.bf import delay_methods
    from openlifu_b200.geo import Point
    from openlifu_b200.sim import run_simulation
    arr, setup, params = _setup()
    delays = delay_methods.Direct().calc_delays(arr, Point(position=(1.0, -2.0, 20.0), units="mm"), params)
    apod = np.array([1.0, 0.0, 0.5, 1.0, 1.0, 0.25])            # one element switched off, ragged weights
    ds, out = run_simulation(arr=arr, params=params, delays=delays, apod=apod, freq=400e3, cycles=3, amplitude=2.0,
                             dt=setup.dt, t_end=setup.t_end, gpu=True)
    want = osc.run_simulation(_scene(arr, params, 1e5), delays=delays, apod=apod, freq=400e3, cycles=3, amplitude=2.0,
                              dt=setup.dt, t_end=setup.t_end)
    assert list(ds.data_vars) == ["p_max", "p_min", "intensity"]                 # kwave_if.py:142-145
    n = tuple(len(params.coords[d]) for d in ("x", "y", "z"))
    for k, dtype, units, long_name in (("p_max", np.float32, "Pa", "PPP"), ("p_min", np.float32, "Pa", "PNP"),
                                       ("intensity", np.float64, "W/cm^2", "Intensity")):
        da = ds[k]
        assert tuple(da.dims) == ("x", "y", "z") and np.asarray(da.data).shape == n
        assert np.asarray(da.data).dtype == dtype and np.asarray(da.data).flags.writeable
        assert da.attrs["units"] == units and da.attrs["long_name"] == long_name
        assert cases.rel_l2(da.data, want[k]) < TOL, k
    assert float(np.asarray(ds["p_min"].data).max()) > 0                          # PNP is stored sign-flipped (:136)
    # the device packaging (lifu_get_packaged) is the reference's numpy expression bit for bit
    from openlifu_b200.sim.kwave_if import package_fields
    host = package_fields(params, out["p_max"], out["p_min"])
    for k in ("p_max", "p_min", "intensity"):
        assert np.array_equal(np.asarray(ds[k].data), np.asarray(host[k].data)), k
    # defaults: delays -> zeros, apod -> ones (kwave_if.py:98-99)
    ds0, _ = run_simulation(arr=arr, params=params, freq=400e3, cycles=3, dt=setup.dt, t_end=setup.t_end)
    want0 = osc.run_simulation(_scene(arr, params, 1e5), freq=400e3, cycles=3, dt=setup.dt, t_end=setup.t_end)
    assert cases.rel_l2(ds0["p_max"].data, want0["p_max"]) < TOL


def test_run_simulation_heterogeneous_and_ref_values_only(lifu_lib, monkeypatch):
    monkeypatch.setenv("LIFU_PACKAGING", "device")
    from openlifu_b200.sim import run_simulation
    arr, setup, params = _setup(steps=80)
    shape = np.asarray(params["sound_speed"].data).shape
    c0, rho0, al = cases.layered_phantom(shape)
    params["sound_speed"].data[...] = c0
    params["density"].data[...] = rho0
    params["attenuation"].data[...] = al
    kw = dict(freq=400e3, cycles=2, dt=1.5e-7, t_end=80 * 1.5e-7)
    ds, out = run_simulation(arr=arr, params=params, **kw)
    want = osc.run_simulation(_scene(arr, params, 1e5), **kw)
    from openlifu_b200.sim.kwave_if import package_fields
    host = package_fields(params, out["p_max"], out["p_min"])
    for k in ("p_max", "p_min", "intensity"):
        assert cases.rel_l2(ds[k].data, want[k]) < TOL, k
        assert np.array_equal(np.asarray(ds[k].data), np.asarray(host[k].data)), k     # impedance map on the device
    # ref_values_only: the medium is the maps' ref_value scalars (kwave_if.py:52-56), the intensity still uses the maps
    ds_r, _ = run_simulation(arr=arr, params=params, ref_values_only=True, **kw)
    sc = _scene(arr, params, 1e5)
    sc.extras["ref_values"] = {k: params[k].attrs["ref_value"] for k in ("sound_speed", "density", "attenuation")}
    want_r = osc.run_simulation(sc, ref_values_only=True, **kw)
    for k in ("p_max", "p_min", "intensity"):
        assert cases.rel_l2(ds_r[k].data, want_r[k]) < TOL, k
    assert cases.rel_l2(ds_r["p_max"].data, ds["p_max"].data) > 1e-2                # really a different medium


def test_calc_solution_wheel_on_gpu(lifu_lib):
    """Per-focus loop, stacking, scaling, aggregation and both analysis engines on real solver output."""
    from openlifu_b200.bf import Pulse, Sequence, focal_patterns
    from openlifu_b200.geo import Point
    from openlifu_b200.plan import Protocol, SolutionAnalysisOptions
    from openlifu_b200.sim import SimSetup
    from openlifu_b200.xdc import Transducer
    arr = Transducer.gen_matrix_array(nx=4, ny=4, pitch=3, kerf=0.5, units="mm", sensitivity=1e5)
    setup = SimSetup(spacing=1.0, x_extent=(-15, 15), y_extent=(-15, 15), z_extent=(-3, 36), dt=2.5e-7, t_end=110 * 2.5e-7)
    pattern = focal_patterns.Wheel(center=True, num_spokes=2, spoke_radius=3, distance_units="mm", target_pressure=0.3, units="MPa")
    pr = Protocol(pulse=Pulse(frequency=400e3, duration=3 / 400e3), sequence=Sequence(pulse_interval=0.01, pulse_count=3, pulse_train_interval=0),
                  focal_pattern=pattern, sim_setup=setup)
    target = Point(position=np.array([0.0, 0.0, 22.0]), units="mm", id="tgt")
    opts = SolutionAnalysisOptions(mainlobe_radius=2.5, beamwidth_radius=5.0, sidelobe_radius=3.0, sidelobe_zmin=1.0, distance_units="mm")
    sol_raw, _, _ = pr.calc_solution(target, arr, simulate=True, scale=False, analysis_options=opts)
    res = sol_raw.simulation_result
    assert tuple(res["p_min"].dims) == ("focal_point_index", "x", "y", "z") and res["p_min"].data.shape[0] == 3
    params = setup.setup_sim_scene(pr.seg_method)
    sc = _scene(arr, params, 1e5)
    geometry = osc.source_geometry(sc)
    for i, focus in enumerate(sol_raw.foci):
        want = osc.run_simulation(sc, delays=sol_raw.delays[i], apod=sol_raw.apodizations[i], freq=400e3, cycles=3,
                                  dt=setup.dt, t_end=setup.t_end, geometry=geometry)
        for k in ("p_max", "p_min", "intensity"):
            assert cases.rel_l2(res[k].data[i], want[k]) < TOL, (i, k)
    # scaled solution: analysis engines agree, every focus reaches the target pressure, aggregation = max / mean
    sol, agg, ana = pr.calc_solution(target, arr, simulate=True, scale=True, analysis_options=opts)
    host = sol.analyze(options=opts, engine="host")
    dev = sol.analyze(options=opts, engine="cuda")
    for k, v in host.__dict__.items():
        if k == "param_constraints" or v is None:
            continue
        np.testing.assert_allclose(np.asarray(getattr(dev, k), dtype=float), np.asarray(v, dtype=float), rtol=1e-6,
                                   equal_nan=True, err_msg=k)
    np.testing.assert_allclose(ana.mainlobe_pnp_MPa, [0.3] * 3, rtol=1e-5)
    r = sol.simulation_result
    assert np.array_equal(agg["p_min"].data, np.max(r["p_min"].data, axis=0))
    np.testing.assert_allclose(agg["intensity"].data, np.mean(r["intensity"].data, axis=0), rtol=1e-12)


def _wheel_protocol(heterogeneous=False):
    from openlifu_b200.bf import Pulse, Sequence, focal_patterns
    from openlifu_b200.geo import Point
    from openlifu_b200.plan import Protocol, SolutionAnalysisOptions
    from openlifu_b200.sim import SimSetup
    from openlifu_b200.xdc import Transducer
    arr = Transducer.gen_matrix_array(nx=4, ny=4, pitch=3, kerf=0.5, units="mm", sensitivity=1e5)
    setup = SimSetup(spacing=1.0, x_extent=(-15, 15), y_extent=(-15, 15), z_extent=(-3, 36), dt=2.5e-7, t_end=110 * 2.5e-7)
    pattern = focal_patterns.Wheel(center=True, num_spokes=3, spoke_radius=3, distance_units="mm", target_pressure=0.3, units="MPa")
    pr = Protocol(pulse=Pulse(frequency=400e3, duration=3 / 400e3), sequence=Sequence(pulse_interval=0.01, pulse_count=4, pulse_train_interval=0),
                  focal_pattern=pattern, sim_setup=setup)
    target = Point(position=np.array([0.0, 0.0, 22.0]), units="mm", id="tgt")
    opts = SolutionAnalysisOptions(mainlobe_radius=2.5, beamwidth_radius=5.0, sidelobe_radius=3.0, sidelobe_zmin=1.0, distance_units="mm")
    volume = None
    if heterogeneous:
        from openlifu_b200 import configs
        pr.seg_method = configs.seg_methods.LabelVolume(materials=dict(configs.PHANTOM_MATERIALS), ref_material="water")
        volume = configs.skull_phantom_labels(setup.get_coords(), centre_mm=(0.0, 0.0, 40.0), r_in=24.0, r_out=28.0)
    return pr, arr, target, opts, volume


@pytest.mark.parametrize("heterogeneous", [False, True])
def test_plan_on_device_is_the_host_plan_bit_for_bit(lifu_lib, heterogeneous):
    """Protocol.calc_solution(on_device=True) keeps every focus' fields in HBM from the solver through
    stacking, Solution.scale, the aggregation over foci and both analyses (csrc/stack.cu) and copies the finished stack
    out once.  Every array and every metric equals the host route exactly."""
    pr, arr, target, opts, volume = _wheel_protocol(heterogeneous)
    sol_h, agg_h, ana_h = pr.calc_solution(target, arr, volume=volume, simulate=True, scale=True, analysis_options=opts,
                                           use_gpu=True, on_device=False)
    sol_d, agg_d, ana_d = pr.calc_solution(target, arr, volume=volume, simulate=True, scale=True, analysis_options=opts,
                                           use_gpu=True, on_device=True)
    assert getattr(sol_d, "_stack", None) is None                      # released
    rh, rd = sol_h.simulation_result, sol_d.simulation_result
    assert tuple(rd["p_min"].dims) == tuple(rh["p_min"].dims) == ("focal_point_index", "x", "y", "z")
    assert list(rd.coords) == list(rh.coords)
    for k in ("p_max", "p_min", "intensity"):
        assert rd[k].data.dtype == rh[k].data.dtype and rd[k].data.shape == rh[k].data.shape
        assert rd[k].attrs == rh[k].attrs and rd[k].name == rh[k].name
        assert np.array_equal(rd[k].data, rh[k].data), k
        assert agg_d[k].data.dtype == agg_h[k].data.dtype and np.array_equal(agg_d[k].data, agg_h[k].data), "aggregated " + k
        assert agg_d[k].attrs == agg_h[k].attrs
    assert sol_d.voltage == sol_h.voltage and np.array_equal(sol_d.apodizations, sol_h.apodizations)
    for k, v in ana_h.__dict__.items():
        if k == "param_constraints" or v is None:
            continue
        assert np.array_equal(np.asarray(getattr(ana_d, k), dtype=float), np.asarray(v, dtype=float), equal_nan=True), k
    # the returned arrays are ordinary writable numpy arrays (Solution.scale can be applied again, in place, on the host)
    rd["p_min"][1].data *= 2.0
    assert rd["p_min"].data[1].max() == 2.0 * rh["p_min"].data[1].max()
    # unscaled route too
    sol_u, _, _ = pr.calc_solution(target, arr, volume=volume, simulate=True, scale=False, analysis_options=opts,
                                   use_gpu=True, on_device=True)
    sol_v, _, _ = pr.calc_solution(target, arr, volume=volume, simulate=True, scale=False, analysis_options=opts,
                                   use_gpu=True, on_device=False)
    for k in ("p_max", "p_min", "intensity"):
        assert np.array_equal(sol_u.simulation_result[k].data, sol_v.simulation_result[k].data), k


def test_field_stack_errors_and_numpy_semantics(lifu_lib):
    """lifu_stack_*: argument / state errors, and the scaling and aggregation arithmetic against numpy on the same data."""
    from openlifu_b200 import _lib
    with pytest.raises(ValueError):
        _lib.FieldStack([0, 4, 4], 2)
    case = cases.small_water_case()
    N, d, Nt, dt = osc.time_axis(cases.scene_of(case), case["dt"], case["t_end"], 0.5)
    offset = np.array([-float(np.mean(c)) * 1e-3 for c in case["coords"]])
    base = 1e5 * np.sin(2 * np.pi * case["freq"] * np.arange(0, case["cycles"] / case["freq"], dt))
    with _lib.LifuSim(N, d, dt, Nt) as sim, _lib.FieldStack(N, 3) as st:
        sim.set_medium(1500.0, 1000.0, 0.0)
        sim.set_elements(case["pos_m"] + offset, case["size_m"], case["angles_deg"], 0.05, 5)
        with pytest.raises(_lib.LifuError, match="lifu_run first"):
            st.put(0, sim)
        host = []
        for f, g in enumerate(([1, 1, 1, 1], [1, 0.5, 0.2, 0], [0.3, 1, 1, 0.7])):
            sim.set_drive(base, [0, 1, 2, 3], g)
            p_max, p_min, _ = sim.run()
            if f == 0:
                with pytest.raises(_lib.LifuError, match="lifu_set_two_z"):
                    st.put(f, sim)
                sim.set_two_z(2 * 1000.0 * 1500.0)
            st.put(f, sim)
            pm = p_max.reshape(N, order="F")
            pn = (-1 * p_min).reshape(N, order="F")
            it = ((np.float32(1e-4) * p_min ** 2) / np.float64(2 * 1000.0 * 1500.0)).reshape(N, order="F")
            host.append([pm.copy(), pn.copy(), it.copy()])
            if f < 2:
                with pytest.raises(_lib.LifuError, match="no fields yet"):
                    st.aggregate()
        for f in range(3):
            got = st.get(f)
            for a, b in zip(got, host[f]):
                assert a.dtype == b.dtype and np.array_equal(a, b)
        s = np.float64(1.7320508075688772) / 3.0 * np.float64(1.1)
        st.scale(1, s)
        host[1][0] *= s
        host[1][1] *= s
        host[1][2] *= s ** 2
        pm, pn, it = st.get()
        for k, arr in enumerate((pm, pn, it)):
            want = np.stack([h[k] for h in host])
            assert arr.shape == want.shape and np.array_equal(arr, want), k
        a_pm, a_pn, a_it = st.aggregate()
        assert np.array_equal(a_pm, np.max(np.stack([h[0] for h in host]), axis=0))
        assert np.array_equal(a_pn, np.max(np.stack([h[1] for h in host]), axis=0))
        assert np.array_equal(a_it, np.mean(np.stack([h[2] for h in host]), axis=0))


def test_simulate_candidates_against_oracle(lifu_lib):
    """three candidate poses of the same array for one target, each an independent simulation with
    its own off-grid source weights; source masks bit-exact and fields within 1e-4 of the oracle run on the baked
    geometry; the analysis ranks the candidates."""
    pr, arr, target, opts, _ = _wheel_protocol()
    rng = np.random.default_rng(11)
    transforms = [np.eye(4)]
    for _ in range(2):
        ay, ax = rng.uniform(-0.15, 0.15, 2)
        ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        m = np.eye(4)
        m[:3, :3] = ry @ rx
        m[:3, 3] = rng.uniform(-2.5, 2.5, 3) * np.array([1, 1, 0.3])
        transforms.append(m)
    results = pr.simulate_candidates(target, arr, transforms, analysis_options=opts, use_gpu=True)
    assert len(results) == 3
    params = pr.sim_setup.setup_sim_scene(pr.seg_method)
    peaks = []
    for i, (sol, ana) in enumerate(results):
        res = sol.simulation_result
        assert tuple(res["p_min"].dims) == ("focal_point_index", "x", "y", "z") and res["p_min"].data.shape[0] == 1
        sc = _scene(sol.transducer, params, 1e5)
        want = osc.run_simulation(sc, delays=sol.delays[0], apod=sol.apodizations[0], freq=400e3, cycles=3,
                                  dt=pr.sim_setup.dt, t_end=pr.sim_setup.t_end)
        for k in ("p_max", "p_min", "intensity"):
            assert cases.rel_l2(res[k].data[0], want[k]) < TOL, (i, k)
        assert len(ana.mainlobe_pnp_MPa) == 1 and ana.mainlobe_pnp_MPa[0] > 0
        peaks.append(ana.mainlobe_pnp_MPa[0])
    # the identity pose reproduces the plain single-focus plan
    sol0, _, _ = pr.__class__(pulse=pr.pulse, sequence=pr.sequence, sim_setup=pr.sim_setup).calc_solution(
        target, arr, simulate=True, scale=False, analysis_options=opts, use_gpu=True)
    assert np.array_equal(results[0][0].simulation_result["p_min"].data, sol0.simulation_result["p_min"].data)
    assert len(set(np.round(peaks, 9))) == 3                                # the poses really differ


def test_label_medium_on_device_equals_the_expanded_maps(lifu_lib, monkeypatch):
    """the param maps of SegmentationMethod._map_params (seg_method.py:84-97) expanded on the DEVICE from
    the label volume and the per-material tables (lifu_set_medium_labels) give bit-identical fields to uploading the three
    float64 maps; run_simulation takes that route only while the maps still hold what _map_params wrote."""
    from openlifu_b200 import _lib, configs
    from openlifu_b200.sim import SimSetup, kwave_if
    from openlifu_b200.xdc import Transducer
    setup = SimSetup(spacing=1.0, x_extent=(-15, 15), y_extent=(-15, 15), z_extent=(-3, 36), dt=1.5e-7, t_end=90 * 1.5e-7)
    seg = configs.seg_methods.LabelVolume(materials=dict(configs.PHANTOM_MATERIALS), ref_material="water")
    volume = configs.skull_phantom_labels(setup.get_coords(), centre_mm=(0.0, 0.0, 40.0), r_in=24.0, r_out=28.0)
    params = setup.setup_sim_scene(seg, volume=volume)
    lm = params.attrs["lifu_label_medium"]
    assert lm["labels"].dtype == np.uint8 and set(np.unique(lm["labels"])) == {0, 1, 2}
    arr = Transducer.gen_matrix_array(nx=3, ny=3, pitch=4, kerf=0.5, units="mm", sensitivity=1e5)
    calls = {"labels": 0, "maps": 0}
    orig_l, orig_m = _lib.LifuSim.set_medium_labels, _lib.LifuSim.set_medium

    def spy_l(self, *a, **k):
        calls["labels"] += 1
        return orig_l(self, *a, **k)

    def spy_m(self, *a, **k):
        calls["maps"] += 1
        return orig_m(self, *a, **k)

    monkeypatch.setattr(_lib.LifuSim, "set_medium_labels", spy_l)
    monkeypatch.setattr(_lib.LifuSim, "set_medium", spy_m)
    kw = dict(arr=arr, params=params, freq=400e3, cycles=3, dt=setup.dt, t_end=setup.t_end, amplitude=1.0, gpu=True)
    kwave_if.clear_sessions()
    ds_l, out_l = kwave_if.run_simulation(**kw)
    assert calls == {"labels": 1, "maps": 0} and out_l["stats"]["homogeneous"] == 0 and out_l["stats"]["absorbing"] == 1
    monkeypatch.setenv("LIFU_MEDIUM_LABELS", "0")
    kwave_if.clear_sessions()
    ds_m, _ = kwave_if.run_simulation(**kw)
    assert calls == {"labels": 1, "maps": 1}
    monkeypatch.delenv("LIFU_MEDIUM_LABELS")
    for k in ("p_max", "p_min", "intensity"):
        assert np.array_equal(ds_l[k].data, ds_m[k].data), k
    # an in-place edit of one voxel of a map: the label route must not be taken any more, and the edit must be simulated
    params["sound_speed"].data[15, 15, 20] = 1700.0
    ds_e, _ = kwave_if.run_simulation(**kw)
    assert calls == {"labels": 1, "maps": 2}
    assert not np.array_equal(ds_e["p_min"].data, ds_m["p_min"].data)
    kwave_if.clear_sessions()
    # bad tables / labels are refused
    with _lib.LifuSim([8, 8, 8], [1e-3] * 3, 1e-7, 2) as sim:
        with pytest.raises(ValueError):
            sim.set_medium_labels(np.zeros((8, 8, 8), dtype=np.uint8), np.ones(40), np.ones(40))
        with pytest.raises(ValueError, match="sound speed must be positive"):
            sim.set_medium_labels(np.full((8, 8, 8), 3, dtype=np.uint8), [1500.0, 1600.0], [1000.0, 1100.0])
