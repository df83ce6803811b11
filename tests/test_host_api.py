"""Host-side mirror of the reference interface: signatures, defaults, error behaviour (no GPU)."""
from __future__ import annotations

import inspect

import numpy as np
import pytest

import openlifu_b200 as ol
from openlifu_b200 import xa
from openlifu_b200.sim import kwave_if


def test_run_simulation_signature_matches_reference():
    """kwave_if.py:80-94: names, order and defaults of the boundary."""
    sig = inspect.signature(kwave_if.run_simulation)
    got = [(p.name, p.default) for p in sig.parameters.values()]
    want = [("arr", inspect._empty), ("params", inspect._empty), ("delays", None), ("apod", None), ("freq", 1e6),
            ("cycles", 20), ("amplitude", 1), ("dt", 0), ("t_end", 0), ("cfl", 0.5), ("bli_tolerance", 0.05),
            ("upsampling_rate", 5), ("gpu", True), ("ref_values_only", False)]
    assert got == want


def _scene():
    arr = ol.Transducer.gen_matrix_array(nx=2, ny=2, pitch=2, kerf=.5, units="mm", sensitivity=1e5)
    ss = ol.SimSetup(dt=2e-7, t_end=3 * 2e-7, x_extent=(-10, 10), y_extent=(-10, 10), z_extent=(-2, 10))
    params = ol.seg_methods.UniformWater().ref_params(ss.get_coords())
    return arr, ss, params


def test_no_cpu_path():
    arr, ss, params = _scene()
    with pytest.raises(RuntimeError, match="no CPU path"):
        kwave_if.run_simulation(arr=arr, params=params, gpu=False)


def test_unit_mismatch_errors():
    arr, ss, params = _scene()
    params.coords["z"].attrs["units"] = "m"
    with pytest.raises(ValueError, match="All coordinates must have the same units"):
        kwave_if.run_simulation(arr=arr, params=params)


def test_get_kgrid_time_axis():
    arr, ss, params = _scene()
    kg = kwave_if.get_kgrid(params.coords, dt=ss.dt, t_end=ss.t_end)
    assert kg["N"] == [21, 21, 13] and kg["Nt"] == 3 and kg["dt"] == 2e-7       # tests/test_sim.py:22-29 grid
    kg = kwave_if.get_kgrid(params.coords)
    assert kg["Nt"] == int(np.floor(np.sqrt(21 ** 2 * 2 + 13 ** 2) * 1e-3 / 1500 / (0.5e-3 / 1500))) + 1


def test_simsetup_validation_and_dict_roundtrip():
    with pytest.raises(ValueError, match="x_extent must be in the form"):
        ol.SimSetup(x_extent=(3, 1))
    with pytest.raises(ValueError, match="units must be a length unit"):
        ol.SimSetup(units="s")
    with pytest.raises(TypeError, match="spacing must be a number"):
        ol.SimSetup(spacing="1")
    ss = ol.SimSetup(spacing=0.5, x_extent=(-5, 5))
    ss2 = ol.SimSetup.from_dict(ss.to_dict())
    assert ss2 == ss
    assert ss.get_size().tolist() == [21, 121, 129]


def test_method_dict_roundtrips():
    for obj in (ol.delay_methods.Direct(c0=1540), ol.apod_methods.MaxAngle(25.0), ol.apod_methods.PiecewiseLinear(80, 30),
                ol.apod_methods.Uniform(0.5)):
        base = ol.DelayMethod if isinstance(obj, ol.DelayMethod) else ol.ApodizationMethod
        assert base.from_dict(obj.to_dict()) == obj
    w = ol.focal_patterns.Wheel(num_spokes=6, spoke_radius=3.0)
    assert ol.FocalPattern.from_dict(w.to_dict()) == w and w.num_foci() == 7
    with pytest.raises(ValueError):
        ol.apod_methods.PiecewiseLinear(zero_angle=10, rolloff_angle=20)
    with pytest.raises(TypeError):
        ol.apod_methods.MaxAngle("x")
    sm = ol.seg_methods.UniformWater()
    assert type(ol.SegmentationMethod.from_dict(sm.to_dict())) is type(sm)


def test_transducer_reference_kats():
    """tests/test_transducer.py:39-70 of the reference: convert_transform and effective origin."""
    arr = ol.Transducer.gen_matrix_array(nx=3, ny=2, units="cm")
    m = np.eye(4)
    m[:3, 3] = [1.0, 2.0, 3.0]
    out = arr.convert_transform(m, units="mm")
    assert np.allclose(out[:3, 3], [0.1, 0.2, 0.3]) and np.allclose(out[:3, :3], np.eye(3))
    apod = np.zeros(6)
    apod[[1, 4]] = 1
    want = arr.get_positions()[[1, 4]].mean(axis=0)
    assert np.allclose(arr.get_effective_origin(apod), want)
    assert np.allclose(arr.get_effective_origin(apod, units="mm"), want * 10)


def test_xa_shim_surface():
    coords = xa.Coordinates({"x": np.arange(3.0), "y": np.arange(2.0)})
    coords["x"].attrs["units"] = "mm"
    a = xa.DataArray(np.arange(6.0).reshape(3, 2), coords=coords, dims=("x", "y"), attrs={"units": "Pa"})
    assert a.dims == ("x", "y") and a.sizes == {"x": 3, "y": 2}
    assert float(a.max()) == 5 and float(a.where(a > 2).max()) == 5 and np.isnan(a.where(a > 2).data[0, 0])
    assert a.isel(x=1).dims == ("y",) and float(a.sel(x=2.0, y=1.0)) == 5
    stacked = xa.concat([xa.Dataset({"p": a}).assign_coords(focal_point_index=i) for i in range(2)], dim="focal_point_index")
    assert stacked["p"].dims == ("focal_point_index", "x", "y")
    view = stacked["p"][1]
    view.data *= 2                                   # writes through (solution.py:334-336 relies on it)
    assert stacked["p"].data[1, 2, 1] == 10 and stacked["p"].data[0, 2, 1] == 5
    agg = stacked["p"].max(dim="focal_point_index", keep_attrs=True)
    assert agg.attrs["units"] == "Pa" and agg.dims == ("x", "y")
    ds = stacked.drop_dims("focal_point_index")
    assert len(ds.data_vars) == 0 and "x" in ds.coords
    line = a.interp(x=xa.DataArray([0.5, 1.5], coords={"s": [0.0, 1.0]}), y=xa.DataArray([0.5, 0.5], coords={"s": [0.0, 1.0]}))
    assert line.dims == ("s",) and np.allclose(line.data, [1.5, 3.5])
    assert (a * xa.DataArray(np.array([1.0, 2.0]), dims=("y",))).shape == (3, 2)


def test_package_fields_matches_reference_expression_bitwise():
    """The packaging step (kwave_if.py:131-146) is evaluated with threaded flat-vector ops; values must equal the
    reference's literal numpy expression bit for bit, dtypes and names included."""
    from openlifu_b200 import configs
    from openlifu_b200.sim import kwave_if
    params = configs.prepare(configs.c3(24))[0]
    sz = tuple(params.coords.sizes.values())
    rng = np.random.default_rng(147)
    p_max = (rng.random(int(np.prod(sz))) * 3e5).astype(np.float32)
    p_min = (-rng.random(int(np.prod(sz))) * 3e5).astype(np.float32)
    ds = kwave_if.package_fields(params, p_max, p_min)
    Z = params["density"].data * params["sound_speed"].data
    assert np.array_equal(ds["p_max"].data, p_max.reshape(sz, order="F")) and ds["p_max"].data.dtype == np.float32
    assert np.array_equal(ds["p_min"].data, -1 * p_min.reshape(sz, order="F")) and ds["p_min"].data.dtype == np.float32
    want = 1e-4 * p_min.reshape(sz, order="F") ** 2 / (2 * Z)
    assert ds["intensity"].data.dtype == np.float64 and np.array_equal(ds["intensity"].data, want)
    assert ds["p_min"].attrs == {"units": "Pa", "long_name": "PNP"} and ds["intensity"].attrs["units"] == "W/cm^2"
    ds["p_min"].data *= 2.0          # Solution.scale works in place: the arrays must be writable


def test_analysis_engine_selection(monkeypatch):
    """Solution.analyze(engine=...): explicit value > $LIFU_ANALYZE > auto (device only for the solver's float32 stack)."""
    from openlifu_b200.plan.solution import _pick_engine
    from openlifu_b200 import xa
    f32 = xa.Dataset({"p_min": xa.DataArray(np.zeros((1, 2, 2, 2), np.float32), dims=("focal_point_index", "x", "y", "z"))})
    f64 = xa.Dataset({"p_min": xa.DataArray(np.zeros((1, 2, 2, 2)), dims=("focal_point_index", "x", "y", "z"))})
    monkeypatch.delenv("LIFU_ANALYZE", raising=False)
    assert _pick_engine("host", f32) == "host" and _pick_engine("cuda", f64) == "cuda"
    assert _pick_engine(None, f64) == "host"                     # not the solver's dtype -> numpy evaluation
    monkeypatch.setenv("LIFU_ANALYZE", "host")
    assert _pick_engine(None, f32) == "host"
    monkeypatch.setenv("LIFU_ANALYZE", "gpu")
    with pytest.raises(ValueError, match="Unknown analysis engine"):
        _pick_engine(None, f32)


def test_output_dict_derives_raw_p_min_on_demand():
    from openlifu_b200.sim.kwave_if import _Output
    out = _Output({"pnp": np.array([1.0, -2.0], np.float32)})
    assert "p_min" not in out
    assert np.array_equal(out["p_min"], np.array([-1.0, 2.0], np.float32)) and "p_min" in out
    with pytest.raises(KeyError):
        out["nope"]


def test_stack_of_fortran_ordered_foci_keeps_layout_and_roundtrips(tmp_path):
    """xa.concat of per-focus Datasets whose arrays are x-fastest (what run_simulation returns): the stack keeps each
    focus' memory order (block copies, and the layout the device analysis stages directly), equals np.stack, reduces
    and serialises like any other array."""
    from openlifu_b200 import xa
    n = (7, 6, 5)
    coords = {d: xa.DataArray(np.linspace(0, 1, k), dims=[d], attrs={"units": "mm"}) for d, k in zip("xyz", n)}
    rng = np.random.default_rng(0)
    fields = [rng.random(n).astype(np.float32) for _ in range(3)]
    pieces = []
    for i, f in enumerate(fields):
        ds = xa.Dataset({"p_min": xa.DataArray(np.asfortranarray(f), coords=coords, dims=("x", "y", "z"), attrs={"units": "Pa"}),
                         "intensity": xa.DataArray(np.asfortranarray(f.astype(np.float64)), coords=coords, dims=("x", "y", "z"),
                                                   attrs={"units": "W/cm^2"})})
        pieces.append(ds.assign_coords(focal_point_index=i))
    st = xa.concat(pieces, dim="focal_point_index")
    data = np.asarray(st["p_min"].data)
    assert data.shape == (3,) + n and np.array_equal(data, np.stack(fields))
    if not hasattr(xa, "__version__"):                       # the shim (real xarray makes its own C-ordered copy)
        assert all(data[i].flags.f_contiguous for i in range(3))
    assert np.array_equal(np.asarray(st["p_min"].max(dim="focal_point_index").data), np.max(np.stack(fields), axis=0))
    f = tmp_path / "stack.nc"
    st.to_netcdf(f, engine="scipy")
    back = xa.open_dataset(f, engine="scipy")
    for k in ("p_min", "intensity"):
        assert np.array_equal(np.asarray(back[k].data), np.asarray(st[k].data)), k


def test_medium_cache_key_sees_a_single_voxel_edit():
    """the medium cache key must change when ONE voxel of a params map is edited in place between two
    run_simulation calls (the reference rebuilds the medium on every call, kwave_if.py:113)."""
    from openlifu_b200.sim.kwave_if import _content_key
    rng = np.random.default_rng(3)
    a = 1500.0 + rng.random((48, 40, 36))
    k0 = _content_key(a)
    assert _content_key(a) == k0 and _content_key(a.copy()) == k0           # content, not identity
    for idx in [(0, 0, 1), (17, 23, 5), (47, 39, 35)]:                       # none of them on a coarse sampling lattice
        b = a.copy()
        b[idx] += 1e-9
        assert _content_key(b) != k0, idx
    f = np.asfortranarray(a)
    assert _content_key(f) != None and _content_key(np.asfortranarray(a)) == _content_key(f)  # noqa: E711
    g = f.copy(order="F")
    g[3, 4, 5] = 0.0
    assert _content_key(g) != _content_key(f)
    s = a[::2]                                                              # non-contiguous view: hashed byte by byte
    t = s.copy()
    t[1, 1, 1] += 1.0
    assert _content_key(s) != _content_key(t)
    assert _content_key(a.astype(np.float32)) != _content_key(a)


def test_per_element_sensitivity_is_applied_per_element_and_announced(caplog):
    """intended (non-compounding) behaviour for arrays whose elements carry their own sensitivity, i.e. after
    Transducer.merge of modules with different sensitivities -- each element's signal is scaled by ITS sensitivity; the
    reference compounds them through an in-place multiplication of the shared input (element.py:145-153).  The
    deviation is logged once at WARNING level."""
    import logging
    from openlifu_b200.xdc import Transducer
    from openlifu_b200.xdc import transducer as tmod
    arr = Transducer.gen_matrix_array(nx=2, ny=2, pitch=4, kerf=0.5, units="mm", sensitivity=2.0)
    for k, el in enumerate(arr.elements):
        el.sensitivity = 1.0 + k
    tmod._WARNED_ELEMENT_SENSITIVITY = False
    sig = np.sin(np.linspace(0, 6, 20))
    with caplog.at_level(logging.WARNING):
        out = arr.calc_output(sig, 1e-7, delays=np.zeros(4), apod=np.ones(4))
    for k in range(4):
        assert np.allclose(out[k, :20], (1.0 + k) * 2.0 * sig, rtol=1e-14)
    assert any("per-element sensitivities" in r.message for r in caplog.records)


def test_gpu_probe_falls_back_to_the_library(monkeypatch):
    """without pynvml the probe asks liblifusim (lifu_device_count) instead of reporting "no GPU"."""
    from openlifu_b200.util import checkgpu
    checkgpu._count.cache_clear()
    monkeypatch.setattr(checkgpu, "_nvml_device_count", lambda: (_ for _ in ()).throw(ImportError("no pynvml")))
    monkeypatch.setattr(checkgpu, "_library_device_count", lambda: 3)
    assert checkgpu.gpu_available() is True
    checkgpu._count.cache_clear()
    monkeypatch.setattr(checkgpu, "_library_device_count", lambda: 0)
    assert checkgpu.gpu_available() is False
    checkgpu._count.cache_clear()


def test_calc_solution_default_gpu_choice_names_the_missing_cpu_path(monkeypatch):
    """use_gpu=None on a box without a B200: an explicit error instead of the reference's CPU fallback."""
    from openlifu_b200 import configs
    from openlifu_b200.plan import Protocol
    from openlifu_b200.plan import protocol as pmod
    monkeypatch.setattr(pmod, "gpu_available", lambda: False)
    cfg = configs.c1()
    proto = Protocol(pulse=cfg["pulse"], sim_setup=cfg["setup"])
    with pytest.raises(RuntimeError, match="no CPU simulation path"):
        proto.calc_solution(cfg["target"], cfg["arr"])
    sol, agg, ana = proto.calc_solution(cfg["target"], cfg["arr"], simulate=False, scale=False)
    assert agg is None and ana is None and sol.delays.shape == (1, 64)
