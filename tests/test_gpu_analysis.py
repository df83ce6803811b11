"""Device beam analysis (the O(V) passes of Solution.analyze, csrc/analysis.cu) against
(i) the golden vectors the real reference produced (tests/golden/ref_plan.json), (ii) the host numpy evaluation
of the same solution -- bit-exact for maxima, selection sizes and line samples, 1e-6 for the float32-weighted
centroid -- and (iii) size-independent properties at the C2 grid size (216^3)."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from openlifu_b200 import _lib, xa
from openlifu_b200.bf import Pulse, Sequence
from openlifu_b200.bf.focal_patterns import Wheel
from openlifu_b200.geo import Point
from openlifu_b200.plan import Solution, SolutionAnalysisOptions
from openlifu_b200.plan.solution_analysis import FocusFrame, trilinear_line
from openlifu_b200.xdc import Transducer

GOLD = Path(__file__).parent / "golden"
CENTROID = ("focal_centroid_lat_mm", "focal_centroid_ele_mm", "focal_centroid_ax_mm")


def _mods():
    return {"xa": xa, "Transducer": Transducer, "Point": Point, "Solution": Solution, "Pulse": Pulse, "Sequence": Sequence,
            "SolutionAnalysisOptions": SolutionAnalysisOptions, "Wheel": Wheel}


def _golden_case():
    import sys
    sys.path.insert(0, str(GOLD))
    from make_reference_plan_goldens import synthetic_case
    return synthetic_case(_mods())


def _plain(a):
    return {k: v for k, v in a.__dict__.items() if k != "param_constraints"}


def _assert_same(got, want, centroid_rtol=1e-6):
    g, w = _plain(got), _plain(want)
    assert g.keys() == w.keys()
    for k in w:
        if w[k] is None:
            assert g[k] is None, k
        elif k in CENTROID:
            np.testing.assert_allclose(g[k], w[k], rtol=centroid_rtol, equal_nan=True, err_msg=k)
        else:
            assert np.array_equal(np.asarray(g[k], dtype=float), np.asarray(w[k], dtype=float), equal_nan=True), \
                f"{k}: {g[k]} != {w[k]}"


def _random_solution(shape=(37, 29, 45), n_foci=3, seed=5, nan_frac=0.0, order="C"):
    rng = np.random.default_rng(seed)
    x = np.linspace(-18, 18, shape[0]); y = np.linspace(-14, 14, shape[1]); z = np.linspace(-2.0, 64, shape[2])
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    foci_mm = [(3.0 * np.cos(2.1 * i) + 1.0, 4.0 * np.sin(1.3 * i), 35.0 + 6 * i) for i in range(n_foci)]
    fields = []
    for i, (fx, fy, fz) in enumerate(foci_mm):
        amp = 1e6 * (0.7 + 0.2 * i)
        f = amp * np.exp(-(((X - fx) / 1.7) ** 2 + ((Y - fy) / 1.5) ** 2 + ((Z - fz) / 7.0) ** 2))
        f += 0.3 * amp * np.exp(-(((X - fx + 8) / 2.0) ** 2 + ((Y - fy - 3) / 2.0) ** 2 + ((Z - fz + 5) / 4.0) ** 2))
        f += 2e3 * rng.random(f.shape)
        fields.append(f)
    p_min = np.stack(fields).astype(np.float32)
    if nan_frac > 0:
        p_min[rng.random(p_min.shape) < nan_frac] = np.nan
    inten = 1e-4 * p_min.astype(np.float64) ** 2 / (2 * 1000.0 * 1500.0)
    if order == "F":       # what run_simulation returns per focus: x fastest
        p_min = np.ascontiguousarray(p_min.transpose(0, 3, 2, 1)).transpose(0, 3, 2, 1)
        inten = np.ascontiguousarray(inten.transpose(0, 3, 2, 1)).transpose(0, 3, 2, 1)
        assert p_min[0].flags.f_contiguous
    coords = {"x": xa.DataArray(x, dims=["x"], attrs={"units": "mm", "long_name": "Lateral"}),
              "y": xa.DataArray(y, dims=["y"], attrs={"units": "mm", "long_name": "Elevation"}),
              "z": xa.DataArray(z, dims=["z"], attrs={"units": "mm", "long_name": "Axial"}),
              "focal_point_index": list(range(n_foci))}
    dims = ["focal_point_index", "x", "y", "z"]
    ds = xa.Dataset({"p_min": xa.DataArray(p_min, dims=dims, attrs={"units": "Pa", "long_name": "PNP"}),
                     "p_max": xa.DataArray((1.1 * p_min).astype(np.float32), dims=dims, attrs={"units": "Pa", "long_name": "PPP"}),
                     "intensity": xa.DataArray(inten, dims=dims, attrs={"units": "W/cm^2", "long_name": "Intensity"})},
                    coords=coords)
    arr = Transducer.gen_matrix_array(nx=4, ny=4, pitch=4, kerf=0.5, units="mm")
    foci = [Point(position=np.array(f), units="mm", id=f"f{i}") for i, f in enumerate(foci_mm)]
    sol = Solution(id="r", transducer=arr, delays=np.zeros((n_foci, 16)), apodizations=np.ones((n_foci, 16)),
                   pulse=Pulse(frequency=400e3, amplitude=1.0, duration=50e-6),
                   sequence=Sequence(pulse_interval=0.01, pulse_count=2 * n_foci, pulse_train_interval=0.1, pulse_train_count=3),
                   voltage=7.0, foci=foci, target=foci[0], simulation_result=ds)
    return sol


@pytest.mark.gpu
def test_cuda_analysis_matches_reference_golden():
    gold = json.loads((GOLD / "ref_plan.json").read_text())
    sol, opts, _ = _golden_case()
    got = sol.analyze(options=opts, engine="cuda")
    for k, w in gold["analysis"].items():
        g = getattr(got, k)
        if w is None:
            assert g is None, k
        else:
            np.testing.assert_allclose(np.asarray(g, dtype=float), np.asarray(w, dtype=float),
                                       rtol=1e-6 if k in CENTROID else 1e-9, atol=1e-12, equal_nan=True, err_msg=k)


@pytest.mark.gpu
def test_cuda_scale_matches_reference_golden():
    gold = json.loads((GOLD / "ref_plan.json").read_text())
    sol, opts, pattern = _golden_case()
    apod_f, v0, v1 = sol.compute_scaling_factors(pattern, sol.analyze(options=opts, engine="cuda"))
    np.testing.assert_allclose(apod_f, gold["scaling"]["apod_factors"], rtol=1e-9)
    np.testing.assert_allclose([v0, v1], [gold["scaling"]["v0"], gold["scaling"]["v1"]], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("units", ["mm", "m"])
@pytest.mark.parametrize("order,nan_frac,n_foci", [("C", 0.0, 3), ("F", 0.0, 2), ("C", 0.02, 3), ("C", 0.0, 1)])
def test_cuda_analysis_equals_host(order, nan_frac, n_foci, units):
    sol = _random_solution(order=order, nan_frac=nan_frac, n_foci=n_foci)
    if units == "mm":
        opts = SolutionAnalysisOptions(mainlobe_radius=2.5, beamwidth_radius=5.0, sidelobe_radius=3.0, sidelobe_zmin=1.0,
                                       distance_units="mm")
    else:
        opts = SolutionAnalysisOptions()
    _assert_same(sol.analyze(options=opts, engine="cuda"), sol.analyze(options=opts, engine="host"))


@pytest.mark.gpu
def test_empty_selections_are_nan():
    """Focus outside the grid: the main lobe selects nothing -> NaN maxima, NaN centroid, NaN beam widths."""
    sol = _random_solution(n_foci=1)
    sol.foci = [Point(position=np.array([0.0, 0.0, 200.0]), units="mm", id="far")]
    opts = SolutionAnalysisOptions(mainlobe_radius=2.5, beamwidth_radius=5.0, sidelobe_radius=3.0, sidelobe_zmin=1.0,
                                   distance_units="mm")
    got = sol.analyze(options=opts, engine="cuda")
    assert np.isnan(got.mainlobe_pnp_MPa[0]) and np.isnan(got.focal_centroid_ax_mm[0]) and np.isnan(got.beamwidth_ax_3dB_mm[0])
    _assert_same(got, sol.analyze(options=opts, engine="host"))


@pytest.mark.gpu
def test_selection_sizes_and_lines_bit_exact():
    """The C-ABI handle directly: selection sizes == numpy mask sums, line samples == the host trilinear gather,
    and the result does not depend on the memory layout of the staged fields."""
    sol = _random_solution(shape=(40, 33, 51), n_foci=2)
    p = np.asarray(sol.simulation_result["p_min"].data)
    I = np.asarray(sol.simulation_result["intensity"].data)
    axes = [np.asarray(sol.simulation_result["p_min"].coords[d].data, dtype=np.float64) for d in ("x", "y", "z")]
    z_ok = axes[2] > 1.0
    frame = FocusFrame(sol.foci[1].get_position(units="mm"), np.array([0.5, -0.25, 0.0]))
    ar = (1.0, 1.0, 5.0)
    dist = frame.distance(axes, ar)
    main, side = dist < 2.5, (dist > 3.0) & z_ok[None, None, :]
    offs = [np.linspace(-5.0 * s, 5.0 * s, 2 * n) for s, n in zip(ar, p.shape[1:])]
    pts = [frame.line(k, o) for k, o in enumerate(offs)]
    results = []
    for order in ("C", "F"):
        with _lib.BeamAnalysis(axes, 2, z_ok=z_ok) as ana:
            for f in range(2):
                ana.set_focus(f, np.asarray(p[f], order=order), np.asarray(I[f], order=order))
            m, lines = ana.run_focus(1, frame.inverse, ar, 2.5, 3.0, np.float32(1e-6), line_pts=pts)
        results.append((m, lines))
        assert m["n_main"] == int(main.sum()) and m["n_side"] == int(side.sum()) and m["n_global"] == int(z_ok.sum()) * 40 * 33
        pm = (p[1] * np.float32(1e-6)).astype(np.float64)
        assert m["main_pnp"] == pm[main].max() and m["side_pnp"] == pm[side].max()
        assert m["main_ipa"] == I[1][main].max() and m["main_ipa_all"] == I[:, main].max()
        assert m["global_ipa_all"] == I[..., z_ok].max()
        for k in range(3):
            want = trilinear_line(pm, axes, pts[k])
            assert np.array_equal(lines[k], want, equal_nan=True)
        cut = np.float32(m["main_pnp"] * 10 ** (-3 / 20))
        sel = main & ((p[1] * np.float32(1e-6)) > cut)
        assert m["n_centroid"] == int(sel.sum())
        np.testing.assert_allclose(m["cen_w"], pm[sel].sum(), rtol=1e-12)
    for key in results[0][0]:
        if key not in ("kernel_ms", "reduce_ms"):
            np.testing.assert_allclose(results[0][0][key], results[1][0][key], rtol=1e-13, err_msg=key)


@pytest.mark.gpu
def test_full_size_properties_c2():
    """216^3 (the C2 inner grid), 2 foci: doubling the pressure doubles every pressure metric exactly and leaves
    the selections unchanged; the main-lobe size matches the ellipsoid volume; staging + analysis time is printed."""
    n = 216
    ax = [np.linspace(-53.75, 53.75, n), np.linspace(-53.75, 53.75, n), np.linspace(-4.0, 103.5, n)]
    X = ax[0][:, None, None]; Y = ax[1][None, :, None]; Z = ax[2][None, None, :]
    rng = np.random.default_rng(11)
    foci = [(0.0, 0.0, 50.0), (4.0, -3.0, 55.0)]
    p = np.stack([(1e6 * np.exp(-((X - fx) / 2.0) ** 2 - ((Y - fy) / 2.0) ** 2 - ((Z - fz) / 9.0) ** 2)).astype(np.float32)
                  for fx, fy, fz in foci])
    p += (1e3 * rng.random(p.shape)).astype(np.float32)
    I = 1e-4 * p.astype(np.float64) ** 2 / 3e6
    z_ok = ax[2] > 1.0
    out = []
    import time
    for scale in (1.0, 2.0):
        t0 = time.perf_counter()
        with _lib.BeamAnalysis(ax, 2, z_ok=z_ok) as ana:
            for f in range(2):
                ana.set_focus(f, p[f] * np.float32(scale), I[f])
            t1 = time.perf_counter()
            ms = [ana.run_focus(f, FocusFrame(np.array(foci[f]), np.zeros(3)).inverse, (1, 1, 5), 2.5, 3.0, np.float32(1e-6))[0]
                  for f in range(2)]
            t2 = time.perf_counter()
        out.append(ms)
        print(f"216^3 x 2 foci: staging {1e3 * (t1 - t0):.1f} ms, analysis {1e3 * (t2 - t1):.2f} ms wall, "
              f"kernels {ms[0]['kernel_ms']:.3f} + {ms[1]['kernel_ms']:.3f} ms, k_focus_reduce {ms[0]['reduce_ms']:.4f} + "
              f"{ms[1]['reduce_ms']:.4f} ms = {20 * n ** 3 / ms[1]['reduce_ms'] / 1e6:.0f} GB/s")
    for f in range(2):
        a, b = out[0][f], out[1][f]
        for k in ("main_pnp", "side_pnp", "global_pnp", "cen_w", "cen_wx", "cen_wy", "cen_wz"):
            assert b[k] == 2.0 * a[k], k
        for k in ("n_main", "n_side", "n_global", "n_centroid", "main_ipa", "main_ipa_all"):
            assert b[k] == a[k], k
        voxel = (107.5 / 215) ** 3
        vol = 4.0 / 3.0 * np.pi * 2.5 * 2.5 * 12.5
        assert abs(a["n_main"] * voxel / vol - 1.0) < 0.02
        assert a["n_global"] == int(z_ok.sum()) * n * n


@pytest.mark.gpu
@pytest.mark.parametrize("aspect", [(1.0, 1.0, 1.0), (1.0, 1.0, 5.0), (3.0, 0.7, 1.3)])
def test_voxels_exactly_on_the_radius(aspect):
    """Axis-aligned focus frame on a 0.5 mm lattice: many voxels sit exactly on the 2.5 mm radius (3-4-5 triples), where
    `dist < r` / `dist > r` must come out as numpy computes them (the kernel's reciprocal pre-test may not decide these)."""
    ax = [np.arange(-20, 21) * 0.5, np.arange(-18, 19) * 0.5, 20.0 + np.arange(0, 61) * 0.5]
    rng = np.random.default_rng(3)
    p = (1e6 * rng.random((41, 37, 61))).astype(np.float32)
    I = p.astype(np.float64) ** 2 * 1e-10
    frame = FocusFrame(np.array([0.0, 0.0, 30.0]), np.zeros(3))
    dist = frame.distance(ax, aspect)
    on = int((dist == 2.5).sum())
    if aspect == (1.0, 1.0, 1.0):
        assert on >= 30
    main, side = dist < 2.5, dist > 2.5
    with _lib.BeamAnalysis(ax, 1) as ana:
        ana.set_focus(0, p, I)
        m, _ = ana.run_focus(0, frame.inverse, aspect, 2.5, 2.5, np.float32(1e-6))
    assert m["n_main"] == int(main.sum()) and m["n_side"] == int(side.sum())
    assert m["n_main"] + m["n_side"] + on == p.size
    pm = (p * np.float32(1e-6)).astype(np.float64)
    assert m["main_pnp"] == pm[main].max() and m["side_pnp"] == pm[side].max() and m["side_ipa"] == I[side].max()
