#!/usr/bin/env python
"""Golden vectors for the planning layer (Solution.analyze / scale / aggregation), produced by running
the REAL reference classes (/root/reference/src/openlifu/plan/solution.py) on the xarray shim in the
build container.  Same import arrangement as make_reference_goldens.py.  Output: ref_plan.json (committed).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_reference_goldens import install_reference  # noqa: E402


def synthetic_case(mod):
    """A two-focus solution on a 1 mm grid with Gaussian main lobes and one side lobe each.
    ``mod`` supplies the classes (reference or ours) so both sides build identical inputs."""
    rng = np.random.default_rng(147)
    x = np.linspace(-15, 15, 31); y = np.linspace(-12, 12, 25); z = np.linspace(20, 70, 51)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    foci_mm = [(2.0, -1.0, 45.0), (-3.0, 2.0, 50.0)]
    fields = []
    for (fx, fy, fz), amp in zip(foci_mm, (1.3e6, 0.9e6)):
        f = amp * np.exp(-(((X - fx) / 2.0) ** 2 + ((Y - fy) / 1.8) ** 2 + ((Z - fz) / 8.0) ** 2))
        f += 0.25 * amp * np.exp(-(((X - fx - 9) / 2.0) ** 2 + ((Y - fy) / 2.0) ** 2 + ((Z - fz + 4) / 4.0) ** 2))
        f += 1e3 * rng.random(f.shape)
        fields.append(f)
    p_min = np.stack(fields).astype(np.float32)
    p_max = (1.05 * p_min).astype(np.float32)
    inten = 1e-4 * p_min.astype(np.float64) ** 2 / (2 * 1000.0 * 1500.0)
    xa = mod["xa"]
    coords = {"x": xa.DataArray(x, dims=["x"], attrs={"units": "mm", "long_name": "Lateral"}),
              "y": xa.DataArray(y, dims=["y"], attrs={"units": "mm", "long_name": "Elevation"}),
              "z": xa.DataArray(z, dims=["z"], attrs={"units": "mm", "long_name": "Axial"}),
              "focal_point_index": [0, 1]}
    dims = ["focal_point_index", "x", "y", "z"]
    ds = xa.Dataset({"p_min": xa.DataArray(p_min, dims=dims, attrs={"units": "Pa", "long_name": "PNP"}),
                     "p_max": xa.DataArray(p_max, dims=dims, attrs={"units": "Pa", "long_name": "PPP"}),
                     "intensity": xa.DataArray(inten, dims=dims, attrs={"units": "W/cm^2", "long_name": "Intensity"})},
                    coords=coords)
    arr = mod["Transducer"].gen_matrix_array(nx=4, ny=4, pitch=4, kerf=0.5, units="mm")
    foci = [mod["Point"](position=np.array(f), units="mm", id=f"f{i}") for i, f in enumerate(foci_mm)]
    delays = np.array([[np.linalg.norm(np.array(f) - el.get_position(units="mm")) * 1e-3 / 1500 for el in arr.elements]
                       for f in foci_mm])
    delays = delays.max(axis=1, keepdims=True) - delays
    apod = np.stack([np.ones(16), np.linspace(0.5, 1.0, 16)])
    sol = mod["Solution"](id="g", transducer=arr, delays=delays, apodizations=apod,
                          pulse=mod["Pulse"](frequency=400e3, amplitude=1.0, duration=50e-6),
                          sequence=mod["Sequence"](pulse_interval=0.01, pulse_count=4, pulse_train_interval=0.1, pulse_train_count=3),
                          voltage=12.0, foci=foci, target=foci[0], simulation_result=ds)
    opts = mod["SolutionAnalysisOptions"](mainlobe_radius=2.5, beamwidth_radius=5.0, sidelobe_radius=3.0, sidelobe_zmin=1.0,
                                          distance_units="mm")
    pattern = mod["Wheel"](center=True, num_spokes=1, spoke_radius=5.0, distance_units="mm", target_pressure=0.8, units="MPa")
    return sol, opts, pattern


def analysis_to_plain(a):
    out = {}
    for k, v in a.__dict__.items():
        if k == "param_constraints":
            continue
        out[k] = [float(x) for x in v] if isinstance(v, list) else (None if v is None else float(v))
    return out


def main():
    install_reference()
    from openlifu.bf import Pulse, Sequence
    from openlifu.bf.focal_patterns import Wheel
    from openlifu.geo import Point
    from openlifu.plan.solution import Solution
    from openlifu.plan.solution_analysis import SolutionAnalysisOptions
    from openlifu.xdc import Transducer
    from openlifu_b200 import xa
    mod = {"xa": xa, "Transducer": Transducer, "Point": Point, "Solution": Solution, "Pulse": Pulse, "Sequence": Sequence,
           "SolutionAnalysisOptions": SolutionAnalysisOptions, "Wheel": Wheel}
    sol, opts, pattern = synthetic_case(mod)
    # the on-disk JSON of a Solution as the reference writes it (plan/solution.py:406-437): ref_solution.json
    from datetime import datetime
    import importlib.util
    import types
    import openlifu.plan.solution as rsol
    stub = types.ModuleType("openlifu.db.subject")           # the encoder module imports the database layer for one name
    stub.Subject = type("Subject", (), {})
    sys.modules["openlifu.db.subject"] = stub
    spec = importlib.util.spec_from_file_location("openlifu_util_json_real", "/root/reference/src/openlifu/util/json.py")
    real_json = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(real_json)
    rsol.PYFUSEncoder = real_json.PYFUSEncoder               # the reference's own encoder (install_reference stubs it)
    sol.date_created = datetime(2024, 1, 2, 3, 4, 5)
    (HERE / "ref_solution.json").write_text(sol.to_json(include_simulation_data=False, compact=False))
    out = {"analysis": analysis_to_plain(sol.analyze(options=opts))}
    ita = sol.get_ita()
    out["ita_sum"] = float(np.asarray(ita.data).sum())
    out["ita_max_per_focus"] = [float(np.asarray(ita.data)[i].max()) for i in range(2)]
    apod_f, v0, v1 = sol.compute_scaling_factors(pattern, sol.analyze(options=opts))
    out["scaling"] = {"apod_factors": [float(a) for a in apod_f], "v0": float(v0), "v1": float(v1)}
    sol.scale(pattern, analysis_options=opts)
    out["scaled"] = {"voltage": float(sol.voltage), "apod_sum": [float(a.sum()) for a in sol.apodizations],
                     "p_min_max": [float(sol.simulation_result["p_min"][i].data.max()) for i in range(2)],
                     "p_max_max": [float(sol.simulation_result["p_max"][i].data.max()) for i in range(2)],
                     "intensity_max": [float(sol.simulation_result["intensity"][i].data.max()) for i in range(2)]}
    out["analysis_scaled"] = analysis_to_plain(sol.analyze(options=opts))
    (HERE / "ref_plan.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "ref_plan.json")


if __name__ == "__main__":
    main()
