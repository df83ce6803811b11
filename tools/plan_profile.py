#!/usr/bin/env python
"""Where does Protocol.calc_solution spend its time on a multi-focus pattern?  (C4-class workload: C2 grid,
Wheel pattern, one B200.)  Runs the whole planning call twice -- beam analysis on the device
(`lifu_analysis_*`) and on the host (numpy) -- and prints one JSON line with the wall time of
every phase and foci/s of the complete call.

    python tools/plan_profile.py [num_spokes=7] [n_inner=216]
"""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    spokes = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    n_inner = int(sys.argv[2]) if len(sys.argv) > 2 else 216
    import __graft_entry__ as ge
    ge.build()
    from openlifu_b200 import configs, xa
    from openlifu_b200.plan import Protocol, Solution
    from openlifu_b200.plan import protocol as protocol_mod
    cfg = configs.c4(n_inner, num_spokes=spokes)
    cfg["focal_pattern"].target_pressure = 1.0e6
    cfg["focal_pattern"].units = "Pa"
    prot = Protocol(pulse=cfg["pulse"], sequence=cfg["sequence"], focal_pattern=cfg["focal_pattern"], sim_setup=cfg["setup"],
                    seg_method=cfg["seg"])
    T = {}

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            T[name] = T.get(name, 0.0) + time.perf_counter() - t0
            return r
        return w

    Protocol._simulate_foci = timed("simulate_foci", Protocol._simulate_foci)
    Solution.analyze = timed("analyze(x2)", Solution.analyze)
    Solution.scale = timed("scale(incl. 1 analyze)", Solution.scale)
    protocol_mod.xa.concat = timed("concat", xa.concat)
    lines = {}
    results = {}
    Protocol._simulate_foci_on_device = timed("simulate_foci", Protocol._simulate_foci_on_device)
    fields = {}
    for engine, on_device in (("cuda", False), ("host", False), ("cuda", False), ("plan_on_device", True), ("plan_on_device", True)):
        os.environ["LIFU_ANALYZE"] = "cuda" if on_device else engine
        T.clear()
        t0 = time.perf_counter()
        sol, agg, ana = prot.calc_solution(cfg["target"], cfg["arr"], simulate=True, scale=True, use_gpu=True,
                                           on_device=on_device)
        tot = time.perf_counter() - t0
        n = len(sol.foci)
        lines[engine] = {"foci": n, "total_s": round(tot, 3), "foci_per_s": round(n / tot, 3),
                         **{k: round(v, 3) for k, v in T.items()}}
        results[engine] = ana
        if engine != "host":
            fields.setdefault(engine, []).append((sol.simulation_result, agg, sol.voltage, sol.apodizations.copy()))
    def diff(a, b):
        out = {}
        for j, tag in ((0, "stack"), (1, "aggregated")):
            for k in ("p_max", "p_min", "intensity"):
                x, y = np.asarray(a[j][k].data), np.asarray(b[j][k].data)
                ne = x != y
                if ne.any():
                    out[f"{tag}.{k}"] = {"differing": int(ne.sum()), "max_rel": float(np.max(np.abs(x[ne] - y[ne]) / np.abs(y[ne]).clip(1e-300)))}
        if a[2] != b[2] or not np.array_equal(a[3], b[3]):
            out["voltage/apod"] = [float(a[2]), float(b[2])]
        return out
    d_route = diff(fields["plan_on_device"][-1], fields["cuda"][-1])
    d_repeat_host = diff(fields["cuda"][0], fields["cuda"][-1])
    d_repeat_dev = diff(fields["plan_on_device"][0], fields["plan_on_device"][-1])
    same = not d_route
    a, b = results["cuda"], results["host"]
    worst = 0.0
    for k, v in b.__dict__.items():
        if k == "param_constraints" or v is None:
            continue
        g = np.asarray(getattr(a, k), dtype=float)
        w = np.asarray(v, dtype=float)
        with np.errstate(invalid="ignore", divide="ignore"):
            d = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
        d = d[~(np.isnan(g) & np.isnan(w))]
        worst = max(worst, float(d.max()) if d.size else 0.0)
    print(json.dumps({"workload": f"C4-class: C2 grid ({n_inner}^3 inner), Wheel with {spokes} spokes + centre, calc_solution(scale=True)",
                      "plan_on_device": lines["plan_on_device"], "plan_on_device_bit_identical_to_host_route": bool(same),
                      "differences": {"device_vs_host_route": d_route, "host_route_run_to_run": d_repeat_host,
                                      "device_route_run_to_run": d_repeat_dev},
                      "analysis_on_device": lines["cuda"], "analysis_on_host": lines["host"],
                      "max_rel_diff_between_engines": worst, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
