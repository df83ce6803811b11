#!/usr/bin/env python
"""Where does the end-to-end time of one run_simulation call go?  (C2 workload, one B200.)
Wraps the pieces of openlifu_b200.sim.kwave_if.run_simulation with wall-clock timers."""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402,F401


def main():
    import __graft_entry__ as ge
    ge.build()
    from openlifu_b200 import _lib, configs
    from openlifu_b200.sim import kwave_if
    cfg = configs.c3(216) if "C3" in sys.argv else configs.c2(216)
    params, foci, beams, cycles = configs.prepare(cfg)
    arr = cfg["arr"]
    T = {}

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            T[name] = T.get(name, 0.0) + time.perf_counter() - t0
            return r
        return w

    kwave_if.package_fields = timed("package_fields", kwave_if.package_fields)
    kwave_if.get_kgrid = timed("get_kgrid", kwave_if.get_kgrid)
    kwave_if.element_geometry = timed("element_geometry", kwave_if.element_geometry)
    _lib.LifuSim.run = timed("sim.run", _lib.LifuSim.run)
    _lib.LifuSim.run_packaged = timed("sim.run_packaged", _lib.LifuSim.run_packaged)
    _lib.LifuSim.set_two_z = timed("sim.set_two_z", _lib.LifuSim.set_two_z)
    kwave_if.package_arrays = timed("package_arrays", kwave_if.package_arrays)
    kwave_if._sample_checksum = timed("checksums", kwave_if._sample_checksum)
    _lib.LifuSim.set_medium = timed("sim.set_medium", _lib.LifuSim.set_medium)
    _lib.LifuSim.set_drive = timed("sim.set_drive", _lib.LifuSim.set_drive)
    _lib.LifuSim.set_elements = timed("sim.set_elements", _lib.LifuSim.set_elements)
    type(arr).drive_plan = timed("drive_plan", type(arr).drive_plan)
    if "nogc" in sys.argv:
        import gc
        gc.disable()
    for it in range(7):
        T.clear()
        delays, apod = beams[it % len(beams)]
        ses = next(iter(kwave_if._SESSIONS.values()), None)
        if ses is not None:
            ses.medium_key = None
        t0 = time.perf_counter()
        ds, out = kwave_if.run_simulation(arr=arr, params=params, delays=delays, apod=apod, freq=cfg["pulse"].frequency,
                                          cycles=cycles, amplitude=cfg["pulse"].amplitude, gpu=True)
        tot = time.perf_counter() - t0
        rest = tot - sum(T.values())
        print(f"call {it}: total {tot*1e3:.1f} ms  loop {out['stats']['loop_ms']:.1f}  setup {out['stats']['setup_ms']:.1f} | "
              + "  ".join(f"{k} {v*1e3:.1f}" for k, v in T.items()) + f"  other {rest*1e3:.1f}", flush=True)


if __name__ == "__main__":
    main()
