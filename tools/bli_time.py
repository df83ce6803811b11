"""How long does the GPU band-limited-interpolant source build (lifu_set_elements) take?  C2 array / grid."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import numpy as np
from openlifu_b200 import _lib, configs
from openlifu_b200.sim.kwave_if import element_geometry, get_kgrid
cfg = configs.c2(216)
params, foci, beams, cycles = configs.prepare(cfg)
kg = get_kgrid(params.coords)
off = [-float(c.mean()) * 1e-3 for c in params.coords.values()]
geo = element_geometry(cfg["arr"], off)
for i in range(3):
    t0 = time.perf_counter()
    sim = _lib.LifuSim(kg["N"], kg["d"], kg["dt"], 4)
    t1 = time.perf_counter()
    sim.set_medium(1500.0, 1000.0, 0.0)
    t2 = time.perf_counter()
    n = sim.set_elements(*geo, 0.05, 5)
    t3 = time.perf_counter()
    n = sim.set_elements(*geo, 0.05, 5)
    t4 = time.perf_counter()
    print(f"handle {i}: create {1e3*(t1-t0):.1f} ms, set_medium {1e3*(t2-t1):.1f}, set_elements {1e3*(t3-t2):.1f}, again {1e3*(t4-t3):.1f} (n_src {n})")
    sim.close()
