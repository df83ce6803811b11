"""Short C2-grid run for profiling under ncu: python tools/prof_run.py [n_time_steps] [n_inner]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import numpy as np
from openlifu_b200 import _lib, configs
from openlifu_b200.sim.kwave_if import element_geometry, get_kgrid

nt = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n_inner = int(sys.argv[2]) if len(sys.argv) > 2 else 216
cfg = configs.c2(n_inner)
params, foci, beams, cycles = configs.prepare(cfg)
kg = get_kgrid(params.coords)
arr = cfg["arr"]
t = np.arange(0, cycles / 400e3, kg["dt"])
base = np.sin(2 * np.pi * 400e3 * t)
sim = _lib.LifuSim(kg["N"], kg["d"], kg["dt"], nt)
sim.set_medium(1500.0, 1000.0, 0.0)
off = [-float(c.mean()) * 1e-3 for c in params.coords.values()]
sim.set_elements(*element_geometry(arr, off), 0.05, 5)
n_delay, gains, bg = arr.drive_plan(kg["dt"], *beams[0])
sim.set_drive(base * bg, n_delay, gains)
pm, pn, st = sim.run()
print(st)
for name, ms, b in sim.profile_stages(reps=3):
    print(f"{name:20s} {ms:8.4f} ms  {b * st['voxels'] / ms / 1e6 if ms else 0:8.1f} GB/s")
