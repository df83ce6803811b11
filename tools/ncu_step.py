"""One time step of every kind (generic source, steady source window, no source) inside a cudaProfilerStart/Stop range,
for `ncu --profile-from-start off`:

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_c2_step \
        python tools/ncu_step.py C2
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import numpy as np
import torch
from openlifu_b200 import _lib, configs
from openlifu_b200.sim.kwave_if import element_geometry, get_kgrid

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 70
cfg = {"C1": configs.c1, "C2": configs.c2, "C3": configs.c3}[wl]()
params, foci, beams, cycles = configs.prepare(cfg)
kg = get_kgrid(params.coords)
arr = cfg["arr"]
base = np.sin(2 * np.pi * cfg["pulse"].frequency * np.arange(0, cycles / cfg["pulse"].frequency, kg["dt"]))
names = ("sound_speed", "density", "attenuation")
homog = all(float(params[k].data.min()) == float(params[k].data.max()) for k in names)
sim = _lib.LifuSim(kg["N"], kg["d"], kg["dt"], min(nt, kg["Nt"]))
if homog:
    sim.set_medium(*[float(params[k].data.flat[0]) for k in names])
else:
    sim.set_medium(*[params[k].data for k in names])
sim.set_elements(*element_geometry(arr, [-float(c.mean()) * 1e-3 for c in params.coords.values()]), 0.05, 5)
n_delay, gains, bg = arr.drive_plan(kg["dt"], *beams[0])
sim.set_drive(base * bg, n_delay, gains)
_, _, st = sim.run()
print(st, file=sys.stderr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
kinds = [1, 0] + ([2] if st.get("steady_source_steps", 0) else [])
for k in kinds:
    sim.profile_stages(reps=1, with_source=k)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
sim.close()
