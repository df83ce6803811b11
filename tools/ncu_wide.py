"""One time step (no source, then source active) of a C5-class phantom grid on ONE GPU inside a cudaProfilerStart/Stop range:

    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kw_ -o /tmp/wide python tools/ncu_wide.py 472
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import torch
import bench
from openlifu_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 472
kinds = [int(k) for k in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
arr, sp, x, y, z = bench.c5_scene(n)
d = [sp * 1e-3] * 3
nt_full, dt = _lib.make_time([n] * 3, d, 1500.0, 0.5)
sim = _lib.LifuSim([n] * 3, d, dt, 4, device=0)
bench.slab_setup(sim, arr, x, y, z, n, sp, dt, False)
_, _, st = sim.run()
print(st, file=sys.stderr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for k in kinds:
    sim.profile_stages(reps=1, with_source=k)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
sim.close()
