"""Short v2 runs for `compute-sanitizer --tool racecheck` (every kernel of the fused pipeline, few time steps):
lossless water with the generic source path, the steady-source window, a heterogeneous absorbing medium."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import numpy as np
from tests import cases

c = cases.v2_small_case(steps=6)
g = cases.run_cuda_case(c)
assert g["stats"]["fft_launches"] == 0
print("water:", g["stats"]["kernel_launches"], "launches")
c = cases.make_case([(-20, 19), (-22, 21), (-3, 32)], 1.0, 3, 3, 4.0, 0.5, (3, -2, 18), 400e3, 10, dt=3e-7, t_end=34 * 3e-7)
g = cases.run_cuda_case(c)
assert g["stats"]["fft_launches"] == 0 and g["stats"]["steady_source_steps"] > 0
print("steady:", g["stats"]["steady_source_steps"], "steady steps,", g["stats"]["kernel_launches"], "launches")
c = cases.v2_small_case(steps=5)
c["c0"], c["rho0"], c["alpha"] = cases.layered_phantom(tuple(c["N"]))
c["dt"], c["t_end"] = 1.5e-7, 5 * 1.5e-7
g = cases.run_cuda_case(c)
assert g["stats"]["fft_launches"] == 0 and g["stats"]["absorbing"] == 1
print("phantom:", g["stats"]["kernel_launches"], "launches")
