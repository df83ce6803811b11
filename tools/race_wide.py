"""Short runs of pipeline wide for compute-sanitizer (racecheck / memcheck / synccheck): every kernel family on non-square
factorisations, a few time steps each -- water with the source path, a heterogeneous absorbing phantom, a long (768-point) x
axis, and the slab code path (routed stores into the own exchange buffer) on one rank."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
from tests import cases

c = cases.wide_case((128, 64, 128), steps=5)
g = cases.run_cuda_case(c, pml=cases.WIDE_PML)
assert g["stats"]["fft_launches"] == 0
print("water 128 x 64 x 128:", g["stats"]["kernel_launches"], "launches")
c = cases.wide_case((64, 128, 64), steps=4)
c["c0"], c["rho0"], c["alpha"] = cases.layered_phantom(tuple(c["N"]))
c["dt"], c["t_end"] = 1.5e-7, 4 * 1.5e-7
g = cases.run_cuda_case(c, pml=cases.WIDE_PML)
assert g["stats"]["fft_launches"] == 0 and g["stats"]["absorbing"] == 1
print("phantom 64 x 128 x 64:", g["stats"]["kernel_launches"], "launches")
c = cases.wide_case((768, 64, 64), steps=3)
c["c0"], c["rho0"], c["alpha"] = cases.layered_phantom(tuple(c["N"]))
c["dt"], c["t_end"] = 1.5e-7, 3 * 1.5e-7
g = cases.run_cuda_case(c, pml=cases.WIDE_PML)
assert g["stats"]["fft_launches"] == 0
print("phantom 768 x 64 x 64:", g["stats"]["kernel_launches"], "launches")
from openlifu_b200 import _lib
c = cases.wide_case((128, 64, 128), steps=4)
g = cases.run_cuda_case_slab(c, 0, 1, _lib.slab_unique_id(), exchange="peer")
assert g["stats"]["fft_launches"] == 0
print("slab path, one rank:", g["stats"]["kernel_launches"], "launches")
