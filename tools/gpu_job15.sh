#!/bin/bash
mkdir -p gpurun_out
for mode in host device host device; do
  LIFU_PACKAGING=$mode python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_pkg_$mode.json 2> gpurun_out/r2_bench_pkg_$mode.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_pkg_$mode.json"))
print("$mode", "e2e", round(d["e2e"]["value"]), "wall", round(d["e2e"]["wall_ms_per_step"],1), "loop", round(d["e2e"]["solver_loop_ms_per_step"],1), "d2h", d["e2e"]["d2h_bytes_per_step"])
PY
done
LIFU_PACKAGING=device python -m pytest tests/test_gpu_api.py -q -k "run_simulation or packag or dataset" 2>&1 | tail -3
