#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -k "small_water or tilted or overhang or edge_cases or candidates or oracle_geometry or wheel" 2>&1 | tail -4
python bench.py --poses 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_poses.json 2> gpurun_out/r2_bench_poses.err; cut -c 1-300 gpurun_out/r2_bench_poses.json; tail -3 gpurun_out/r2_bench_poses.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_poses.json"))
print("poses: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "wall", d["e2e"]["wall_ms_per_step"], "loop", d["e2e"]["solver_loop_ms_per_step"])
PY
python tools/bli_time.py 2>&1 | tail -5
