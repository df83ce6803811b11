#!/bin/bash
# compute-sanitizer over the fused pipelines (v2, v3), the steady-source path and the plan stack; logs -> gpurun_out/r2_san_*.log
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool pytest-args...
  name=$1; tool=$2; shift 2
  timeout 900 $S --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest "$@" -q -x -p no:cacheprovider > gpurun_out/r2_san_${name}.log 2>&1
  echo "$name ($tool): rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r2_san_${name}.log | tr '\n' ' ')"
}
run v2_race racecheck tests/test_gpu_parity.py -k "v2_small_water or v2_homogeneous_absorbing or v2_steady_source_matches"
run v2_sync synccheck tests/test_gpu_parity.py -k "v2_small_water or v2_heterogeneous_lossless or v2_steady_source_other"
run v2_mem memcheck tests/test_gpu_parity.py -k "v2_small_water or v2_mixed_axes or v2_steady_source_matches"
run v3_race racecheck tests/test_gpu_v3.py -k "odd_grid or radix_5"
run v3_mem memcheck tests/test_gpu_v3.py -k "odd_grid or radix_5 or heterogeneous_absorbing or edge"
run stack_mem memcheck tests/test_gpu_api.py -k "field_stack or plan_on_device or label_medium"
