#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q --maxfail=10 -k "steady or v2_small_water or v2_matches_v1 or c2_grid_short or v2_heterogeneous" > gpurun_out/r2_steady_tests.log 2>&1
tail -30 gpurun_out/r2_steady_tests.log
python tools/perf_variants.py C2 749 "" LIFU_SOURCE_STEADY=0 > gpurun_out/r2_variants3.jsonl 2> gpurun_out/r2_variants3.err
cut -c 1-2500 gpurun_out/r2_variants3.jsonl; tail -3 gpurun_out/r2_variants3.err
python tools/perf_variants.py C3 749 "" > gpurun_out/r2_variants3_c3.jsonl 2> gpurun_out/r2_variants3_c3.err
cut -c 1-400 gpurun_out/r2_variants3_c3.jsonl; tail -3 gpurun_out/r2_variants3_c3.err
timeout 600 python tools/single_grid.py 728 6 v3 > gpurun_out/r2_768_v3_stages.jsonl 2> gpurun_out/r2_768_v3_stages.err
cat gpurun_out/r2_768_v3_stages.jsonl; tail -3 gpurun_out/r2_768_v3_stages.err
python -m pytest tests/test_gpu_v3.py -q -k "c1_full or radix" 2>&1 | tail -3
