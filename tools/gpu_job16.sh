#!/bin/bash
# wide pipeline: parity tests, then 512^3 / 768^3 phantom grids and the 256^3 C2 grid against the other pipelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_wide_tests.log 2>&1
tail -40 gpurun_out/r2_wide_tests.log
timeout 300 python tools/single_grid.py 472 6 v2 v1 > gpurun_out/r2_wide_512.jsonl 2> gpurun_out/r2_wide_512.err; cut -c 1-900 gpurun_out/r2_wide_512.jsonl; tail -3 gpurun_out/r2_wide_512.err
timeout 400 python tools/single_grid.py 728 4 v2 v1 > gpurun_out/r2_wide_768.jsonl 2> gpurun_out/r2_wide_768.err; cut -c 1-900 gpurun_out/r2_wide_768.jsonl; tail -3 gpurun_out/r2_wide_768.err
timeout 300 python tools/perf_variants.py C2 100 "" LIFU_WIDE_SQUARE=1 > gpurun_out/r2_wide_c2.jsonl 2> gpurun_out/r2_wide_c2.err; cut -c 1-1200 gpurun_out/r2_wide_c2.jsonl; tail -3 gpurun_out/r2_wide_c2.err
