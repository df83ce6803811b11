#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_v3.py -q --maxfail=30 > gpurun_out/r2_v3tests.log 2>&1
tail -40 gpurun_out/r2_v3tests.log
python tools/perf_variants.py C1 229 "" LIFU_PIPELINE=v1 > gpurun_out/r2_c1_variants.jsonl 2> gpurun_out/r2_c1_variants.err
cut -c 1-1200 gpurun_out/r2_c1_variants.jsonl
tail -3 gpurun_out/r2_c1_variants.err
timeout 600 python tools/single_grid.py 728 10 auto v1 > gpurun_out/r2_768_single.jsonl 2> gpurun_out/r2_768_single.err
cat gpurun_out/r2_768_single.jsonl; tail -3 gpurun_out/r2_768_single.err
python tools/perf_variants.py C2 200 "" > gpurun_out/r2_variants2.jsonl 2> gpurun_out/r2_variants2.err
cut -c 1-300 gpurun_out/r2_variants2.jsonl
