#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_slab.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_wide_tests4.log 2>&1
tail -5 gpurun_out/r2_wide_tests4.log
timeout 300 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_final.jsonl 2> gpurun_out/r2_wide_512_final.err; cut -c 1-900 gpurun_out/r2_wide_512_final.jsonl; tail -3 gpurun_out/r2_wide_512_final.err
timeout 400 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768_final.jsonl 2> gpurun_out/r2_wide_768_final.err; cut -c 1-900 gpurun_out/r2_wide_768_final.jsonl; tail -3 gpurun_out/r2_wide_768_final.err
