#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider -k "phantom or absorbing or long" > gpurun_out/r2_wide_tests3.log 2>&1
tail -5 gpurun_out/r2_wide_tests3.log
timeout 300 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_sm.jsonl 2> gpurun_out/r2_wide_512_sm.err; cut -c 1-900 gpurun_out/r2_wide_512_sm.jsonl; tail -3 gpurun_out/r2_wide_512_sm.err
timeout 400 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768_sm.jsonl 2> gpurun_out/r2_wide_768_sm.err; cut -c 1-900 gpurun_out/r2_wide_768_sm.jsonl; tail -3 gpurun_out/r2_wide_768_sm.err
