#!/bin/bash
mkdir -p gpurun_out
LIFU_WIDE_ZPERSIST=1 timeout 600 python -m pytest tests/test_gpu_wide.py tests/test_gpu_slab.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_wide_tests5.log 2>&1
tail -4 gpurun_out/r2_wide_tests5.log
for zp in 1 0; do
LIFU_WIDE_ZPERSIST=$zp timeout 300 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_zp$zp.jsonl 2> gpurun_out/r2_wide_512_zp.err; cut -c 1-800 gpurun_out/r2_wide_512_zp$zp.jsonl; tail -2 gpurun_out/r2_wide_512_zp.err
done
LIFU_WIDE_ZPERSIST=1 timeout 300 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768_zp1.jsonl 2> gpurun_out/r2_wide_768_zp.err; cut -c 1-800 gpurun_out/r2_wide_768_zp1.jsonl; tail -2 gpurun_out/r2_wide_768_zp.err
