#!/bin/bash
mkdir -p gpurun_out
LIFU_WIDE_LANES=16 timeout 600 python -m pytest tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider -k "long or phantom or 128-128" > gpurun_out/r2_wide_tests_l16.log 2>&1
tail -5 gpurun_out/r2_wide_tests_l16.log
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_slabwide_tests2.log 2>&1
tail -5 gpurun_out/r2_slabwide_tests2.log
