#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider -x > gpurun_out/r2_slabwide_tests.log 2>&1
tail -30 gpurun_out/r2_slabwide_tests.log
