#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_analysis.py -q --maxfail=10 > gpurun_out/r2_api_tests.log 2>&1
tail -30 gpurun_out/r2_api_tests.log
python -m pytest tests/test_gpu_parity.py -q -k "rank_two" 2>&1 | tail -3
python tools/plan_profile.py 7 216 > gpurun_out/r2_plan_profile.json 2> gpurun_out/r2_plan_profile.err
cat gpurun_out/r2_plan_profile.json; tail -5 gpurun_out/r2_plan_profile.err
