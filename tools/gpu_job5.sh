#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_analysis.py -q --maxfail=10 > gpurun_out/r2_api_tests.log 2>&1
tail -30 gpurun_out/r2_api_tests.log
python -m pytest tests/test_gpu_parity.py -q -k "rank_two" 2>&1 | tail -3
python -m pytest tests/test_gpu_v3.py -q --maxfail=20 > gpurun_out/r2_v3tests2.log 2>&1
tail -15 gpurun_out/r2_v3tests2.log
python tools/perf_variants.py C1 229 LIFU_PIPELINE=v3 LIFU_PIPELINE=v1 > gpurun_out/r2_c1_variants2.jsonl 2> gpurun_out/r2_c1_variants2.err
cut -c 1-900 gpurun_out/r2_c1_variants2.jsonl; tail -3 gpurun_out/r2_c1_variants2.err
timeout 600 python tools/single_grid.py 728 6 v3 > gpurun_out/r2_768_v3_stages2.jsonl 2> gpurun_out/r2_768_v3_stages2.err
cat gpurun_out/r2_768_v3_stages2.jsonl; tail -3 gpurun_out/r2_768_v3_stages2.err
python tools/plan_profile.py 7 216 > gpurun_out/r2_plan_profile.json 2> gpurun_out/r2_plan_profile.err
cat gpurun_out/r2_plan_profile.json; tail -5 gpurun_out/r2_plan_profile.err
