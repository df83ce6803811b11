#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_v3.py -q --maxfail=20 > gpurun_out/r2_v3tests3.log 2>&1
tail -15 gpurun_out/r2_v3tests3.log
python -m pytest tests/test_gpu_api.py -q -k "field_stack" 2>&1 | tail -15
python tools/perf_variants.py C1 229 LIFU_PIPELINE=v3 LIFU_PIPELINE=v3,LIFU_V3_LS=8 LIFU_PIPELINE=v3,LIFU_V3_LS=8,LIFU_V3_LX=8,LIFU_V3_TS=128,LIFU_V3_TX=128 > gpurun_out/r2_c1_variants3.jsonl 2> gpurun_out/r2_c1_variants3.err
cut -c 1-700 gpurun_out/r2_c1_variants3.jsonl; tail -3 gpurun_out/r2_c1_variants3.err
rm -f gpurun_out/r2_768_v3_sweep.jsonl
for cfg in "" "LIFU_V3_TS=128" "LIFU_V3_TX=512" "LIFU_V3_LS=4 LIFU_V3_TS=128" "LIFU_V3_LX=4 LIFU_V3_TX=128" "LIFU_V3_LS=16 LIFU_V3_TS=256" ; do
  echo "## $cfg" >> gpurun_out/r2_768_v3_sweep.jsonl
  env $cfg timeout 300 python tools/single_grid.py 728 4 v3 >> gpurun_out/r2_768_v3_sweep.jsonl 2>> gpurun_out/r2_768_v3_sweep.err
done
cat gpurun_out/r2_768_v3_sweep.jsonl; tail -3 gpurun_out/r2_768_v3_sweep.err
python tools/plan_profile.py 7 216 > gpurun_out/r2_plan_profile.json 2> gpurun_out/r2_plan_profile.err
cat gpurun_out/r2_plan_profile.json; tail -5 gpurun_out/r2_plan_profile.err
