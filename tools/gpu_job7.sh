#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_v3.py -q --maxfail=20 > gpurun_out/r2_v3tests4.log 2>&1
tail -5 gpurun_out/r2_v3tests4.log
python -m pytest tests/test_gpu_api.py -q -k "field_stack or plan_on_device" 2>&1 | tail -5
EXP=$PWD/openlifu-python_b200/lib/liblifusim_exp.so
rm -f gpurun_out/r2_768_v3_sweep2.jsonl
for cfg in "" "LIFUSIM_LIB=$EXP LIFU_V3_RMAX=8 LIFU_V3_TS=512 LIFU_V3_TX=512" "LIFUSIM_LIB=$EXP LIFU_V3_RMAX=8 LIFU_V3_TS=256 LIFU_V3_TX=256" "LIFU_V3_RMAX=8" ; do
  echo "## $cfg" >> gpurun_out/r2_768_v3_sweep2.jsonl
  env $cfg timeout 300 python tools/single_grid.py 728 4 v3 >> gpurun_out/r2_768_v3_sweep2.jsonl 2>> gpurun_out/r2_768_v3_sweep2.err
done
cat gpurun_out/r2_768_v3_sweep2.jsonl; tail -3 gpurun_out/r2_768_v3_sweep2.err
python tools/perf_variants.py C1 229 LIFU_PIPELINE=v3 LIFU_PIPELINE=v3,LIFUSIM_LIB=$EXP,LIFU_V3_RMAX=9,LIFU_V3_TS=256,LIFU_V3_TX=256 LIFU_PIPELINE=v3,LIFUSIM_LIB=$EXP,LIFU_V3_RMAX=9,LIFU_V3_TS=128,LIFU_V3_TX=128,LIFU_V3_LS=8,LIFU_V3_LX=8 > gpurun_out/r2_c1_variants4.jsonl 2> gpurun_out/r2_c1_variants4.err
cut -c 1-260 gpurun_out/r2_c1_variants4.jsonl; tail -3 gpurun_out/r2_c1_variants4.err
python tools/plan_profile.py 7 216 > gpurun_out/r2_plan_profile.json 2> gpurun_out/r2_plan_profile.err
cat gpurun_out/r2_plan_profile.json; tail -5 gpurun_out/r2_plan_profile.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2_c2_step python tools/ncu_step.py C2 > gpurun_out/r2_ncu_c2.log 2>&1
tail -3 gpurun_out/r2_ncu_c2.log
ncu -i gpurun_out/r2_c2_step.ncu-rep --page raw --csv > gpurun_out/r2_c2_step_raw.csv 2>/dev/null
wc -l gpurun_out/r2_c2_step_raw.csv
