#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:kw_ -f -o /tmp/r2_wide512 python tools/ncu_wide.py 472 0 > gpurun_out/r2_ncu_wide512.log 2>&1
tail -2 gpurun_out/r2_ncu_wide512.log
ncu -i /tmp/r2_wide512.ncu-rep --page raw --csv > gpurun_out/r2_wide512_step_raw.csv 2>/dev/null
wc -l gpurun_out/r2_wide512_step_raw.csv
ncu -i /tmp/r2_wide512.ncu-rep --page source --csv -k regex:kw_z --launch-skip 1 --launch-count 1 > gpurun_out/r2_wide512_zdiv_source.csv 2>/dev/null
wc -l gpurun_out/r2_wide512_zdiv_source.csv
ls -la /tmp/r2_wide512.ncu-rep
