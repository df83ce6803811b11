#!/bin/bash
# ncu --set full captures of one step of every kind (C2, C3, C1); reports stay in /tmp, CSV tables come back
mkdir -p gpurun_out
for wl in C2 C3 C1; do
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o /tmp/r2_${wl}_step python tools/ncu_step.py $wl > gpurun_out/r2_ncu_${wl}.log 2>&1
  tail -2 gpurun_out/r2_ncu_${wl}.log
  ncu -i /tmp/r2_${wl}_step.ncu-rep --page raw --csv > gpurun_out/r2_${wl}_step_raw.csv 2>/dev/null
  wc -l gpurun_out/r2_${wl}_step_raw.csv
done
# launch list of the bench command (cold-cache, serialised: the kernels' SHARES of the step are what must agree with bench.py)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_run.log 2>&1
wc -l gpurun_out/r2_launches.csv
bash tools/gpu_job_sanitizer.sh
du -sh gpurun_out
