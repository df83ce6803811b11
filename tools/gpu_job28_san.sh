#!/bin/bash
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in racecheck memcheck synccheck; do
  timeout 600 $S --tool $tool --error-exitcode 9 --print-limit 5 python tools/race_wide.py > gpurun_out/r2_san_wide_$tool.log 2>&1
  echo "wide ($tool): rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|launches' gpurun_out/r2_san_wide_$tool.log | tr '\n' ' ')"
done
