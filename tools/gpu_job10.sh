#!/bin/bash
mkdir -p gpurun_out
timeout 2400 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python tools/race_v2.py > gpurun_out/r2_san_v2_race.log 2>&1
echo "v2 racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|launches' gpurun_out/r2_san_v2_race.log | tr '\n' ' ')"
