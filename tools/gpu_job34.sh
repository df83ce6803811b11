#!/bin/bash
mkdir -p gpurun_out
timeout 70 python tools/single_grid.py 728 3 v2 > gpurun_out/r2_wide_768_final2.jsonl 2> gpurun_out/r2_wide_768_final2.err; cut -c 1-900 gpurun_out/r2_wide_768_final2.jsonl; tail -2 gpurun_out/r2_wide_768_final2.err
