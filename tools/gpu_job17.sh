#!/bin/bash
# wide pipeline: H-plane padding experiment, parity, 512^3 / 768^3 timings
mkdir -p gpurun_out
timeout 300 python tools/perf_variants.py C2 60 LIFU_WIDE_SQUARE=1,LIFU_WIDE_HPAD=0 LIFU_WIDE_SQUARE=1,LIFU_WIDE_HPAD=16 LIFU_WIDE_SQUARE=1,LIFU_WIDE_HPAD=80 > gpurun_out/r2_wide_hpad_c2.jsonl 2> gpurun_out/r2_wide_hpad_c2.err
python - <<PY
import json
for l in open("gpurun_out/r2_wide_hpad_c2.jsonl"):
    d=json.loads(l); print(d["variant"], d["ms_per_step"], d["stages_nosrc"])
PY
tail -3 gpurun_out/r2_wide_hpad_c2.err
timeout 900 python -m pytest tests/test_gpu_wide.py -q --tb=short -p no:cacheprovider > gpurun_out/r2_wide_tests.log 2>&1
tail -5 gpurun_out/r2_wide_tests.log
for pad in 16 0; do
LIFU_WIDE_HPAD=$pad timeout 300 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_pad$pad.jsonl 2> gpurun_out/r2_wide_512.err; cut -c 1-900 gpurun_out/r2_wide_512_pad$pad.jsonl; tail -3 gpurun_out/r2_wide_512.err
done
timeout 400 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768.jsonl 2> gpurun_out/r2_wide_768.err; cut -c 1-900 gpurun_out/r2_wide_768.jsonl; tail -3 gpurun_out/r2_wide_768.err
