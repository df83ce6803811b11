#!/bin/bash
# round-2 GPU job 1: traversal-order variants, bench line, smoke, full GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_job1_gpu.txt
python tools/perf_variants.py C2 200 LIFU_V2_ORDER=0,LIFU_PM_ALWAYS=1 LIFU_V2_ORDER=0 LIFU_V2_ORDER=1 LIFU_V2_ORDER=2 LIFU_V2_ORDER=3 > gpurun_out/r2_variants1.jsonl 2> gpurun_out/r2_variants1.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1
rm -f gpurun_out/r2_parity1.jsonl
LIFU_PARITY_LOG=gpurun_out/r2_parity1.jsonl python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest1.log 2>&1
tail -5 gpurun_out/r2_gputest1.log
cat gpurun_out/r2_variants1.jsonl | cut -c 1-400
