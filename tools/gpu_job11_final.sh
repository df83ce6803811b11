#!/bin/bash
# final 1-GPU pass: variants of the gradient y-inverse, full GPU test suite, smoke, bench lines
mkdir -p gpurun_out
python tools/perf_variants.py C2 749 "" LIFU_V2_YGRAD=split > gpurun_out/r2_variants_ygrad.jsonl 2> gpurun_out/r2_variants_ygrad.err
cut -c 1-420 gpurun_out/r2_variants_ygrad.jsonl; tail -2 gpurun_out/r2_variants_ygrad.err
LIFU_V2_YGRAD=split python -m pytest tests/test_gpu_parity.py -q -k "v2_small_water or v2_matches_v1 or v2_mixed_axes or v2_heterogeneous_lossless" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -3 gpurun_out/r2_smoke_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; cut -c 1-1500 gpurun_out/r2_bench_c2.json; tail -3 gpurun_out/r2_bench_c2.err
python bench.py --workload C3 --steps 3 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; cut -c 1-700 gpurun_out/r2_bench_c3.json; tail -3 gpurun_out/r2_bench_c3.err
python bench.py --workload C1 --steps 5 --warmup 3 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; cut -c 1-700 gpurun_out/r2_bench_c1.json; tail -3 gpurun_out/r2_bench_c1.err
python bench.py --poses 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_poses.json 2> gpurun_out/r2_bench_poses.err; cut -c 1-900 gpurun_out/r2_bench_poses.json; tail -3 gpurun_out/r2_bench_poses.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c 1-900 gpurun_out/r2_bench_reference.json
rm -f gpurun_out/r2_parity_final.jsonl
LIFU_PARITY_LOG=gpurun_out/r2_parity_final.jsonl python -m pytest tests -m gpu -q > gpurun_out/r2_gputest_final.log 2>&1
tail -8 gpurun_out/r2_gputest_final.log
