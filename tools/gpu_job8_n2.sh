#!/bin/bash
# 2-GPU job: multi-rank slab parity, the bench line with its slab leg, racecheck on the peer-store exchange
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m pytest tests/test_gpu_slab.py -q -k "two_rank" > gpurun_out/r2_slab_tests_n2.log 2>&1
tail -5 gpurun_out/r2_slab_tests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
cut -c 1-3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --target-processes all --error-exitcode 9 --print-limit 5 \
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/slab_rank.py --case water --exchange peer --out gpurun_out/r2_slab_race.json > gpurun_out/r2_san_slab_peer.log 2>&1
echo "slab peer racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/r2_san_slab_peer.log | tr '\n' ' ')"
