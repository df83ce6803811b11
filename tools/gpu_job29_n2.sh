#!/bin/bash
# 2-GPU: NCCL-transport slab test after the block-size fix, all two-rank slab tests, final bench line with its slab leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py -q -k "two_rank" --tb=short -p no:cacheprovider > gpurun_out/r2_slab_tests_n2_final.log 2>&1
tail -4 gpurun_out/r2_slab_tests_n2_final.log | cut -c 1-600
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n2_final.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps({k: v for k, v in d["slab"].items() if k != "stages"}))
PY
tail -2 gpurun_out/r2_bench_n2_final.err
