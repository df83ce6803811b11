#!/bin/bash
# 2-GPU job: multi-rank slab parity (fused passes with routed stores), the bench line with its slab leg
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_slab.py -q -k "two_rank" --tb=short -p no:cacheprovider > gpurun_out/r2_slabwide_tests_n2.log 2>&1
tail -25 gpurun_out/r2_slabwide_tests_n2.log | cut -c 1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2_wide.json 2> gpurun_out/r2_bench_n2_wide.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n2_wide.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("slab"))[:3000])
PY
tail -5 gpurun_out/r2_bench_n2_wide.err
