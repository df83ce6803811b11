#!/bin/bash
# experiment build (16-lane tiles on 32-thread lines): single-GPU 512^3 / 768^3, then the 2-GPU slab leg if two GPUs are visible
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
TAG=${1:-lanes16}
timeout 300 python tools/single_grid.py 472 6 v2 > gpurun_out/r2_wide_512_$TAG.jsonl 2> gpurun_out/r2_wide_512_$TAG.err; cut -c 1-900 gpurun_out/r2_wide_512_$TAG.jsonl; tail -3 gpurun_out/r2_wide_512_$TAG.err
timeout 400 python tools/single_grid.py 728 4 v2 > gpurun_out/r2_wide_768_$TAG.jsonl 2> gpurun_out/r2_wide_768_$TAG.err; cut -c 1-900 gpurun_out/r2_wide_768_$TAG.jsonl; tail -3 gpurun_out/r2_wide_768_$TAG.err
if [ "$N" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload C5 --n-inner 472 --time-steps 10 --steps 2 --warmup 1 > gpurun_out/r2_c5_512_n${N}_$TAG.json 2> gpurun_out/r2_c5_512_n${N}_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_c5_512_n${N}_$TAG.json"))
print("C5-512 value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["exchange"])
print(d["stages"])
PY
tail -3 gpurun_out/r2_c5_512_n${N}_$TAG.err
fi
