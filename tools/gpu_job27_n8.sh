#!/bin/bash
# 8-GPU confirmation of the default configuration: bench line with its slab leg, C5 (768^3 over 8 GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n${N}_final.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("slab"))[:2500])
PY
tail -3 gpurun_out/r2_bench_n${N}_final.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --workload C5 --n-inner 728 --time-steps 20 --steps 2 --warmup 1 > gpurun_out/r2_c5_768_n${N}_final.json 2> gpurun_out/r2_c5_768_n${N}_final.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_c5_768_n${N}_final.json"))
print("C5 value", d["value"], "ms/step", d["ms_per_step"], d["roofline"]["step"], d["roofline"]["exchange"])
print(d["stages"])
PY
tail -3 gpurun_out/r2_c5_768_n${N}_final.err
