#!/bin/bash
# 8-GPU job: 8-rank slab parity on the fused passes, the bench line with its slab leg, C5 (768^3 over 8 GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 300 python -m pytest tests/test_gpu_slab.py -q -k "many_rank and $N and phantom" --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c 1-800
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n${N}_wide.json 2> gpurun_out/r2_bench_n${N}_wide.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n${N}_wide.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("slab"))[:2500])
PY
tail -3 gpurun_out/r2_bench_n${N}_wide.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --workload C5 --n-inner 728 --time-steps 20 --steps 2 --warmup 1 > gpurun_out/r2_c5_768_n${N}_wide.json 2> gpurun_out/r2_c5_768_n${N}_wide.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_c5_768_n${N}_wide.json"))
print("C5 value", d["value"], "ms/step", d["ms_per_step"], d["config"]["workload"], d["detail"], d["roofline"]["step"], d["roofline"]["exchange"])
print(d["stages"])
PY
tail -3 gpurun_out/r2_c5_768_n${N}_wide.err
