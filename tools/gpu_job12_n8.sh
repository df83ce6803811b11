#!/bin/bash
# N-GPU scaling check of the bench line (foci weak scaling + slab leg), N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
cut -c 1-600 gpurun_out/r2_bench_n$N.json; python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n$N.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "slab", {k: d["slab"][k] for k in ("value","single_gpu_value","efficiency_vs_n1","rel_l2_vs_single","exchange","ms_per_time_step")})
PY
tail -3 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 2 --warmup 3 --poses 8 --no-slab-leg > gpurun_out/r2_bench_poses_n$N.json 2> gpurun_out/r2_bench_poses_n$N.err
cut -c 1-400 gpurun_out/r2_bench_poses_n$N.json; tail -2 gpurun_out/r2_bench_poses_n$N.err
if [ "$N" -ge 4 ]; then
  python -m pytest tests/test_gpu_slab.py -q -k "many_rank and 4" 2>&1 | tail -3
fi
