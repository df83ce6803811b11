#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; cut -c 1-250 gpurun_out/r2_bench_default.json; tail -2 gpurun_out/r2_bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_default.json")); r=d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel", r["kernel"], "frac", r["frac"], "traffic", r["traffic"], "alg", r["algorithmic_bytes"], "step frac", r["step"]["frac"], "cpu", d["cpu_baseline"]["value"])
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputest_final2.log 2>&1
tail -4 gpurun_out/r2_gputest_final2.log
