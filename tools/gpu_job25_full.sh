#!/bin/bash
# full GPU suite + smoke + default bench line at HEAD
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gputest_wide.log 2>&1
tail -6 gpurun_out/r2_gputest_wide.log
python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_final2.json")); r=d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel", r["kernel"], "frac", r["frac"], "step frac", r["step"]["frac"], "ms/step", r["step"]["ms_per_time_step"], "cpu", d["cpu_baseline"]["value"])
PY
tail -2 gpurun_out/r2_bench_final2.err
