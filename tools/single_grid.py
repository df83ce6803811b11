"""One C5-class phantom grid on ONE GPU through the ordinary handle: python tools/single_grid.py n_inner time_steps [pipeline ...]"""
import json
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 728
ts = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ref = None
for pipe in (sys.argv[3:] or ["auto"]):
    if pipe == "auto":
        os.environ.pop("LIFU_PIPELINE", None)
    else:
        os.environ["LIFU_PIPELINE"] = pipe
    r = bench.single_measure(0, n, ts, steps=2, warmup=1, profile=True)
    rec = {"pipeline": pipe, "n_inner": n, "time_steps": ts, "Mvox_step_per_s": r["value"], "ms_per_time_step": r["ms_per_time_step"],
           "fft_launches": r["fft_launches"], "stages": r["stages"]}
    if ref is None:
        ref = r
    else:
        rec["rel_l2_vs_first"] = {k: bench.rel_l2(r[k], ref[k]) for k in ("p_max", "p_min")}
    print(json.dumps(rec), flush=True)
