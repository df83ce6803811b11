#!/usr/bin/env python
"""FP32-vs-FP64 floor of the oracle at config scale, and the decimated FP64 golden the GPU tests compare with.

    python tools/parity_floor.py C2 [--steps 749] [--stride 4] [--out tests/golden/c2_oracle_f64_sub6.npz]
    python tools/parity_floor.py C3 --steps 240

CPU only (numpy / torch): runs the oracle's time loop (oracle/solver.py, torch backend = the same loop on all host
threads) twice on the headline grid (216^3 inner -> 256^3), in float32 and in float64, prints the relative L2 distance
between the two (the rounding floor any float32 implementation sits on) and stores every `stride`-th voxel of the float64
fields as a small committed fixture.  tests/test_gpu_config_scale.py compares the CUDA result with both the live float32
oracle (full grid) and this float64 golden (sub-lattice).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)

from tests import cases  # noqa: E402
from tests.config_cases import c2_case, c3_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["C2", "C3"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--stride", type=int, default=6)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    case = c2_case(args.steps) if args.workload == "C2" else c3_case(args.steps or 240)
    res = {}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        t0 = time.perf_counter()
        res[name] = cases.run_oracle_case(case, dtype=dt, backend="torch")
        print(f"oracle {name}: {res[name]['Nt']} steps in {time.perf_counter() - t0:.0f} s", flush=True)
    n = case["N"]
    floor = {k: cases.rel_l2(res["f32"][k], res["f64"][k]) for k in ("p_max", "p_min")}
    s = args.stride
    sub = {}
    for k in ("p_max", "p_min"):
        full = res["f64"][k].reshape(n, order="F")
        sub[k] = np.ascontiguousarray(full[::s, ::s, ::s])
    out = Path(args.out or ROOT / "tests" / "golden" / f"{args.workload.lower()}_oracle_f64_sub{s}.npz")
    meta = {"workload": args.workload, "steps": int(res["f64"]["Nt"]), "stride": s, "N": list(n),
            "floor_f32_vs_f64_rel_l2": floor, "n_src": int(res["f64"]["src_idx"].size),
            "src_idx_sum": int(res["f64"]["src_idx"].sum()), "generator": "tools/parity_floor.py"}
    np.savez_compressed(out, p_max=sub["p_max"], p_min=sub["p_min"], meta=json.dumps(meta))
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
