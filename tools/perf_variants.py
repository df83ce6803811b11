"""Time-loop variants of the C2 / C3 grid on one B200, back to back in one process:

    python tools/perf_variants.py [C2|C3] [n_time_steps] VAR=VAL,VAR=VAL ...   (each argument is one variant: env settings)

For every variant: a fresh solver handle with the environment switches set, two warm-up runs, then the CUDA-event
time of the graph-replayed time loop (best of 3) and the per-stage profile with and without the source.
"""
import json
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "openlifu-python_b200"), str(ROOT)]
import numpy as np
from openlifu_b200 import _lib, configs
from openlifu_b200.sim.kwave_if import element_geometry, get_kgrid

args = sys.argv[1:]
wl = args.pop(0) if args and args[0] in ("C2", "C3", "C1") else "C2"
nt = int(args.pop(0)) if args and args[0].isdigit() else 200
variants = args or [""]
cfg = {"C1": configs.c1, "C2": configs.c2, "C3": configs.c3}[wl]()
params, foci, beams, cycles = configs.prepare(cfg)
kg = get_kgrid(params.coords)
arr = cfg["arr"]
t = np.arange(0, cycles / cfg["pulse"].frequency, kg["dt"])
base = np.sin(2 * np.pi * cfg["pulse"].frequency * t)
off = [-float(c.mean()) * 1e-3 for c in params.coords.values()]
geom = element_geometry(arr, off)
names = ("sound_speed", "density", "attenuation")
homog = all(float(params[k].data.min()) == float(params[k].data.max()) for k in names)
out = []
for var in variants:
    env = dict(kv.split("=", 1) for kv in var.split(",") if kv)
    for k, v in env.items():
        os.environ[k] = v
    sim = _lib.LifuSim(kg["N"], kg["d"], kg["dt"], min(nt, kg["Nt"]))
    if homog:
        sim.set_medium(*[float(params[k].data.flat[0]) for k in names])
    else:
        sim.set_medium(*[params[k].data for k in names])
    sim.set_elements(*geom, 0.05, 5)
    n_delay, gains, bg = arr.drive_plan(kg["dt"], *beams[0])
    sim.set_drive(base * bg, n_delay, gains)
    best = None
    for rep in range(5):
        pm, pn, st = sim.run()
        if rep >= 2:
            best = st["loop_ms"] if best is None else min(best, st["loop_ms"])
    rec = {"variant": var or "default", "workload": wl, "steps": st["steps"], "source_steps": st["source_steps"],
           "loop_ms": best, "ms_per_step": best / st["steps"], "checksum": float(np.abs(pm).sum(dtype=np.float64)),
           "fft_launches": st["fft_launches"]}
    rec["steady_source_steps"] = st.get("steady_source_steps", 0)
    for ws, tag in ((1, "src"), (0, "nosrc"), (2, "steady")):
        if ws == 2 and not rec["steady_source_steps"]:
            continue
        prof = sim.profile_stages(reps=5, with_source=ws)
        rec["stages_" + tag] = {n: round(ms, 4) for n, ms, b in prof}
        rec["step_%s_ms" % tag] = round(sum(ms for _, ms, _ in prof), 4)
    sim.close()
    for k in env:
        os.environ.pop(k, None)
    out.append(rec)
    print(json.dumps(rec), flush=True)
