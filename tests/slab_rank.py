"""torchrun worker of the multi-GPU slab parity test (tests/test_gpu_slab.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tests/slab_rank.py --case water|phantom --exchange auto|nccl|peer --out result.json

Every rank runs its slab through the C ABI; the planes are gathered on rank 0, which compares them with
the single-GPU solve of the same scene (cuFFT pipeline) and with the CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT / "openlifu-python_b200"), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def build_case(name, steps):
    from tests import cases
    if name.startswith("wide"):
        # 128 x 64 x 128 expanded (8 x 16 factorisations on x and z): the fused passes with routed stores on > 1 rank
        case = cases.wide_case((128, 64, 128), steps=steps)
        if name == "wide_phantom":
            case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
            case["dt"], case["t_end"] = 1.5e-7, steps * 1.5e-7
        return case
    case = cases.v2_small_case(steps=steps)                       # inner 40x44x36 -> 64^3 with PML
    if name == "phantom":
        case["c0"], case["rho0"], case["alpha"] = cases.layered_phantom(tuple(case["N"]))
        case["dt"], case["t_end"] = 1.5e-7, steps * 1.5e-7
    elif name == "lossy":
        case["alpha"], case["c0"], case["rho0"] = 0.75, 1540.0, 1050.0
    return case


def api_main(args, rank, world, local):
    """Public API: every rank calls run_simulation with the same arguments; the Dataset comes back whole."""
    import torch.distributed as dist
    from openlifu_b200.geo import Point
    from openlifu_b200.seg import seg_methods
    from openlifu_b200.sim import SimSetup, kwave_if
    from openlifu_b200.bf import delay_methods
    from openlifu_b200.xdc import Transducer
    from tests import cases
    os.environ["LIFU_DEVICE"] = str(local)
    arr = Transducer.gen_matrix_array(nx=2, ny=2, pitch=3, kerf=0.5, units="mm", sensitivity=1e5)
    setup = SimSetup(spacing=1, x_extent=(-20, 19), y_extent=(-22, 21), z_extent=(-3, 32), dt=3e-7, t_end=args.steps * 3e-7)
    params = setup.setup_sim_scene(seg_methods.UniformWater())
    if args.case == "phantom":
        c0, rho0, al = cases.layered_phantom(tuple(params["sound_speed"].data.shape))
        params["sound_speed"].data[...] = c0
        params["density"].data[...] = rho0
        params["attenuation"].data[...] = al
    delays = delay_methods.Direct().calc_delays(arr, Point(position=(0, 0, 18), units="mm"), params)
    kw = dict(arr=arr, params=params, delays=delays, apod=np.ones(4), freq=400e3, cycles=2, dt=setup.dt, t_end=setup.t_end)
    os.environ["LIFU_MULTI_GPU"] = "slab"
    os.environ["LIFU_SLAB_EXCHANGE"] = args.exchange
    ds, out = kwave_if.run_simulation(**kw)
    ds2, _ = kwave_if.run_simulation(**kw)                       # cached handle, second collective run
    os.environ["LIFU_MULTI_GPU"] = "foci"
    os.environ["LIFU_PIPELINE"] = "v1"
    one, _ = kwave_if.run_simulation(**kw)                       # this rank alone on its own GPU
    res = {"case": args.case, "world": world, "api": True, "shape": list(ds["p_min"].data.shape),
           "vs_single": {k: cases.rel_l2(ds[k].data, one[k].data) for k in ("p_max", "p_min", "intensity")},
           "repeat_equal": bool(np.array_equal(ds["p_min"].data, ds2["p_min"].data)),
           "finite": bool(np.isfinite(ds["p_max"].data).all())}
    allres = [None] * world
    dist.all_gather_object(allres, res)
    kwave_if.clear_sessions()
    if rank == 0:
        res["all_ranks_vs_single"] = [max(r["vs_single"].values()) for r in allres]
        print(json.dumps(res), flush=True)
        if args.out:
            Path(args.out).write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="water")
    ap.add_argument("--exchange", default="auto")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--api", action="store_true", help="go through openlifu_b200.sim.run_simulation (LIFU_MULTI_GPU=slab)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                               # plumbing only: id broadcast + result gather
    from openlifu_b200 import _lib
    from tests import cases
    if args.api:
        return api_main(args, rank, world, local)
    ids = [_lib.slab_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    case = build_case(args.case, args.steps)
    got = cases.run_cuda_case_slab(case, rank, world, ids[0], exchange=args.exchange, device=local)
    parts = [None] * world
    dist.gather_object({"lay": got["layout"], "p_max": got["p_max"], "p_min": got["p_min"], "stats": got["stats"]},
                       parts if rank == 0 else None, dst=0)
    if rank == 0:
        N = got["N"]
        full = {k: np.concatenate([p[k] for p in parts]) for k in ("p_max", "p_min")}
        assert full["p_max"].size == int(np.prod(N)), (full["p_max"].size, N)
        planes = [(p["lay"]["sensor_z0"], p["lay"]["sensor_nz"]) for p in parts]
        one = cases.run_cuda_case(case, pipeline="v1", device=local)
        res = {"case": args.case, "world": world, "exchange": parts[0]["lay"]["exchange"], "planes": planes,
               "n_exp": list(got["stats"]["n_exp"]), "loop_ms": [p["stats"]["loop_ms"] for p in parts],
               "vs_single": {k: cases.rel_l2(full[k], one[k]) for k in ("p_max", "p_min")},
               "finite": bool(np.isfinite(full["p_max"]).all() and np.isfinite(full["p_min"]).all())}
        if not args.no_oracle:
            want = cases.run_oracle_case(case)
            res["vs_oracle"] = {k: cases.rel_l2(full[k], want[k]) for k in ("p_max", "p_min")}
        print(json.dumps(res), flush=True)
        if args.out:
            Path(args.out).write_text(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
