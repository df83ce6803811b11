"""``Protocol.calc_solution``: per-focus beamform -> simulate -> stack -> scale -> aggregate -> analyze.

Mirrors /root/reference/src/openlifu/plan/protocol.py (``Protocol:33``, ``beamform:129-132``,
``check_target:207-224``, ``fix_pulse_mismatch:226-240``, ``calc_solution:242-398``): same
attributes, keyword arguments, defaults, errors and returned triple
``(Solution, aggregated Dataset, SolutionAnalysis)``.  ``run_simulation`` is bound at module level
exactly like the reference (``protocol.py:24``) so it can be patched by name.

What differs is the execution of the per-focus loop (``protocol.py:318-339``): the iterations are
independent, so with several B200s visible the foci are sharded one simulation per GPU
(focus i -> device i mod G, one worker thread per device; each device keeps its own solver handle,
medium and source weights -- no data-path collective).  With one device, or when
``run_simulation`` has been replaced (tests), the loop runs serially in the reference's order.
"""
from __future__ import annotations

import json
import logging
import math
import os
from concurrent.futures import ThreadPoolExecutor
from copy import deepcopy
from dataclasses import asdict, dataclass, field
from datetime import datetime
from enum import Enum
from pathlib import Path
from typing import Any, Dict, List, Tuple

import numpy as np
import pandas as pd

from .. import bf, geo, seg, sim, xa, xdc
from ..geo import Point
from ..sim import kwave_if, run_simulation
from ..util.checkgpu import gpu_available
from ..xdc import Transducer
from .param_constraint import ParameterConstraint
from .solution import Solution, _Encoder
from .solution_analysis import SolutionAnalysis, SolutionAnalysisOptions
from .target_constraints import TargetConstraints

OnPulseMismatchAction = Enum("OnPulseMismatchAction", ["ERROR", "ROUND", "ROUNDUP", "ROUNDDOWN"])


def _visible_devices() -> int:
    """Number of GPUs the foci may be sharded over (``LIFU_FOCI_GPUS`` caps it; 1 under torchrun,
    where each rank owns one device)."""
    if "LOCAL_RANK" in os.environ or "LIFU_DEVICE" in os.environ:
        return 1
    try:
        from pynvml import nvmlDeviceGetCount, nvmlInit, nvmlShutdown
        nvmlInit()
        n = nvmlDeviceGetCount()
        nvmlShutdown()
    except Exception:  # noqa: BLE001
        n = 1
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES")
    if cvd:
        n = min(n, len([x for x in cvd.split(",") if x.strip()]))
    cap = int(os.environ.get("LIFU_FOCI_GPUS", "0") or 0)
    return max(1, min(n, cap) if cap > 0 else n)


def _dist_world() -> Tuple[int, int]:
    """(world_size, rank) of the torch.distributed job this process belongs to, (1, 0) outside one."""
    try:
        import torch.distributed as dist
    except Exception:  # noqa: BLE001
        return 1, 0
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(), dist.get_rank()


def _gather_foci(mine: Dict[int, Any], n_foci: int, world: int, coords) -> list:
    """All-gather the per-focus Datasets of a rank-sharded sweep (focus i lives on rank i mod world).
    Every variable is exchanged in its own dtype (p_max/p_min float32, intensity float64), in slots
    of ceil(n_foci / world) foci per rank; the only collective of the sweep, outside the solver."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    template = next(iter(mine.values())) if mine else None
    meta = [None]
    if rank == 0:                       # rank 0 always owns focus 0
        meta = [[(k, template[k].data.dtype.str, tuple(template[k].data.shape), tuple(template[k].dims),
                  dict(template[k].attrs), template[k].name) for k in template.data_vars]]
    dist.broadcast_object_list(meta, src=0)
    slots = (n_foci + world - 1) // world
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", kwave_if._device()) if on_gpu else torch.device("cpu")   # the solver's GPU (LOCAL_RANK)
    gathered = {}
    for name, dtype, shape, _, _, _ in meta[0]:
        local = np.zeros((slots,) + shape, dtype=np.dtype(dtype))
        for i, ds in mine.items():
            local[i // world] = ds[name].data
        send = torch.from_numpy(local).to(dev)
        recv = torch.empty((world * slots,) + tuple(send.shape[1:]), dtype=send.dtype, device=dev)
        dist.all_gather_into_tensor(recv, send)                               # rank-major concatenation
        gathered[name] = recv.cpu().numpy().reshape((world, slots) + shape)
    out = []
    for i in range(n_foci):
        if i in mine:
            out.append(mine[i])
            continue
        vars_ = {}
        for name, _, _, dims, attrs, da_name in meta[0]:
            data = np.array(gathered[name][i % world, i // world])          # writable copy (Solution.scale works in place)
            vars_[name] = xa.DataArray(data, coords=coords, dims=dims, name=da_name, attrs=dict(attrs))
        out.append(xa.Dataset(vars_))
    return out


def candidate_transducer(transducer: Transducer, transform) -> Transducer:
    """The transducer placed at one candidate pose the way the reference does it: a ``TransformedTransducer`` carrying
    the 4x4 ``transform``, baked into element positions / orientations (``xdc/transducer.py:412-417``)."""
    from dataclasses import fields as dc_fields
    from ..xdc.transducer import TransformedTransducer
    kw = {f.name: deepcopy(getattr(transducer, f.name)) for f in dc_fields(Transducer)}
    return TransformedTransducer(transform=np.asarray(transform, dtype=np.float64), **kw).bake()


@dataclass
class Protocol:
    id: str = "protocol"
    name: str = "Protocol"
    description: str = ""
    allowed_roles: List[str] = field(default_factory=list)
    pulse: bf.Pulse = field(default_factory=bf.Pulse)
    sequence: bf.Sequence = field(default_factory=bf.Sequence)
    focal_pattern: bf.FocalPattern = field(default_factory=bf.focal_patterns.SinglePoint)
    sim_setup: sim.SimSetup = field(default_factory=sim.SimSetup)
    delay_method: bf.DelayMethod = field(default_factory=bf.delay_methods.Direct)
    apod_method: bf.ApodizationMethod = field(default_factory=bf.apod_methods.Uniform)
    seg_method: seg.SegmentationMethod = field(default_factory=seg.seg_methods.UniformWater)
    param_constraints: dict = field(default_factory=dict)
    target_constraints: List[TargetConstraints] = field(default_factory=list)
    analysis_options: SolutionAnalysisOptions = field(default_factory=SolutionAnalysisOptions)
    virtual_fit_options: Any = None      # virtual fit is out of scope here; the field is carried through untouched

    def __post_init__(self):
        self.logger = logging.getLogger(__name__)

    # ------------------------------------------------------------------------------ (de)serialisation
    @staticmethod
    def from_dict(d: Dict[str, Any]) -> "Protocol":
        d = dict(d)
        d["pulse"] = bf.Pulse.from_dict(d.get("pulse", {}))
        d["sequence"] = bf.Sequence.from_dict(d.get("sequence", {}))
        d["focal_pattern"] = bf.FocalPattern.from_dict(d.get("focal_pattern", {"class": "SinglePoint"}))
        d["sim_setup"] = sim.SimSetup.from_dict(d.get("sim_setup", {}))
        d["delay_method"] = bf.DelayMethod.from_dict(d.get("delay_method", {"class": "Direct"}))
        d["apod_method"] = bf.ApodizationMethod.from_dict(d.get("apod_method", {"class": "Uniform"}))
        seg_d = dict(d.get("seg_method", {"class": "UniformWater"}))
        if "materials" in d:
            mats = d.pop("materials")
            seg_d["materials"] = {k: (m if isinstance(m, seg.Material) else seg.Material.from_dict(m)) for k, m in mats.items()}
        d["seg_method"] = seg.SegmentationMethod.from_dict(seg_d)
        d["param_constraints"] = {k: (v if isinstance(v, ParameterConstraint) else ParameterConstraint.from_dict(v))
                                  for k, v in d.get("param_constraints", {}).items()}
        if "target_constraints" in d:
            d["target_constraints"] = [t if isinstance(t, TargetConstraints) else TargetConstraints.from_dict(t)
                                       for t in d["target_constraints"]]
        d["analysis_options"] = SolutionAnalysisOptions.from_dict(d.get("analysis_options", {}))
        return Protocol(**d)

    def to_dict(self):
        vf = self.virtual_fit_options
        return {
            "id": self.id,
            "name": self.name,
            "description": self.description,
            "allowed_roles": self.allowed_roles,
            "pulse": self.pulse.to_dict(),
            "sequence": self.sequence.to_dict(),
            "focal_pattern": self.focal_pattern.to_dict(),
            "sim_setup": asdict(self.sim_setup),
            "delay_method": self.delay_method.to_dict(),
            "apod_method": self.apod_method.to_dict(),
            "seg_method": self.seg_method.to_dict(),
            "param_constraints": {k: pc.to_dict() for k, pc in self.param_constraints.items()},
            "target_constraints": [tc.to_dict() for tc in self.target_constraints],
            "virtual_fit_options": vf.to_dict() if hasattr(vf, "to_dict") else vf,
            "analysis_options": self.analysis_options.to_dict(),
        }

    @staticmethod
    def from_file(filename):
        with open(filename) as f:
            return Protocol.from_dict(json.load(f))

    @staticmethod
    def from_json(json_string: str) -> "Protocol":
        return Protocol.from_dict(json.loads(json_string))

    def to_json(self, compact: bool) -> str:
        if compact:
            return json.dumps(self.to_dict(), separators=(",", ":"), cls=_Encoder)
        return json.dumps(self.to_dict(), indent=4, cls=_Encoder)

    def to_file(self, filename: str):
        Path(filename).parent.mkdir(parents=True, exist_ok=True)
        Path(filename).write_text(self.to_json(compact=False))

    def to_table(self) -> pd.DataFrame:
        parts = [pd.DataFrame.from_records([{"Category": "", "Name": n, "Value": v, "Unit": ""}
                                            for n, v in (("ID", self.id), ("Name", self.name), ("Description", self.description))])]

        def add(category, sub):
            sub = sub.copy()
            sub.insert(0, "Category", category)
            parts.append(sub)

        for cat, obj in (("Pulse", self.pulse), ("Sequence", self.sequence), ("Focal Pattern", self.focal_pattern),
                         ("Delay Method", self.delay_method), ("Apodization Method", self.apod_method),
                         ("Segmentation Method", self.seg_method), ("Simulation Setup", self.sim_setup)):
            add(cat, obj.to_table())
        for tc in self.target_constraints:
            add("Target Constraints", tc.to_table())
        for pid, pc in self.param_constraints.items():
            tp = pc.to_table()
            tp["Value"] = tp["Value"].str.replace("value", pid)
            add("Parameter Constraints", tp)
        return pd.concat(parts, ignore_index=True)

    # ------------------------------------------------------------------------------ planning
    def beamform(self, arr: xdc.Transducer, target: geo.Point, params):
        delays = self.delay_method.calc_delays(arr, target, params)
        apod = self.apod_method.calc_apodization(arr, target, params)
        return delays, apod

    def check_target(self, target: Point):
        if isinstance(target, list):
            raise ValueError(f"Input target {target} not supposed to be a list!")
        for tc in self.target_constraints:
            if tc.dim in target.dims:
                tc.check_bounds(target.get_position(dim=tc.dim, units=tc.units))

    def fix_pulse_mismatch(self, on_pulse_mismatch: OnPulseMismatchAction, foci: List[Point]):
        """Make ``sequence.pulse_count`` a multiple of the number of foci, in place."""
        n = len(foci)
        if on_pulse_mismatch is OnPulseMismatchAction.ERROR:
            raise ValueError(f"Pulse Count {self.sequence.pulse_count} is not a multiple of the number of foci {n}")
        ratio = self.sequence.pulse_count / n
        rounder = {OnPulseMismatchAction.ROUND: round, OnPulseMismatchAction.ROUNDUP: math.ceil,
                   OnPulseMismatchAction.ROUNDDOWN: math.floor}.get(on_pulse_mismatch)
        if rounder is not None:
            self.sequence.pulse_count = rounder(ratio) * n
        self.logger.warning(f"Pulse Count {self.sequence.pulse_count} is not a multiple of the number of foci {n}."
                            f"Rounding to {self.sequence.pulse_count}.")

    def _simulate_foci(self, transducer, params, beams, cycles, sim_options, voltage, use_gpu):
        """Run one simulation per focus.  Serial unless several GPUs are visible and the stock
        ``run_simulation`` is in place; then focus i runs on device i mod G."""
        return self._simulate_jobs([(transducer, d, a) for d, a in beams], params, cycles, sim_options, voltage, use_gpu)

    def _simulate_jobs(self, jobs, params, cycles, sim_options, voltage, use_gpu):
        """One simulation per job = (transducer, delays, apodization): the independent units of a sweep (foci of a
        pattern, candidate poses of a virtual fit).  Job i runs on rank i mod world of a torch.distributed job, or on
        device i mod G of this process; no data-path collective."""
        def one(i):
            arr, delays, apod = jobs[i]
            ds, _ = run_simulation(arr=arr, params=params, delays=delays, apod=apod, freq=self.pulse.frequency,
                                   cycles=cycles, dt=sim_options.dt, t_end=sim_options.t_end, cfl=sim_options.cfl,
                                   amplitude=self.pulse.amplitude * voltage, gpu=use_gpu)
            return ds

        world, rank = _dist_world()
        if world > 1 and kwave_if.multi_gpu_mode()[0] == "slab":
            world = 1                   # the ranks share every simulation (slab decomposition): same loop on all of them
        if world > 1 and len(jobs) > 1:
            # one process per GPU (torchrun): rank r simulates jobs r, r + world, ...; the fields are
            # all-gathered so that every rank holds the full stack, as the reference's serial loop would
            mine = {i: one(i) for i in range(rank, len(jobs), world)}
            return _gather_foci(mine, len(jobs), world, params.coords)
        n_dev = _visible_devices() if (use_gpu and run_simulation is kwave_if.run_simulation and len(jobs) > 1) else 1
        if n_dev <= 1:
            return [one(i) for i in range(len(jobs))]
        results: list = [None] * len(jobs)

        def worker(dev):
            with kwave_if.use_device(dev):
                for i in range(dev, len(jobs), n_dev):
                    self.logger.info(f"Simulate job {i} on GPU {dev}...")
                    results[i] = one(i)

        with ThreadPoolExecutor(max_workers=n_dev) as pool:
            for f in [pool.submit(worker, d) for d in range(n_dev)]:
                f.result()
        return results

    def simulate_candidates(self, target: Point, transducer: Transducer, transforms, volume=None,
                            sim_options: sim.SimSetup | None = None,
                            analysis_options: SolutionAnalysisOptions | None = None, use_gpu: bool | None = None,
                            voltage: float = 1.0, analyze: bool = True):
        """Simulate a batch of candidate transducer poses for ONE target.

        The reference's virtual fit (``virtual_fit.py:230-467``) returns its best ``top_n_candidates`` poses as 4x4
        transforms and stops there -- it never simulates them.  Here every candidate is placed with the reference's own
        mechanism, ``TransformedTransducer(transform=M).bake()`` (``xdc/transducer.py:412-417``: elements mapped by
        ``inv(M)``, ``transducer.py:297-301``), beamformed onto the target and simulated on the protocol's grid; the
        candidates are independent simulations and are sharded exactly like the foci of a pattern (candidate i -> rank /
        device i mod G).  The off-grid source weights are rebuilt on the GPU per pose (10 ms on the 256^3 grid); medium,
        solver handle and FFT tables are shared by all candidates.

        Returns a list with one ``(Solution, SolutionAnalysis | None)`` per transform, in input order; every Solution holds
        its baked transducer, the delays / apodizations and a one-focus ``simulation_result``."""
        if use_gpu is None:
            use_gpu = gpu_available()
            if not use_gpu:
                raise RuntimeError("Protocol.simulate_candidates: no B200-class CUDA device was found and openlifu_b200 has "
                                   "no CPU simulation path")
        sim_options = self.sim_setup if sim_options is None else sim_options
        analysis_options = self.analysis_options if analysis_options is None else analysis_options
        self.check_target(target)
        params = sim_options.setup_sim_scene(self.seg_method, volume=volume)
        cycles = np.min([np.round(self.pulse.duration * self.pulse.frequency), 20])
        arrays = [candidate_transducer(transducer, m) for m in transforms]
        jobs = []
        for arr in arrays:
            delays, apod = self.beamform(arr=arr, target=target, params=params)
            jobs.append((arr, delays, apod))
        outputs = self._simulate_jobs(jobs, params, cycles, sim_options, voltage, use_gpu)
        results = []
        for i, (ds, (arr, delays, apod)) in enumerate(zip(outputs, jobs)):
            stacked = xa.concat([ds.assign_coords(focal_point_index=0)], dim="focal_point_index")
            stamp = datetime.now().strftime("%Y%m%d_%H%M%S_%f")
            sol = Solution(id=f"candidate_{i}_{stamp}", name=f"Candidate {i}", protocol_id=self.id, transducer=arr,
                           delays=np.stack([delays], axis=0), apodizations=np.stack([apod], axis=0), pulse=self.pulse,
                           voltage=voltage, sequence=self.sequence, foci=[target], target=target, simulation_result=stacked,
                           approved=False, description=f"Candidate pose {i} of {len(jobs)} for target {target.id}")
            ana = sol.analyze(options=analysis_options, param_constraints=self.param_constraints) if analyze else None
            results.append((sol, ana))
        return results

    def _on_device_ok(self, use_gpu, n_foci) -> bool:
        """Can the whole plan stay on one GPU?  (stock ``run_simulation``, one process, one worker device)"""
        if not use_gpu or run_simulation is not kwave_if.run_simulation:
            return False
        if os.environ.get("LIFU_ANALYZE", "cuda") == "host" or kwave_if.multi_gpu_mode()[0] == "slab":
            return False                    # the device route analyses on the device and needs one whole grid per GPU
        if _dist_world()[0] > 1:
            return False
        return n_foci <= 1 or _visible_devices() <= 1 or os.environ.get("LIFU_FOCI_GPUS", "") == "1"

    def _simulate_foci_on_device(self, transducer, params, beams, cycles, sim_options, voltage):
        """One simulation per focus, every result packaged by the GPU into a device-resident stack (no host copies)."""
        from .. import _lib
        n = [len(c) for c in params.coords.values()]
        stack = _lib.FieldStack(n, len(beams), device=kwave_if._device())
        try:
            for i, (delays, apod) in enumerate(beams):
                self.logger.info(f"Simulate focus {i} (fields stay on the device)...")
                kwave_if.run_simulation_into(stack, i, arr=transducer, params=params, delays=delays, apod=apod,
                                             freq=self.pulse.frequency, cycles=cycles, dt=sim_options.dt,
                                             t_end=sim_options.t_end, cfl=sim_options.cfl,
                                             amplitude=self.pulse.amplitude * voltage, gpu=True)
        except Exception:
            stack.close()
            raise
        return stack

    def calc_solution(self, target: Point, transducer: Transducer, volume=None, session=None, simulate: bool = True,
                      scale: bool = True, sim_options: sim.SimSetup | None = None,
                      analysis_options: SolutionAnalysisOptions | None = None,
                      on_pulse_mismatch: OnPulseMismatchAction = OnPulseMismatchAction.ERROR,
                      use_gpu: bool | None = None, voltage: float = 1.0,
                      on_device: bool | None = None) -> Tuple[Solution, Any, SolutionAnalysis]:
        """Delays/apodizations per focus, simulated fields, scaling to the target pressure and
        the beam analysis (reference semantics, protocol.py:242-398).

        ``on_device`` (not in the reference; ``None`` -> on unless ``$LIFU_PLAN_ON_DEVICE`` == "0"): keep the fields of
        every focus in HBM from the solver through stacking, ``Solution.scale``, the aggregation over foci and both beam
        analyses (``_lib.FieldStack``, csrc/stack.cu), and copy the finished stack to the host once.  Same numbers, bit
        for bit, as the host route (tests/test_gpu_api.py); applies when one GPU runs the whole plan (one process, one
        worker device, stock ``run_simulation``), otherwise the host route is taken."""
        if use_gpu is None:
            use_gpu = gpu_available()
            if not use_gpu and simulate:
                # the reference falls back to the k-Wave OMP binary here (protocol.py:294-295); this package has no
                # CPU solver, so say what is missing instead of failing later inside run_simulation
                raise RuntimeError("Protocol.calc_solution: no B200-class CUDA device was found (NVML and liblifusim "
                                   "probes) and openlifu_b200 has no CPU simulation path; pass simulate=False to "
                                   "beamform only")
        sim_options = self.sim_setup if sim_options is None else sim_options
        analysis_options = self.analysis_options if analysis_options is None else analysis_options
        self.check_target(target)
        params = sim_options.setup_sim_scene(self.seg_method, volume=volume)

        foci: List[Point] = self.focal_pattern.get_targets(target)
        simulation_cycles = np.min([np.round(self.pulse.duration * self.pulse.frequency), 20])
        if (self.sequence.pulse_count % len(foci)) != 0:
            self.fix_pulse_mismatch(on_pulse_mismatch, foci)

        beams = []
        for focus in foci:
            self.logger.info(f"Beamform for focus {focus}...")
            beams.append(self.beamform(arr=transducer, target=focus, params=params))
        if on_device is None:
            on_device = os.environ.get("LIFU_PLAN_ON_DEVICE", "1") != "0"
        stacked = xa.Dataset()
        stack = None
        if simulate and on_device and self._on_device_ok(use_gpu, len(foci)):
            stack = self._simulate_foci_on_device(transducer, params, beams, simulation_cycles, sim_options, voltage)
            # metadata-only stand-in (zero-stride arrays) until the finished stack is copied out below
            shape = (len(foci),) + tuple(len(c) for c in params.coords.values())
            blank = kwave_if.result_dataset(params, np.broadcast_to(np.float32(0), shape), np.broadcast_to(np.float32(0), shape),
                                            np.broadcast_to(np.float64(0), shape), leading=("focal_point_index",))
            stacked = blank.assign_coords(focal_point_index=np.arange(len(foci)))
        elif simulate:
            outputs = self._simulate_foci(transducer, params, beams, simulation_cycles, sim_options, voltage, use_gpu)
            stacked = xa.concat([o.assign_coords(focal_point_index=i) for i, o in enumerate(outputs)],
                                dim="focal_point_index")

        timestamp = datetime.now().strftime("%Y%m%d_%H%M%S_%f")
        solution_id = timestamp if session is None else f"{session.id}_{timestamp}"
        description = (f"A solution computed for the {self.name} protocol with transducer {transducer.name}"
                       f" for target {target.id}."
                       f" This solution was created for the session {session.id} for subject {session.subject_id}."
                       if session is not None else "")
        solution = Solution(id=solution_id, name=f"Solution {timestamp}", protocol_id=self.id, transducer=transducer,
                            delays=np.stack([b[0] for b in beams], axis=0),
                            apodizations=np.stack([b[1] for b in beams], axis=0), pulse=self.pulse, voltage=voltage,
                            sequence=self.sequence, foci=foci, target=target, simulation_result=stacked, approved=False,
                            description=description)
        if stack is not None:
            solution._stack = stack                 # Solution.scale / analyze read and rescale the device-resident fields
        try:
            if scale:
                if not simulate:
                    msg = f"Cannot scale solution {solution.id} if simulation is not enabled!"
                    self.logger.error(msg=msg)
                    raise ValueError(msg)
                self.logger.info(f"Scaling solution {solution.id}...")
                solution.scale(self.focal_pattern, analysis_options=analysis_options)

            if not simulate:
                return solution, None, None
            # pressures: max over foci; intensity: mean over foci (protocol.py:382-392)
            res = solution.simulation_result
            aggregated = deepcopy(res.drop_dims("focal_point_index"))     # every field carries the dim: copy what is left
            if stack is not None:
                agg_pmax, agg_pnp, agg_int = stack.aggregate()
                for name, data in (("p_min", agg_pnp), ("p_max", agg_pmax), ("intensity", agg_int)):
                    v = res[name]
                    aggregated[name] = xa.DataArray(data, coords=params.coords, dims=tuple(params.dims), name=v.name,
                                                    attrs=dict(v.attrs))
            else:
                aggregated["p_min"] = res["p_min"].max(dim="focal_point_index", keep_attrs=True)
                aggregated["p_max"] = res["p_max"].max(dim="focal_point_index", keep_attrs=True)
                aggregated["intensity"] = res["intensity"].mean(dim="focal_point_index", keep_attrs=True)
            analysis = solution.analyze(options=analysis_options, param_constraints=self.param_constraints)
            if stack is not None:
                # the one device -> host copy of the plan: the finished (scaled) stack replaces the stand-in
                pm, pn, it = stack.get()
                full = kwave_if.result_dataset(params, pm, pn, it, leading=("focal_point_index",))
                solution.simulation_result = full.assign_coords(focal_point_index=np.arange(len(foci)))
            return solution, aggregated, analysis
        finally:
            if stack is not None:
                solution._stack = None
                stack.close()
