"""``run_simulation``: the drop-in boundary of the treatment-planning hot path.

Same name, keyword arguments, defaults, errors and return value as
/root/reference/src/openlifu/sim/kwave_if.py:80-146 (module name kept so that
``openlifu.sim.kwave_if.run_simulation`` / ``openlifu.plan.protocol.run_simulation`` can be
pointed here unchanged).  Instead of building k-wave-python objects, writing HDF5 and spawning
the kspaceFirstOrder binary, it drives the C ABI of ``liblifusim.so`` (include/lifusim.h):
hand-written sm_100a kernels around cuFFT on one B200.  There is NO CPU path: ``gpu=False``
raises.

What is reused between calls (the reference recomputes all of it per focus):
solver handle + FFT plans per grid, off-grid source weights per (transducer, grid), medium maps
per params object.  Only ``(delays, apod)`` -- 2 x n_elements numbers -- go to the GPU per focus.
"""
from __future__ import annotations

import contextlib
import logging
import os
import threading
import zlib
from typing import List

import numpy as np

from .. import _lib, xa
from ..util.content import content_key as _content_key
from ..util.units import getunitconversion

log = logging.getLogger(__name__)

_SESSIONS: dict = {}
_MAX_SESSIONS = 2          # cached solver handles per device
_LOCK = threading.Lock()
_TLS = threading.local()


@contextlib.contextmanager
def use_device(device: int):
    """Pin ``run_simulation`` calls made by this thread to one GPU (foci sharding: one worker
    thread per device).  Handles on distinct devices run concurrently."""
    prev = getattr(_TLS, "device", None)
    _TLS.device = int(device)
    try:
        yield
    finally:
        _TLS.device = prev


def _same_units(objs, what):
    units = [o.attrs["units"] for o in objs]
    if not all(u == units[0] for u in units):
        raise ValueError(f"All {what} must have the same units")
    return units[0]


def get_kgrid(coords, t_end=0, dt=0, sound_speed_ref=1500, cfl=0.5):
    """Grid sizes, spacings [m] and time axis (kwave_if.py:13-27).  Returns a dict instead of a
    kWaveGrid.  Auto time stepping uses ``sound_speed_ref`` (default 1500), not the medium."""
    scl = getunitconversion(_same_units([coords[d] for d in coords.dims], "coordinates"), "m")
    sz = [len(c) for c in coords.values()]
    dx = [float(np.diff(c.data)[0] * scl) for c in coords.values()]
    if dt == 0 or t_end == 0:
        nt, dt_ = _lib.make_time(sz, dx, float(sound_speed_ref), float(cfl))
    else:
        nt, dt_ = int(round(t_end / dt)), float(dt)
    return {"N": sz, "d": dx, "Nt": nt, "dt": dt_}


def _device():
    if getattr(_TLS, "device", None) is not None:
        return _TLS.device
    if "LIFU_DEVICE" in os.environ:
        return int(os.environ["LIFU_DEVICE"])
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    return 0


def multi_gpu_mode():
    """How a torch.distributed job uses its ranks (``LIFU_MULTI_GPU``): ``"foci"`` (default) -- every
    ``run_simulation`` call is one rank's own simulation, sweeps are sharded by ``Protocol``; ``"slab"``
    -- every rank calls ``run_simulation`` with the SAME arguments and the grid is decomposed into z
    slabs over the ranks (grids too large for one GPU).  Returns (mode, world, rank)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            mode = os.environ.get("LIFU_MULTI_GPU", "foci")
            if mode not in ("foci", "slab"):
                raise ValueError(f"LIFU_MULTI_GPU must be 'foci' or 'slab', got {mode!r}")
            return mode, dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return "foci", 1, 0


class _Session:
    """Solver handle and what is cached on it."""

    def __init__(self, key, kg, device, slab=None):
        self.key = key
        self.sim = _lib.LifuSim(kg["N"], kg["d"], kg["dt"], kg["Nt"], device=device, slab=slab)
        self.geometry_key = None
        self.medium_key = None
        self.uniform = {}              # medium key -> "all three maps are constant" (decided once per maps object)
        self.two_z_key = None          # which density / sound-speed maps the device-side impedance was made from
        self.z_uniform = {}
        self.n_src = 0


def _session(kg, device, slab_world=1, slab_rank=0):
    key = (tuple(kg["N"]), tuple(kg["d"]), kg["dt"], kg["Nt"], slab_world, device)
    with _LOCK:
        s = _SESSIONS.get(key)
        if s is None:
            mine = [k for k in _SESSIONS if k[-1] == device]
            while len(mine) >= _MAX_SESSIONS:
                _SESSIONS.pop(mine.pop(0)).sim.close()
            slab = None
            if slab_world > 1:
                # collective: every rank reaches this point with the same key (same arguments by contract)
                import torch.distributed as dist
                ids = [_lib.slab_unique_id() if slab_rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                slab = (slab_rank, slab_world, ids[0], os.environ.get("LIFU_SLAB_EXCHANGE", "auto"))
            s = _SESSIONS[key] = _Session(key, kg, device, slab=slab)
    return s


def _gather_planes(local, layout, n, world):
    """All-gather the ranks' inner planes of a sensor vector (ragged along z) into the full x-fastest vector."""
    import torch
    import torch.distributed as dist
    plane = n[0] * n[1]
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", _device()) if on_gpu else torch.device("cpu")    # the solver's GPU, not torch's current one
    meta = torch.tensor([layout["sensor_z0"], layout["sensor_nz"]], dtype=torch.int64, device=dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [tuple(int(v) for v in m.tolist()) for m in metas]
    cap = max(nz for _, nz in metas) * plane
    send = torch.zeros(cap, dtype=torch.float32, device=dev)
    send[:local.size] = torch.from_numpy(local).to(dev)
    recv = torch.empty(world * cap, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, cap)
    out = np.empty(plane * n[2], dtype=np.float32)
    for r, (z0, nz) in enumerate(metas):
        out[z0 * plane:(z0 + nz) * plane] = recv[r, :nz * plane]
    return out


def clear_sessions():
    """Release every cached solver handle (GPU memory, FFT plans)."""
    with _LOCK:
        while _SESSIONS:
            _SESSIONS.popitem()[1].sim.close()


def element_geometry(arr, translation_m):
    """What get_karray (kwave_if.py:29-47) feeds kWaveArray.add_rect_element: centres [m] shifted
    by the array offset, (width, length) [m], (el, az, roll) in degrees."""
    pos = np.array([el.get_position(units="m") for el in arr.elements], dtype=np.float64) + np.asarray(translation_m)
    size = np.array([el.get_size(units="m") for el in arr.elements], dtype=np.float64)
    ang = np.array([el.get_angle(units="deg") for el in arr.elements], dtype=np.float64)
    return pos, size, ang


def drive_plan(arr, dt, delays, apod):
    """Integer delay samples, per-element gains and the base-signal gain that ``Transducer.calc_output`` encodes
    (xdc/transducer.py:95-112, xdc/element.py:144-154).  Uses ``arr.drive_plan`` when the transducer has it (this
    package's mirror); a reference ``openlifu.xdc.Transducer`` is read through its public attributes."""
    if hasattr(arr, "drive_plan"):
        return arr.drive_plan(dt, delays, apod)
    if getattr(arr, "impulse_response", None) is not None:
        raise NotImplementedError("array impulse responses are not supported on the simulation path")
    n_delay = np.array([int(d / dt) for d in delays], dtype=np.int32)
    gains = []
    for a, el in zip(apod, arr.elements):
        g = float(a)
        ir = getattr(el, "impulse_response", None)
        if ir is not None:
            if len(ir) != 1:
                raise NotImplementedError("array impulse responses are not supported on the simulation path")
            g *= float(ir[0])
        if getattr(el, "sensitivity", None) is not None:
            g *= float(el.sensitivity)
        gains.append(g)
    sens = getattr(arr, "sensitivity", None)
    return n_delay, np.array(gains, dtype=np.float64), (1.0 if sens is None else float(sens))


def _setup_run(arr, params, delays, apod, freq, cycles, amplitude, dt, t_end, cfl, bli_tolerance, upsampling_rate, gpu,
               ref_values_only):
    """Everything of ``run_simulation`` up to the solve: grid + time axis, drive, medium, source geometry on the cached
    solver handle.  Returns (session, kgrid dict, delay samples, slab?, world, medium_changed)."""
    if not gpu:
        raise RuntimeError("openlifu_b200.run_simulation has no CPU path (gpu=False): the solve runs on a B200 "
                           "through liblifusim.so only")
    n_el = arr.numelements()
    delays = np.zeros(n_el) if delays is None else delays
    apod = np.ones(n_el) if apod is None else apod
    kg = get_kgrid(params.coords, dt=dt, t_end=t_end, cfl=cfl)
    t = np.arange(0, cycles / freq, kg["dt"])
    input_signal = amplitude * np.sin(2 * np.pi * freq * t)
    n_delay, gains, base_gain = drive_plan(arr, kg["dt"], delays, apod)
    scl = getunitconversion(_same_units([params[d] for d in params.dims], "dimensions"), "m")
    array_offset: List[float] = [-float(c.mean()) * scl for c in params.coords.values()]

    mode, world, rank = multi_gpu_mode()
    slab = mode == "slab"
    ses = _session(kg, _device(), world if slab else 1, rank)
    sim = ses.sim
    # medium (get_medium, kwave_if.py:49-63): alpha_power 0.9; alpha_mode='no_dispersion' is what the
    # reference asks for but the k-Wave binary only receives alpha_coeff/alpha_power (ledger A7)
    alpha_mode = os.environ.get("LIFU_ALPHA_MODE", "binary")
    names = ("sound_speed", "density", "attenuation")
    if ref_values_only:
        mkey = ("ref",) + tuple(float(params[k].attrs["ref_value"]) for k in names) + (alpha_mode,)
        if ses.medium_key != mkey:
            sim.set_medium(*[float(params[k].attrs["ref_value"]) for k in names], alpha_power=0.9, alpha_mode=alpha_mode)
    else:
        maps = [params[k].data for k in names]
        mkey = ("map",) + tuple(_content_key(m) for m in maps) + (alpha_mode,)
        if ses.medium_key != mkey:
            if mkey not in ses.uniform:
                ses.uniform = {mkey: all(float(m.min()) == float(m.max()) for m in maps)}
            if ses.uniform[mkey]:
                sim.set_medium(*[float(m.flat[0]) for m in maps], alpha_power=0.9, alpha_mode=alpha_mode)
            elif slab:
                lo, nz = sim.layout["medium_z0"], sim.layout["medium_nz"]      # only the planes this rank reads
                sim.set_medium(*[m[:, :, lo:lo + nz] for m in maps], alpha_power=0.9, alpha_mode=alpha_mode, plane0=lo)
            else:
                lm = params.attrs.get("lifu_label_medium") if hasattr(params, "attrs") else None
                if (lm is not None and os.environ.get("LIFU_MEDIUM_LABELS", "1") != "0"
                        and lm["keys"] == tuple(mkey[1:4]) and lm["labels"].shape == tuple(kg["N"])):
                    # the maps are still the ones _map_params expanded from this label volume: upload the labels (one
                    # byte per voxel) and the per-label tables, expand on the device (seg_method.py:84-97)
                    sim.set_medium_labels(lm["labels"], *[lm["lut"][k] for k in names], alpha_power=0.9, alpha_mode=alpha_mode)
                else:
                    sim.set_medium(*maps, alpha_power=0.9, alpha_mode=alpha_mode)
    medium_changed = ses.medium_key != mkey
    ses.medium_key = mkey
    # source geometry (get_karray + get_array_binary_mask + BLI weights), cached per transducer
    pos, size, ang = element_geometry(arr, array_offset)
    gkey = (zlib.adler32(pos.tobytes()), zlib.adler32(size.tobytes()), zlib.adler32(ang.tobytes()),
            float(bli_tolerance), int(upsampling_rate))
    if ses.geometry_key != gkey:
        log.info("Computing off-grid source weights on the GPU")
        ses.n_src = sim.set_elements(pos, size, ang, bli_tolerance, upsampling_rate)
        ses.geometry_key = gkey
    sim.set_drive(input_signal * base_gain, n_delay, gains,
                  source_mode=os.environ.get("LIFU_SOURCE_MODE", "additive"))
    return ses, kg, n_delay, slab, world, medium_changed


def _set_impedance(ses, params, medium_changed):
    """2 * density * sound_speed of the packaging step on the device: always from the params maps (kwave_if.py:140),
    whatever medium was used."""
    rho, c = params["density"].data, params["sound_speed"].data
    zkey = (_content_key(rho), _content_key(c))
    if medium_changed or ses.two_z_key != zkey:
        if zkey not in ses.z_uniform:
            ses.z_uniform = {zkey: float(rho.min()) == float(rho.max()) and float(c.min()) == float(c.max())}
        if ses.z_uniform[zkey]:
            ses.sim.set_two_z(2 * (rho.flat[0] * c.flat[0]))
        else:
            ses.sim.set_two_z(_two_z_flat(params))
        ses.two_z_key = zkey


def run_simulation(arr,
                   params,
                   delays: np.ndarray | None = None,
                   apod: np.ndarray | None = None,
                   freq: float = 1e6,
                   cycles: float = 20,
                   amplitude: float = 1,
                   dt: float = 0,
                   t_end: float = 0,
                   cfl: float = 0.5,
                   bli_tolerance: float = 0.05,
                   upsampling_rate: int = 5,
                   gpu: bool = True,
                   ref_values_only: bool = False):
    ses, kg, n_delay, slab, world, medium_changed = _setup_run(arr, params, delays, apod, freq, cycles, amplitude, dt, t_end,
                                                               cfl, bli_tolerance, upsampling_rate, gpu, ref_values_only)
    sim = ses.sim
    log.info("Running simulation")
    # LIFU_PACKAGING=device computes -p_min and the intensity on the GPU (lifu_get_packaged, bit-identical); the default
    # stays on the host: the extra 8 bytes per voxel of device->host copy into fresh pageable memory cost more than the
    # threaded host expression saves (bench.py e2e on C2: 642 ms vs 761 ms per call)
    if slab or os.environ.get("LIFU_PACKAGING", "host") != "device":
        p_max_flat, p_min_flat, stats = sim.run()
        if slab:
            p_max_flat = _gather_planes(p_max_flat, sim.layout, kg["N"], world)
            p_min_flat = _gather_planes(p_min_flat, sim.layout, kg["N"], world)
        log.info("Simulation Complete")
        output = {"p_max": p_max_flat, "p_min": p_min_flat, "stats": stats, "n_src": ses.n_src,
                  "Nt": kg["Nt"], "dt": kg["dt"], "delay_samples": n_delay}
        return package_fields(params, output["p_max"], output["p_min"]), output
    _set_impedance(ses, params, medium_changed)
    p_max_flat, pnp_flat, inten_flat, stats = sim.run_packaged()
    log.info("Simulation Complete")
    output = _Output({"p_max": p_max_flat, "pnp": pnp_flat, "stats": stats, "n_src": ses.n_src,
                      "Nt": kg["Nt"], "dt": kg["dt"], "delay_samples": n_delay})
    return package_arrays(params, p_max_flat, pnp_flat, inten_flat), output


def run_simulation_into(stack, focus: int, *, arr, params, delays=None, apod=None, freq=1e6, cycles=20, amplitude=1, dt=0,
                        t_end=0, cfl=0.5, bli_tolerance=0.05, upsampling_rate=5, gpu=True, ref_values_only=False):
    """``run_simulation`` whose result stays on the device: the same solve, packaged by the GPU into slot ``focus`` of a
    ``_lib.FieldStack`` (p_max, -p_min, intensity -- the arrays of the Dataset ``run_simulation`` returns, bit for bit)
    instead of being copied to the host.  Used by ``Protocol.calc_solution(on_device=True)``; returns the solver stats."""
    ses, kg, n_delay, slab, world, medium_changed = _setup_run(arr, params, delays, apod, freq, cycles, amplitude, dt, t_end,
                                                               cfl, bli_tolerance, upsampling_rate, gpu, ref_values_only)
    if slab:
        raise RuntimeError("run_simulation_into: not available for slab-decomposed solves")
    if stack.device != _device():
        raise ValueError(f"run_simulation_into: the stack lives on device {stack.device}, the solver on {_device()}")
    _set_impedance(ses, params, medium_changed)
    log.info("Running simulation")
    stats = ses.sim.run_resident()                      # no host copy of the raw sensor vectors
    stack.put(focus, ses.sim)
    log.info("Simulation Complete")
    return {"stats": stats, "n_src": ses.n_src, "Nt": kg["Nt"], "dt": kg["dt"], "delay_samples": n_delay}


def result_dataset(params, p_max, pnp, intensity, leading=()):
    """The Dataset of kwave_if.py:131-146 from arrays already in their final form; ``leading``: extra leading dims
    (e.g. ("focal_point_index",) for a stack of foci)."""
    dims = tuple(leading) + tuple(params.dims)
    return xa.Dataset({
        "p_max": xa.DataArray(p_max, coords=params.coords, dims=dims, name="p_max", attrs={"units": "Pa", "long_name": "PPP"}),
        "p_min": xa.DataArray(pnp, coords=params.coords, dims=dims, name="p_min", attrs={"units": "Pa", "long_name": "PNP"}),
        "intensity": xa.DataArray(intensity, coords=params.coords, dims=dims, name="I",
                                  attrs={"units": "W/cm^2", "long_name": "Intensity"})})


class _Output(dict):
    """The opaque second return value (callers bind it to ``_``, protocol.py:324).  ``"p_min"`` -- the raw minimum the
    k-Wave binary would have returned -- is derived from the stored sign-flipped field only when somebody asks."""

    def __missing__(self, key):
        if key == "p_min":
            self[key] = -1 * self["pnp"]
            return self[key]
        raise KeyError(key)


def package_arrays(params, p_max_flat, pnp_flat, inten_flat):
    """Flat x-fastest vectors already in their final form (device packaging) -> the Dataset of kwave_if.py:131-146."""
    sz = list(params.coords.sizes.values())
    p_max = xa.DataArray(p_max_flat.reshape(sz, order="F"), coords=params.coords, name="p_max",
                         attrs={"units": "Pa", "long_name": "PPP"})
    p_min = xa.DataArray(pnp_flat.reshape(sz, order="F"), coords=params.coords, name="p_min",
                         attrs={"units": "Pa", "long_name": "PNP"})
    intensity = xa.DataArray(inten_flat.reshape(sz, order="F"), coords=params.coords, name="I",
                             attrs={"units": "W/cm^2", "long_name": "Intensity"})
    return xa.Dataset({"p_max": p_max, "p_min": p_min, "intensity": intensity})


_Z2_CACHE: dict = {}


def _two_z_flat(params):
    """2 * density * sound_speed (float64) flattened x fastest, cached per CONTENT of the params maps."""
    rho, c = params["density"].data, params["sound_speed"].data
    key = (_content_key(rho), _content_key(c))
    hit = _Z2_CACHE.get(key)
    if hit is None:
        if len(_Z2_CACHE) >= 4:
            _Z2_CACHE.clear()
        hit = _Z2_CACHE[key] = np.ascontiguousarray((2 * (rho * c)).transpose(2, 1, 0)).reshape(-1)
    return hit


_THREADS_SET = False


def _host_threads(torch):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the packaging expression is memory-bound host work, so give
    each rank its share of the host cores (once)."""
    global _THREADS_SET
    if _THREADS_SET:
        return
    _THREADS_SET = True
    local = int(os.environ.get("LOCAL_WORLD_SIZE", "0") or 0)
    if local >= 1 and torch.get_num_threads() == 1:
        torch.set_num_threads(max(1, min(16, (os.cpu_count() or 1) // local)))


def package_fields(params, p_max_flat, p_min_flat):
    """Flat x-fastest float32 sensor vectors -> the Dataset of kwave_if.py:131-146: p_max; p_min = -p_min;
    intensity = 1e-4 * p_min**2 / (2 * density * sound_speed) as float64 named 'I' under the key 'intensity'.
    Same IEEE operations in the same precisions as the reference's numpy expression (float32 square and
    scale, float64 divide), evaluated on flat x-fastest vectors with all host threads."""
    sz = list(params.coords.sizes.values())
    try:
        import torch
        _host_threads(torch)
        # chunked, with the results written straight into their final arrays: only the 12 bytes per voxel that are
        # handed out touch fresh pages (the float32 / float64 intermediates live in a few recycled megabytes)
        n = p_min_flat.size
        neg = np.empty(n, dtype=np.float32)
        inten = np.empty(n, dtype=np.float64)
        tp, tneg, tint = torch.from_numpy(p_min_flat), torch.from_numpy(neg), torch.from_numpy(inten)
        tz = torch.from_numpy(_two_z_flat(params))
        scale = np.float32(1e-4)
        for lo in range(0, n, 1 << 20):
            hi = min(n, lo + (1 << 20))
            x = tp[lo:hi]
            torch.neg(x, out=tneg[lo:hi])
            sq = torch.square(x)
            sq.mul_(scale)
            torch.div(sq.double(), tz[lo:hi], out=tint[lo:hi])
    except ImportError:
        neg = -1 * p_min_flat
        inten = (np.float32(1e-4) * p_min_flat ** 2) / _two_z_flat(params)
    p_max = xa.DataArray(p_max_flat.reshape(sz, order="F"), coords=params.coords, name="p_max",
                         attrs={"units": "Pa", "long_name": "PPP"})
    p_min = xa.DataArray(neg.reshape(sz, order="F"), coords=params.coords, name="p_min",
                         attrs={"units": "Pa", "long_name": "PNP"})
    intensity = xa.DataArray(inten.reshape(sz, order="F"), coords=params.coords, name="I",
                             attrs={"units": "W/cm^2", "long_name": "Intensity"})
    return xa.Dataset({"p_max": p_max, "p_min": p_min, "intensity": intensity})
