"""ctypes binding of liblifusim.so (the C ABI in include/lifusim.h).

No CPU fallback: importing this module without the built library, or creating a solver
without a B200, raises.  Pointers handed to the library may be numpy (host) buffers or raw
device pointers (ints, e.g. ``torch.Tensor.data_ptr()``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("LIFUSIM_LIB", _HERE.parent / "lib" / "liblifusim.so"))


class LifuError(RuntimeError):
    pass


class lifu_grid(C.Structure):
    _fields_ = [("n", C.c_int32 * 3), ("pml", C.c_int32 * 3), ("d", C.c_double * 3), ("dt", C.c_double),
                ("nt", C.c_int32), ("pml_alpha", C.c_double), ("c_ref", C.c_double)]


class lifu_stats(C.Structure):
    _fields_ = [("voxels", C.c_int64), ("n_exp", C.c_int32 * 3), ("pml", C.c_int32 * 3), ("steps", C.c_int32),
                ("source_steps", C.c_int32), ("kernel_launches", C.c_int64), ("fft_launches", C.c_int64),
                ("loop_ms", C.c_double), ("setup_ms", C.c_double), ("bytes_per_voxel_step", C.c_double),
                ("homogeneous", C.c_int32), ("absorbing", C.c_int32), ("steady_source_steps", C.c_int32),
                ("reserved", C.c_int32)]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out


class lifu_slab_desc(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("exchange", C.c_int32), ("nccl_id", C.c_ubyte * 128)]


class lifu_slab_layout(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("exchange", C.c_int32), ("z0", C.c_int32), ("nz", C.c_int32),
                ("sensor_z0", C.c_int32), ("sensor_nz", C.c_int32), ("medium_z0", C.c_int32), ("medium_nz", C.c_int32)]


class lifu_focus_query(C.Structure):
    _fields_ = [("w", (C.c_double * 4) * 3), ("aspect", C.c_double * 3), ("mainlobe_radius", C.c_double),
                ("sidelobe_radius", C.c_double), ("centroid_factor", C.c_double), ("pnp_scale", C.c_float),
                ("n_line", C.c_int32 * 3)]


class lifu_focus_metrics(C.Structure):
    _fields_ = [("main_pnp", C.c_double), ("side_pnp", C.c_double), ("global_pnp", C.c_double),
                ("main_ipa", C.c_double), ("side_ipa", C.c_double), ("global_ipa", C.c_double),
                ("main_ipa_all", C.c_double), ("global_ipa_all", C.c_double),
                ("n_main", C.c_int64), ("n_side", C.c_int64), ("n_global", C.c_int64),
                ("cen_w", C.c_double), ("cen_wx", C.c_double), ("cen_wy", C.c_double), ("cen_wz", C.c_double),
                ("n_centroid", C.c_int64), ("kernel_ms", C.c_double), ("reduce_ms", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


EXCHANGE_MODES = {"auto": 0, "nccl": 1, "peer": 2}
ALPHA_MODES = {"binary": 0, "no_dispersion": 1, "no_absorption": 2}
SOURCE_MODES = {"additive": 0, "additive-no-correction": 1}

_lib = None


def load():
    """Load the shared library once; raise a clear error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LifuError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). There is no CPU fallback for the simulation path.")
    if "LIFU_NCCL_LIB" not in os.environ:
        # slab decomposition dlopen()s NCCL: point it at the copy PyTorch bundles (a process can hold only one
        # libnccl.so.2, and torch cannot be imported on top of an older system NCCL)
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for base in (spec.submodule_search_locations if spec else []):
                cand = Path(base) / "lib" / "libnccl.so.2"
                if cand.exists():
                    os.environ["LIFU_NCCL_LIB"] = str(cand)
                    break
        except Exception:  # noqa: BLE001
            pass
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "lifu_abi_version": (C.c_int, []),
        "lifu_last_error": (C.c_char_p, []),
        "lifu_make_time": (C.c_int, [C.POINTER(i32), C.POINTER(f64), f64, f64, C.POINTER(i32), C.POINTER(f64)]),
        "lifu_pml_auto": (C.c_int, [C.POINTER(i32), C.POINTER(i32)]),
        "lifu_device_count": (C.c_int, [C.POINTER(i32)]),
        "lifu_create": (C.c_int, [C.POINTER(lifu_grid), C.c_int, vp, C.POINTER(vp)]),
        "lifu_destroy": (C.c_int, [vp]),
        "lifu_set_medium": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int, C.c_int]),
        "lifu_set_medium_f64": (C.c_int, [vp, vp, vp, vp, C.POINTER(i64), C.c_float, C.c_int]),
        "lifu_set_medium_labels": (C.c_int, [vp, vp, C.POINTER(i64), i32, vp, vp, vp, C.c_float, C.c_int]),
        "lifu_set_elements": (C.c_int, [vp, i32, vp, vp, vp, f64, i32, C.POINTER(i64)]),
        "lifu_set_source_geometry": (C.c_int, [vp, vp, vp, vp, vp, i64, i64, i32]),
        "lifu_get_source_sizes": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]),
        "lifu_get_source_geometry": (C.c_int, [vp, vp, vp, vp, vp]),
        "lifu_set_drive": (C.c_int, [vp, vp, i32, vp, vp, i32, C.c_int]),
        "lifu_run": (C.c_int, [vp, vp, vp, C.POINTER(lifu_stats)]),
        "lifu_set_two_z": (C.c_int, [vp, vp, i64]),
        "lifu_get_packaged": (C.c_int, [vp, vp, vp, vp]),
        "lifu_get_field": (C.c_int, [vp, C.c_int, vp]),
        "lifu_get_info": (C.c_int, [vp, C.POINTER(lifu_stats)]),
        "lifu_profile_stages": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(f64),
                                          C.POINTER(f64), C.POINTER(C.c_int)]),
        "lifu_slab_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
        "lifu_create_slab": (C.c_int, [C.POINTER(lifu_grid), C.c_int, vp, C.POINTER(lifu_slab_desc), C.POINTER(vp)]),
        "lifu_slab_layout_of": (C.c_int, [vp, C.POINTER(lifu_slab_layout)]),
        "lifu_set_medium_planes": (C.c_int, [vp, vp, vp, vp, C.c_float, C.c_int, C.c_int, i32, i32]),
        "lifu_analysis_create": (C.c_int, [C.c_int, vp, C.POINTER(i32), i32, vp, vp, vp, vp, C.POINTER(vp)]),
        "lifu_analysis_set_focus": (C.c_int, [vp, i32, vp, vp, C.POINTER(i64)]),
        "lifu_analysis_run_focus": (C.c_int, [vp, i32, C.POINTER(lifu_focus_query), vp, C.POINTER(lifu_focus_metrics), vp]),
        "lifu_analysis_destroy": (C.c_int, [vp]),
        "lifu_stack_create": (C.c_int, [C.c_int, vp, C.POINTER(i32), i32, C.POINTER(vp)]),
        "lifu_stack_destroy": (C.c_int, [vp]),
        "lifu_stack_put": (C.c_int, [vp, i32, vp]),
        "lifu_stack_scale": (C.c_int, [vp, i32, f64, f64]),
        "lifu_stack_pointers": (C.c_int, [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "lifu_stack_get": (C.c_int, [vp, i32, vp, vp, vp]),
        "lifu_stack_aggregate": (C.c_int, [vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.lifu_abi_version() != 1:
        raise LifuError(f"liblifusim ABI {lib.lifu_abi_version()} != 1")
    _lib = lib
    return lib


EXPORTED = ["lifu_abi_version", "lifu_last_error", "lifu_make_time", "lifu_pml_auto", "lifu_device_count", "lifu_create", "lifu_destroy",
            "lifu_set_medium", "lifu_set_medium_f64", "lifu_set_medium_labels", "lifu_set_elements", "lifu_set_source_geometry", "lifu_get_source_sizes",
            "lifu_get_source_geometry", "lifu_set_drive", "lifu_run", "lifu_set_two_z", "lifu_get_packaged", "lifu_get_field", "lifu_get_info",
            "lifu_profile_stages", "lifu_slab_unique_id", "lifu_create_slab", "lifu_slab_layout_of",
            "lifu_set_medium_planes", "lifu_analysis_create", "lifu_analysis_set_focus", "lifu_analysis_run_focus",
            "lifu_analysis_destroy", "lifu_stack_create", "lifu_stack_destroy", "lifu_stack_put", "lifu_stack_scale",
            "lifu_stack_pointers", "lifu_stack_get", "lifu_stack_aggregate"]


def _check(rc):
    if rc != 0:
        msg = load().lifu_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise LifuError(f"liblifusim error {rc}: {msg}")


def _ptr(a):
    """numpy array -> host pointer; int -> raw (device) pointer; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)


def make_time(n, d, c_ref=1500.0, cfl=0.5):
    nt, dt = C.c_int32(), C.c_double()
    _check(load().lifu_make_time((C.c_int32 * 3)(*[int(v) for v in n]), (C.c_double * 3)(*[float(v) for v in d]),
                                 c_ref, cfl, C.byref(nt), C.byref(dt)))
    return nt.value, dt.value


def pml_auto(n):
    out = (C.c_int32 * 3)()
    _check(load().lifu_pml_auto((C.c_int32 * 3)(*[int(v) for v in n]), out))
    return tuple(out)


def device_count() -> int:
    """B200-class devices the library can run on; 0 without a driver / device (never raises once the library is built)."""
    n = C.c_int32()
    _check(load().lifu_device_count(C.byref(n)))
    return int(n.value)


def slab_unique_id() -> bytes:
    """ncclUniqueId for a slab-decomposed solve: rank 0 makes it, every rank passes it to LifuSim(slab=...)."""
    buf = (C.c_ubyte * 128)()
    _check(load().lifu_slab_unique_id(buf))
    return bytes(buf)


class LifuSim:
    """One solver handle = one (device, stream).  Thin, typed wrapper over the C ABI.
    ``slab=(rank, nranks, nccl_id, exchange)`` makes it one rank's share of a z-slab decomposed grid
    (collective creation; exchange: "auto" | "nccl" | "peer")."""

    def __init__(self, n, d, dt, nt, pml=(-1, -1, -1), device=0, stream=0, pml_alpha=0.0, c_ref=0.0, slab=None):
        self._h = C.c_void_p()
        g = lifu_grid()
        g.n[:] = [int(v) for v in n]
        g.pml[:] = [int(v) for v in pml]
        g.d[:] = [float(v) for v in d]
        g.dt, g.nt, g.pml_alpha, g.c_ref = float(dt), int(nt), float(pml_alpha), float(c_ref)
        self.n = tuple(int(v) for v in n)
        self._lib = load()
        self.layout = None
        if slab is None:
            _check(self._lib.lifu_create(C.byref(g), int(device), C.c_void_p(int(stream) or None), C.byref(self._h)))
        else:
            rank, nranks, nccl_id, exchange = (tuple(slab) + ("auto",))[:4]
            sd = lifu_slab_desc()
            sd.rank, sd.nranks = int(rank), int(nranks)
            sd.exchange = EXCHANGE_MODES[exchange] if isinstance(exchange, str) else int(exchange)
            sd.nccl_id[:] = list(bytes(nccl_id))
            _check(self._lib.lifu_create_slab(C.byref(g), int(device), C.c_void_p(int(stream) or None), C.byref(sd),
                                              C.byref(self._h)))
            lay = lifu_slab_layout()
            _check(self._lib.lifu_slab_layout_of(self._h, C.byref(lay)))
            self.layout = {k: getattr(lay, k) for k, _ in lay._fields_}
        self._keep = []
        self._stage = None             # page-locked result staging buffers (run()), allocated on first use
        self._stage3 = None            # ... and of run_packaged()

    def close(self):
        if self._h:
            self._lib.lifu_destroy(self._h)
            self._h = C.c_void_p()
        self._stage = None
        self._stage3 = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ medium
    def set_medium(self, c0, rho0, alpha_db=None, alpha_power=0.9, alpha_mode="binary", device_ptrs=False, plane0=None):
        """Scalars -> homogeneous; else (Nx,Ny,Nz) maps (any layout; converted to x-fastest float32)
        or, with ``device_ptrs=True``, raw device pointers to x-fastest float32 inner-grid maps.
        ``plane0``: the maps hold only the inner z planes [plane0, plane0 + maps.shape[2]) (slab handles:
        see ``self.layout['medium_z0'/'medium_nz']``)."""
        mode = ALPHA_MODES[alpha_mode] if isinstance(alpha_mode, str) else int(alpha_mode)
        if device_ptrs:
            _check(self._lib.lifu_set_medium(self._h, _ptr(c0), _ptr(rho0), _ptr(alpha_db), alpha_power, mode, 0))
            return
        homog = np.ndim(c0) == 0 and np.ndim(rho0) == 0 and (alpha_db is None or np.ndim(alpha_db) == 0)
        if homog:
            arrs = [np.array([v if v is not None else 0.0], dtype=np.float32) for v in (c0, rho0, alpha_db)]
            _check(self._lib.lifu_set_medium(self._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), alpha_power, mode, 1))
            return
        if plane0 is None and self.layout is None:
            # float64 maps of the whole grid (what SimSetup.setup_sim_scene produces): hand them over as they are,
            # the float32 rounding and the x-fastest re-layout happen on the device
            maps = [c0, rho0, alpha_db]
            if all(m is None or (isinstance(m, np.ndarray) and m.dtype == np.float64 and m.shape == self.n) for m in maps):
                ref = maps[0]
                dense = ref.flags.c_contiguous or ref.flags.f_contiguous
                same = all(m is None or m.strides == ref.strides for m in maps)
                if dense and same:
                    st = (C.c_int64 * 3)(*[v // 8 for v in ref.strides])
                    # host -> device from memory that is already resident is fine as a pageable copy (23 ms against 12 ms
                    # through page-locked staging for three 216^3 maps); the device -> host side into fresh arrays is
                    # what needs staging (run())
                    _check(self._lib.lifu_set_medium_f64(self._h, _ptr(maps[0]), _ptr(maps[1]), _ptr(maps[2]), st,
                                                         alpha_power, mode))
                    return
        shape = self.n
        if plane0 is not None:
            nzp = max(np.shape(v)[2] for v in (c0, rho0, alpha_db) if np.ndim(v) == 3)
            shape = (self.n[0], self.n[1], nzp)
        arrs = []
        for v in (c0, rho0, 0.0 if alpha_db is None else alpha_db):
            full = np.broadcast_to(np.asarray(v, dtype=np.float32), shape)
            arrs.append(np.ascontiguousarray(full.transpose(2, 1, 0)))   # x fastest
        if plane0 is None:
            _check(self._lib.lifu_set_medium(self._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), alpha_power, mode, 0))
        else:
            _check(self._lib.lifu_set_medium_planes(self._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), alpha_power,
                                                    mode, 0, int(plane0), int(shape[2])))

    def set_medium_labels(self, labels, c0, rho0, alpha_db=None, alpha_power=0.9, alpha_mode="binary"):
        """Medium from a label volume (integer array of shape (Nx, Ny, Nz), any dense layout, values < 32) and per-label
        tables: the maps are expanded on the device (``lifu_set_medium_labels``)."""
        mode = ALPHA_MODES[alpha_mode] if isinstance(alpha_mode, str) else int(alpha_mode)
        lab = np.asarray(labels)
        if lab.shape != self.n:
            raise ValueError(f"labels must have shape {self.n}")
        if lab.dtype != np.uint8:
            if lab.size and (lab.min() < 0 or lab.max() > 255):
                raise ValueError("labels must lie in [0, 255]")
            lab = lab.astype(np.uint8)
        if not (lab.flags.c_contiguous or lab.flags.f_contiguous):
            lab = np.ascontiguousarray(lab)
        t = [np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for v in (c0, rho0)]
        ta = None if alpha_db is None else np.ascontiguousarray(alpha_db, dtype=np.float64).reshape(-1)
        st = (C.c_int64 * 3)(*[int(v) for v in lab.strides])
        _check(self._lib.lifu_set_medium_labels(self._h, _ptr(lab), st, t[0].size, _ptr(t[0]), _ptr(t[1]), _ptr(ta),
                                                alpha_power, mode))

    # ------------------------------------------------------------------ source geometry
    def set_elements(self, pos_m, size_m, angle_deg, bli_tolerance=0.05, upsampling_rate=5):
        pos = np.ascontiguousarray(pos_m, dtype=np.float64).reshape(-1, 3)
        size = np.ascontiguousarray(size_m, dtype=np.float64).reshape(-1, 2)
        ang = np.ascontiguousarray(angle_deg, dtype=np.float64).reshape(-1, 3)
        n_src = C.c_int64()
        _check(self._lib.lifu_set_elements(self._h, pos.shape[0], _ptr(pos), _ptr(size), _ptr(ang),
                                           float(bli_tolerance), int(upsampling_rate), C.byref(n_src)))
        return n_src.value

    def set_source_geometry(self, idx, row_ptr, col, w, n_el):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float32)
        _check(self._lib.lifu_set_source_geometry(self._h, _ptr(idx), _ptr(row_ptr), _ptr(col), _ptr(w),
                                                  idx.size, col.size, int(n_el)))

    def get_source_geometry(self):
        n_src, nnz, n_el = C.c_int64(), C.c_int64(), C.c_int32()
        _check(self._lib.lifu_get_source_sizes(self._h, C.byref(n_src), C.byref(nnz), C.byref(n_el)))
        idx = np.empty(n_src.value, dtype=np.int64)
        row_ptr = np.empty(n_src.value + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int32)
        w = np.empty(nnz.value, dtype=np.float32)
        _check(self._lib.lifu_get_source_geometry(self._h, _ptr(idx), _ptr(row_ptr), _ptr(col), _ptr(w)))
        return idx, row_ptr, col, w, n_el.value

    # ------------------------------------------------------------------ drive + run
    def set_drive(self, base_signal, delay_samples, gains, source_mode="additive"):
        base = np.ascontiguousarray(base_signal, dtype=np.float32)
        dly = np.ascontiguousarray(delay_samples, dtype=np.int32)
        g = np.ascontiguousarray(gains, dtype=np.float32)
        mode = SOURCE_MODES[source_mode] if isinstance(source_mode, str) else int(source_mode)
        _check(self._lib.lifu_set_drive(self._h, _ptr(base), base.size, _ptr(dly), _ptr(g), dly.size, mode))

    def run(self, p_max=None, p_min=None):
        """Run the time loop.  ``p_max``/``p_min``: None -> new numpy arrays are returned; numpy
        float32 arrays of Nx*Ny*Nz; or raw device pointers (ints)."""
        nvox = int(np.prod(self.n)) if self.layout is None else self.n[0] * self.n[1] * self.layout["sensor_nz"]
        own = p_max is None
        st = lifu_stats()
        if own and nvox >= (1 << 20):
            # large results: DMA into page-locked staging buffers kept on the handle, then a threaded copy into the fresh
            # arrays that are handed out (a pageable device->host copy would fault the fresh pages in one driver thread)
            stage = self._pinned_stage(nvox)
            if stage is not None:
                import torch
                _check(self._lib.lifu_run(self._h, C.c_void_p(stage[0].data_ptr()), C.c_void_p(stage[1].data_ptr()), C.byref(st)))
                p_max = np.empty(nvox, dtype=np.float32)
                p_min = np.empty(nvox, dtype=np.float32)
                torch.from_numpy(p_max).copy_(stage[0])
                torch.from_numpy(p_min).copy_(stage[1])
                return p_max, p_min, st.as_dict()
        if own:
            p_max = np.empty(nvox, dtype=np.float32)
            p_min = np.empty(nvox, dtype=np.float32)
        _check(self._lib.lifu_run(self._h, _ptr(p_max), _ptr(p_min), C.byref(st)))
        return p_max, p_min, st.as_dict()

    def run_resident(self):
        """Run the time loop and leave p_max / p_min on the device (``FieldStack.put`` / ``run_packaged`` pick them up)."""
        st = lifu_stats()
        _check(self._lib.lifu_run(self._h, None, None, C.byref(st)))
        return st.as_dict()

    def _pinned_stage(self, nvox):
        """Two page-locked float32 buffers of nvox elements (torch is the allocator); None when torch / CUDA pinning is
        not available."""
        cur = self._stage
        if cur is not None and cur[0].numel() == nvox:
            return cur
        try:
            import torch
            self._stage = (torch.empty(nvox, dtype=torch.float32, pin_memory=True),
                           torch.empty(nvox, dtype=torch.float32, pin_memory=True))
        except Exception:  # noqa: BLE001
            self._stage = None
        return self._stage

    def set_two_z(self, two_z):
        """2 * density * sound_speed for the intensity: a float, or a float64 array with one value per inner-grid
        voxel, x fastest."""
        a = np.ascontiguousarray(two_z, dtype=np.float64).reshape(-1)
        _check(self._lib.lifu_set_two_z(self._h, _ptr(a), a.size))

    def run_packaged(self):
        """Time loop + the packaging of kwave_if.py:136-141 on the device: (p_max, -p_min, intensity, stats) as flat
        x-fastest host arrays (float32, float32, float64).  Large results come through page-locked staging buffers kept
        on the handle and a threaded copy into the fresh arrays that are handed out (as in ``run``)."""
        st = lifu_stats()
        _check(self._lib.lifu_run(self._h, None, None, C.byref(st)))
        nvox = int(np.prod(self.n))
        p_max = np.empty(nvox, dtype=np.float32)
        pnp = np.empty(nvox, dtype=np.float32)
        inten = np.empty(nvox, dtype=np.float64)
        stage = self._pinned_stage3(nvox) if nvox >= (1 << 20) else None
        if stage is not None:
            import torch
            _check(self._lib.lifu_get_packaged(self._h, C.c_void_p(stage[0].data_ptr()), C.c_void_p(stage[1].data_ptr()),
                                               C.c_void_p(stage[2].data_ptr())))
            torch.from_numpy(p_max).copy_(stage[0])
            torch.from_numpy(pnp).copy_(stage[1])
            torch.from_numpy(inten).copy_(stage[2])
        else:
            _check(self._lib.lifu_get_packaged(self._h, _ptr(p_max), _ptr(pnp), _ptr(inten)))
        return p_max, pnp, inten, st.as_dict()

    def _pinned_stage3(self, nvox):
        cur = getattr(self, "_stage3", None)
        if cur is not None and cur[0].numel() == nvox:
            return cur
        try:
            import torch
            self._stage3 = (torch.empty(nvox, dtype=torch.float32, pin_memory=True),
                            torch.empty(nvox, dtype=torch.float32, pin_memory=True),
                            torch.empty(nvox, dtype=torch.float64, pin_memory=True))
        except Exception:  # noqa: BLE001
            self._stage3 = None
        return self._stage3

    def get_field(self, which):
        st = lifu_stats()
        _check(self._lib.lifu_get_info(self._h, C.byref(st)))
        N = list(st.n_exp)
        if self.layout is not None:
            N[2] = self.layout["nz"]
        out = np.empty(int(np.prod(N)), dtype=np.float32)
        _check(self._lib.lifu_get_field(self._h, int(which), _ptr(out)))
        return out.reshape(N[2], N[1], N[0]).transpose(2, 1, 0)

    def profile_stages(self, reps=5, with_source=True, max_stages=32):
        """Mean CUDA-event time per stage of one time step: list of (name, ms, bytes_per_voxel).
        with_source: False / 0 no source, True / 1 source-active step, 2 step of the steady source window."""
        stride = 32
        names = C.create_string_buffer(max_stages * stride)
        ms = (C.c_double * max_stages)()
        bpv = (C.c_double * max_stages)()
        n = C.c_int()
        _check(self._lib.lifu_profile_stages(self._h, int(reps), int(with_source), max_stages, names, stride,
                                             ms, bpv, C.byref(n)))
        raw = names.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(), ms[i], bpv[i]) for i in range(n.value)]

    def info(self):
        st = lifu_stats()
        _check(self._lib.lifu_get_info(self._h, C.byref(st)))
        return st.as_dict()


class BeamAnalysis:
    """Device-side beam analysis handle (``lifu_analysis_*``): holds the fields of every focus of a solution
    and evaluates the O(V) passes of ``Solution.analyze`` on the GPU.  No CPU fallback."""

    def __init__(self, axes, n_foci, z_ok=None, device=0, stream=0):
        self._h = C.c_void_p()
        self._lib = load()
        self.axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        self.n = tuple(int(a.size) for a in self.axes)
        self.n_foci = int(n_foci)
        zk = None if z_ok is None else np.ascontiguousarray(z_ok, dtype=np.uint8)
        _check(self._lib.lifu_analysis_create(int(device), C.c_void_p(int(stream) or None), (C.c_int32 * 3)(*self.n),
                                              self.n_foci, _ptr(self.axes[0]), _ptr(self.axes[1]), _ptr(self.axes[2]),
                                              _ptr(zk), C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.lifu_analysis_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_focus(self, focus, pnp, ipa, strides=None):
        """Stage one focus' fields.  numpy arrays of shape (Nx,Ny,Nz) in any dense layout (float32 pressure,
        float64 intensity; both the same layout), or raw device pointers with explicit element ``strides``."""
        if strides is None:
            pnp = np.asarray(pnp)
            if pnp.dtype != np.float32 or pnp.shape != self.n:
                raise ValueError(f"pnp must be a float32 array of shape {self.n}")
            ipa = np.asarray(ipa)
            if ipa.dtype != np.float64 or ipa.shape != self.n:
                raise ValueError(f"ipa must be a float64 array of shape {self.n}")
            if not (pnp.flags.c_contiguous or pnp.flags.f_contiguous):
                pnp = np.ascontiguousarray(pnp)
            strides = tuple(s // pnp.itemsize for s in pnp.strides)
            if tuple(s // ipa.itemsize for s in ipa.strides) != strides:
                ipa = np.asarray(ipa, order="C" if pnp.flags.c_contiguous else "F")
                if tuple(s // ipa.itemsize for s in ipa.strides) != strides:
                    ipa = np.ascontiguousarray(ipa)
                    pnp = np.ascontiguousarray(pnp)
                    strides = tuple(s // pnp.itemsize for s in pnp.strides)
        _check(self._lib.lifu_analysis_set_focus(self._h, int(focus), _ptr(pnp), _ptr(ipa),
                                                 (C.c_int64 * 3)(*[int(s) for s in strides])))

    def run_focus(self, focus, w, aspect, mainlobe_radius, sidelobe_radius, pnp_scale, line_pts=None,
                  centroid_factor=10 ** (-3 / 20)):
        """Metrics of one focus.  ``w``: inverse focus matrix (4x4 or 3x4); ``line_pts``: list of three (n,3)
        arrays of grid coordinates (or None).  Returns (metrics dict, [line value arrays])."""
        q = lifu_focus_query()
        w = np.asarray(w, dtype=np.float64)
        for i in range(3):
            for j in range(4):
                q.w[i][j] = float(w[i, j])
            q.aspect[i] = float(aspect[i])
        q.mainlobe_radius, q.sidelobe_radius = float(mainlobe_radius), float(sidelobe_radius)
        q.centroid_factor = float(centroid_factor)
        q.pnp_scale = float(pnp_scale)
        pts = vals = None
        sizes = [0, 0, 0]
        if line_pts is not None:
            sizes = [int(np.shape(p)[0]) for p in line_pts]
            pts = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) for p in line_pts]))
            vals = np.empty(pts.shape[0], dtype=np.float64)
        q.n_line[:] = sizes
        m = lifu_focus_metrics()
        _check(self._lib.lifu_analysis_run_focus(self._h, int(focus), C.byref(q), _ptr(pts), C.byref(m), _ptr(vals)))
        lines = []
        if vals is not None:
            o = 0
            for k in sizes:
                lines.append(vals[o:o + k])
                o += k
        return m.as_dict(), lines


class FieldStack:
    """The fields of every focus of one plan, resident on the device (``lifu_stack_*``): device-side packaging,
    ``Solution.scale``, aggregation over foci and the hand-over to ``BeamAnalysis`` without a host round trip.
    Host arrays come out as (n_foci, Nx, Ny, Nz) views of x-fastest storage -- the layout ``run_simulation`` returns."""

    def __init__(self, n, n_foci, device=0, stream=0):
        self._h = C.c_void_p()
        self._lib = load()
        self.n = tuple(int(v) for v in n)
        self.n_foci = int(n_foci)
        self.device = int(device)
        _check(self._lib.lifu_stack_create(self.device, C.c_void_p(int(stream) or None), (C.c_int32 * 3)(*self.n),
                                           self.n_foci, C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.lifu_stack_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def put(self, focus, sim: "LifuSim"):
        """Package the last run of ``sim`` (which needs ``set_two_z``) into slot ``focus``."""
        _check(self._lib.lifu_stack_put(self._h, int(focus), sim._h))

    def scale(self, focus, s, s2=None):
        """pressures *= s, intensity *= s2 (default ``s ** 2`` evaluated on ``s`` as given, e.g. a numpy float64)."""
        s2 = s ** 2 if s2 is None else s2
        _check(self._lib.lifu_stack_scale(self._h, int(focus), float(s), float(s2)))

    def pointers(self, focus):
        """(p_max, pnp, intensity) device pointers of one focus, and the element strides of (x, y, z)."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self._lib.lifu_stack_pointers(self._h, int(focus), C.byref(a), C.byref(b), C.byref(c)))
        return (a.value, b.value, c.value), (1, self.n[0], self.n[0] * self.n[1])

    def _host(self, count):
        """Fresh host arrays for `count` foci, pages touched by all host threads before the device -> host copy."""
        nv = int(np.prod(self.n))
        out = [np.empty(count * nv, dtype=np.float32), np.empty(count * nv, dtype=np.float32),
               np.empty(count * nv, dtype=np.float64)]
        try:
            import torch
            for a in out:
                torch.from_numpy(a).zero_()
        except ImportError:
            pass
        return out

    def _view(self, flat, count):
        nx, ny, nz = self.n
        a = flat.reshape(count, nz, ny, nx).transpose(0, 3, 2, 1)
        return a[0] if count == 1 else a

    def get(self, focus=None):
        """Host copies (p_max, pnp, intensity) of one focus -- (Nx, Ny, Nz) -- or of all of them -- (F, Nx, Ny, Nz)."""
        count = 1 if focus is not None else self.n_foci
        pm, pn, it = self._host(count)
        _check(self._lib.lifu_stack_get(self._h, -1 if focus is None else int(focus), _ptr(pm), _ptr(pn), _ptr(it)))
        if focus is None:
            nx, ny, nz = self.n
            return tuple(a.reshape(count, nz, ny, nx).transpose(0, 3, 2, 1) for a in (pm, pn, it))
        return tuple(self._view(a, 1) for a in (pm, pn, it))

    def aggregate(self):
        """(max over foci of p_max, max over foci of pnp, mean over foci of the intensity) as (Nx, Ny, Nz) host arrays."""
        pm, pn, it = self._host(1)
        _check(self._lib.lifu_stack_aggregate(self._h, _ptr(pm), _ptr(pn), _ptr(it)))
        return tuple(self._view(a, 1) for a in (pm, pn, it))
