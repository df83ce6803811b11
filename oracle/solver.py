"""Oracle: the kspaceFirstOrder3D time loop (TEST INFRASTRUCTURE, parity unpinned).

Restates what ``kspaceFirstOrder3D(...)`` (called at
/root/reference/src/openlifu/sim/kwave_if.py:124-129 with the options of :117-123) and
the ``kspaceFirstOrder-OMP`` binary compute: first-order k-space pseudospectral
acoustics with a split-field PML (Treeby & Cox 2010; Treeby et al. 2012).  The
dependency (k-wave-python 0.4.0 + binary) is not in the container, so each
behavioural assumption is a switch in :class:`Assumptions`.

Arrays are numpy C-ordered (Nx,Ny,Nz); outputs are flattened in Fortran order
(x fastest) like the dependency's ``p_max`` / ``p_min`` vectors (kwave_if.py:132-141).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.fft as sfft

from . import kgrid as kg


@dataclass
class Assumptions:
    pml_range: tuple = (10, 40)          # A3
    pml_alpha: float = 2.0               # A4
    source_kspace_correction: bool = True  # A6: additive source filtered with cos(c_ref k dt/2)
    absorb_eta: bool = True              # A7: binary ignores alpha_mode='no_dispersion' -> eta term kept
    absorb_tau: bool = True
    alpha_power: float = 0.9             # kwave_if.py:55,61
    staggered_density: bool = True       # A9
    record_start: int = 0                # A8: p_max/p_min sampled from the first step
    pml_size: tuple | None = None        # explicit PML thickness per axis (known-answer tests; None = pml_auto, A3)


@dataclass
class SolverInputs:
    """Everything the time loop needs, already on the *inner* grid (no PML)."""
    N: tuple                 # inner grid size (Nx,Ny,Nz)
    d: tuple                 # spacing in m
    dt: float
    Nt: int
    c0: object               # scalar or (Nx,Ny,Nz)
    rho0: object
    alpha_db: object         # dB/(MHz^y cm); scalar or map
    src_idx: np.ndarray      # Fortran-order linear indices on the inner grid, sorted
    src_p: np.ndarray        # (n_src, L) pressure source signals *before* k-Wave source scaling


def _cast(a, dtype):
    return np.asarray(a).astype(dtype)


def simulate(inp: SolverInputs, dtype=np.float32, asm: Assumptions | None = None, workers: int = -1,
             return_p_final: bool = False, max_steps: int | None = None, progress=None, backend: str = "numpy"):
    """Run the time loop; returns dict(p_max, p_min [Fortran-flat over inner grid], pml, N_exp, ...).

    ``backend="torch"`` runs the SAME step with torch CPU tensors (MKL FFTs and element-wise operations on all host
    threads) -- the k-Wave-OMP-like timing arm of bench.py; ``"numpy"`` (scipy.fft with ``workers`` threads, single-
    threaded element-wise operations) is the checker the parity tests use."""
    asm = asm or Assumptions()
    rdt = np.dtype(dtype)
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    N_in = tuple(int(v) for v in inp.N)
    d = tuple(float(v) for v in inp.d)
    dt = float(inp.dt)
    pml = tuple(int(v) for v in asm.pml_size) if asm.pml_size is not None else kg.optimal_pml_size(N_in, asm.pml_range)
    N = tuple(N_in[a] + 2 * pml[a] for a in range(3))

    homogeneous = np.ndim(inp.c0) == 0 and np.ndim(inp.rho0) == 0 and np.ndim(inp.alpha_db) == 0
    c0 = kg.expand_edge(np.asarray(inp.c0, dtype=np.float64), pml)
    rho0 = kg.expand_edge(np.asarray(inp.rho0, dtype=np.float64), pml)
    alpha_db = kg.expand_edge(np.asarray(inp.alpha_db, dtype=np.float64), pml)
    c_ref = float(np.max(c0))
    absorbing = bool(np.any(alpha_db != 0))

    ops = kg.KOps(N, d, dt, c_ref, asm.alpha_power if absorbing else None)
    kappa = _cast(ops.kappa, rdt)
    gp = [_cast(ops.ddx_pos * ops.kappa, cdt), _cast(ops.ddy_pos * ops.kappa, cdt), _cast(ops.ddz_pos * ops.kappa, cdt)]
    gn = [_cast(ops.ddx_neg * ops.kappa, cdt), _cast(ops.ddy_neg * ops.kappa, cdt), _cast(ops.ddz_neg * ops.kappa, cdt)]
    del kappa
    src_kappa = _cast(ops.source_kappa, rdt)

    shp = [(N[0], 1, 1), (1, N[1], 1), (1, 1, N[2])]
    pml_c = [_cast(kg.pml_profile(N[a], d[a], dt, c_ref, pml[a], asm.pml_alpha, False), rdt).reshape(shp[a]) for a in range(3)]
    pml_sg = [_cast(kg.pml_profile(N[a], d[a], dt, c_ref, pml[a], asm.pml_alpha, True), rdt).reshape(shp[a]) for a in range(3)]

    if rho0.ndim == 0:
        dt_rho0_sg = [rdt.type(dt / float(rho0))] * 3
    else:
        dt_rho0_sg = [_cast(dt / (kg.staggered_density(rho0, a) if asm.staggered_density else rho0), rdt) for a in range(3)]
    dt_rho0 = _cast(dt * rho0, rdt)
    rho0_r = _cast(rho0, rdt)
    c2 = _cast(c0 ** 2, rdt)

    if absorbing:
        y = asm.alpha_power
        a_np = kg.db2neper(alpha_db, y)
        tau = _cast(-2.0 * a_np * c0 ** (y - 1.0), rdt) if asm.absorb_tau else None
        eta = _cast(2.0 * a_np * c0 ** y * np.tan(np.pi * y / 2.0), rdt) if asm.absorb_eta else None
        nabla1 = _cast(ops.nabla1, rdt)
        nabla2 = _cast(ops.nabla2, rdt)

    # ---- source: inner-grid F-order linear index -> expanded-grid subscripts, k-Wave scaling (A6)
    src_idx = np.asarray(inp.src_idx, dtype=np.int64)
    si = src_idx % N_in[0] + pml[0]
    sj = (src_idx // N_in[0]) % N_in[1] + pml[1]
    sk = src_idx // (N_in[0] * N_in[1]) + pml[2]
    c0_src = c0[si, sj, sk] if c0.ndim else float(c0)
    scale = 2.0 * dt / (3.0 * c0_src * d[0])
    src_p = _cast(np.asarray(inp.src_p) * (scale[:, None] if np.ndim(scale) else scale), rdt)
    L = src_p.shape[1]

    axes = (0, 1, 2)
    Nt = int(inp.Nt) if max_steps is None else min(int(inp.Nt), int(max_steps))
    if backend == "torch":
        return _loop_torch(locals(), return_p_final, progress)
    if backend != "numpy":
        raise ValueError(f"unknown backend {backend!r}")

    def fwd(a):
        return sfft.rfftn(a, axes=axes, workers=workers)

    def inv(A):
        return sfft.irfftn(A, s=N, axes=axes, workers=workers)

    p = np.zeros(N, dtype=rdt)
    u = [np.zeros(N, dtype=rdt) for _ in range(3)]
    rho = [np.zeros(N, dtype=rdt) for _ in range(3)]
    inner = tuple(slice(pml[a], pml[a] + N_in[a]) for a in range(3))
    p_max = np.full(N_in, -np.inf, dtype=rdt)
    p_min = np.full(N_in, np.inf, dtype=rdt)

    import time as _time
    t_loop0 = _time.perf_counter()
    for t in range(Nt):
        # (1) pressure gradient -> particle velocity on the staggered grid
        P = fwd(p)
        for a in range(3):
            g = inv(gp[a] * P)
            u[a] = pml_sg[a] * (pml_sg[a] * u[a] - dt_rho0_sg[a] * g)
        # (2) velocity divergence terms
        du = [inv(gn[a] * fwd(u[a])) for a in range(3)]
        # (3) split density update
        for a in range(3):
            rho[a] = pml_c[a] * (pml_c[a] * rho[a] - dt_rho0 * du[a])
        # (4) additive pressure source, same field into each split component
        if t < L:
            S = np.zeros(N, dtype=rdt)
            S[si, sj, sk] = src_p[:, t]
            if asm.source_kspace_correction:
                S = inv(src_kappa * fwd(S)).astype(rdt, copy=False)
            for a in range(3):
                rho[a] = rho[a] + S
        # (5) equation of state
        rsum = rho[0] + rho[1] + rho[2]
        if absorbing:
            acc = rsum
            if tau is not None:
                acc = acc + tau * inv(nabla1 * fwd(rho0_r * (du[0] + du[1] + du[2])))
            if eta is not None:
                acc = acc - eta * inv(nabla2 * fwd(rsum))
            p = (c2 * acc).astype(rdt, copy=False)
        else:
            p = (c2 * rsum).astype(rdt, copy=False)
        # (6) sensor: running max / min over the inner grid
        if t >= asm.record_start:
            pi = p[inner]
            np.maximum(p_max, pi, out=p_max)
            np.minimum(p_min, pi, out=p_min)
        if progress is not None:
            progress(t, p)

    loop_s = _time.perf_counter() - t_loop0
    out = {
        "loop_s": loop_s,
        "p_max": p_max.flatten("F"),
        "p_min": p_min.flatten("F"),
        "pml": pml, "N_exp": N, "c_ref": c_ref, "Nt": Nt, "L": L,
        "homogeneous": homogeneous, "absorbing": absorbing,
    }
    if return_p_final:
        out["p_final"] = p
        out["u_final"] = u
    return out


def _loop_torch(v, return_p_final, progress):
    """The time loop of :func:`simulate` on torch CPU tensors (same operators, same order of operations)."""
    import time as _time

    import torch
    if v["workers"] and v["workers"] > 0:
        torch.set_num_threads(int(v["workers"]))
    N, N_in, pml, rdt, asm = v["N"], v["N_in"], v["pml"], v["rdt"], v["asm"]
    tdt = torch.float32 if rdt == np.float32 else torch.float64

    def T(a):
        if a is None:
            return None
        if isinstance(a, (list, tuple)):
            return [T(x) for x in a]
        if isinstance(a, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(a))
        return a                                                    # numpy scalars broadcast as Python numbers
    gp, gn, pml_c, pml_sg = T(v["gp"]), T(v["gn"]), T(v["pml_c"]), T(v["pml_sg"])
    dt_rho0_sg = [T(x) if isinstance(x, np.ndarray) else float(x) for x in v["dt_rho0_sg"]]
    dt_rho0 = T(v["dt_rho0"]) if np.ndim(v["dt_rho0"]) else float(v["dt_rho0"])
    c2 = T(v["c2"]) if np.ndim(v["c2"]) else float(v["c2"])
    rho0_r = T(v["rho0_r"]) if np.ndim(v["rho0_r"]) else float(v["rho0_r"])
    src_kappa = T(v["src_kappa"])
    absorbing = v["absorbing"]
    if absorbing:
        tau = v["tau"] if v["tau"] is None else (T(v["tau"]) if np.ndim(v["tau"]) else float(v["tau"]))
        eta = v["eta"] if v["eta"] is None else (T(v["eta"]) if np.ndim(v["eta"]) else float(v["eta"]))
        nabla1, nabla2 = T(v["nabla1"]), T(v["nabla2"])
    si, sj, sk = (torch.from_numpy(np.asarray(a)) for a in (v["si"], v["sj"], v["sk"]))
    src_p = torch.from_numpy(np.ascontiguousarray(v["src_p"]))
    L, Nt = v["L"], v["Nt"]
    dims = (0, 1, 2)

    def fwd(a):
        return torch.fft.rfftn(a, dim=dims)

    def inv(A):
        return torch.fft.irfftn(A, s=N, dim=dims)

    p = torch.zeros(N, dtype=tdt)
    u = [torch.zeros(N, dtype=tdt) for _ in range(3)]
    rho = [torch.zeros(N, dtype=tdt) for _ in range(3)]
    inner = tuple(slice(pml[a], pml[a] + N_in[a]) for a in range(3))
    p_max = torch.full(N_in, -float("inf"), dtype=tdt)
    p_min = torch.full(N_in, float("inf"), dtype=tdt)
    t_loop0 = _time.perf_counter()
    for t in range(Nt):
        P = fwd(p)
        for a in range(3):
            g = inv(gp[a] * P)
            u[a] = pml_sg[a] * (pml_sg[a] * u[a] - dt_rho0_sg[a] * g)
        du = [inv(gn[a] * fwd(u[a])) for a in range(3)]
        for a in range(3):
            rho[a] = pml_c[a] * (pml_c[a] * rho[a] - dt_rho0 * du[a])
        if t < L:
            S = torch.zeros(N, dtype=tdt)
            S[si, sj, sk] = src_p[:, t]
            if asm.source_kspace_correction:
                S = inv(src_kappa * fwd(S))
            for a in range(3):
                rho[a] = rho[a] + S
        rsum = rho[0] + rho[1] + rho[2]
        if absorbing:
            acc = rsum
            if tau is not None:
                acc = acc + tau * inv(nabla1 * fwd(rho0_r * (du[0] + du[1] + du[2])))
            if eta is not None:
                acc = acc - eta * inv(nabla2 * fwd(rsum))
            p = c2 * acc
        else:
            p = c2 * rsum
        if t >= asm.record_start:
            pi = p[inner]
            torch.maximum(p_max, pi, out=p_max)
            torch.minimum(p_min, pi, out=p_min)
        if progress is not None:
            progress(t, p.numpy())
    loop_s = _time.perf_counter() - t_loop0
    out = {"loop_s": loop_s, "p_max": p_max.numpy().flatten("F"), "p_min": p_min.numpy().flatten("F"), "pml": pml,
           "N_exp": N, "c_ref": v["c_ref"], "Nt": Nt, "L": L, "homogeneous": v["homogeneous"], "absorbing": absorbing}
    if return_p_final:
        out["p_final"] = p.numpy()
        out["u_final"] = [x.numpy() for x in u]
    return out
