"""Oracle: off-grid rectangular sources by band-limited interpolation (TEST INFRASTRUCTURE).

Restates ``kWaveArray.add_rect_element / get_array_binary_mask /
get_distributed_source_signal`` of k-wave-python 0.4.0 as driven by
/root/reference/src/openlifu/sim/kwave_if.py:29-47 (get_karray) and :71-78 (get_source).
Algorithm: Wise, Cox, Jaros, Treeby, JASA 146(1) 2019 (ledger A5, A10 in SURVEY.md 8c).
All indices are 0-based; linear indices are Fortran order (x fastest) on the *inner*
(un-expanded) grid, exactly like ``matlab_find`` on ``source.p_mask``.
"""
from __future__ import annotations

import numpy as np

from .kgrid import x_vec


def rotation_matrix_deg(theta_xyz) -> np.ndarray:
    """make_cart_rect rotation: R = Rz(theta[2]) @ Ry(theta[1]) @ Rx(theta[0]), degrees (A10).
    The reference passes theta = (el, az, roll) (xdc/element.py:216-226, kwave_if.py:42-43)."""
    tx, ty, tz = (np.deg2rad(float(t)) for t in theta_xyz)
    Rx = np.array([[1, 0, 0], [0, np.cos(tx), -np.sin(tx)], [0, np.sin(tx), np.cos(tx)]])
    Ry = np.array([[np.cos(ty), 0, np.sin(ty)], [0, 1, 0], [-np.sin(ty), 0, np.cos(ty)]])
    Rz = np.array([[np.cos(tz), -np.sin(tz), 0], [np.sin(tz), np.cos(tz), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def rect_point_counts(Lx: float, Ly: float, dx: float, upsampling_rate: float):
    """m_grid = area/dx^2, m_int = ceil(m_grid*upsampling), re-gridded to
    npts_x = round(sqrt(m_int*Lx/Ly)), npts_y = round(m_int/npts_x) (A5)."""
    m_grid = (Lx * Ly) / dx ** 2
    m_int = int(np.ceil(m_grid * upsampling_rate))
    npts_x = int(np.round(np.sqrt(m_int * Lx / Ly)))
    npts_y = int(np.round(m_int / npts_x))
    return m_grid, npts_x, npts_y


def make_cart_rect(pos, Lx, Ly, theta_deg, npts_x, npts_y) -> np.ndarray:
    """Cell-centred lattice on the rotated rectangle, returned as (3, npts_x*npts_y)."""
    d_x = 2.0 / npts_x
    d_y = 2.0 / npts_y
    p_x = np.linspace(-1 + d_x / 2, 1 - d_x / 2, npts_x)
    p_y = np.linspace(-1 + d_y / 2, 1 - d_y / 2, npts_y)
    P_x, P_y = np.meshgrid(p_x, p_y, indexing="ij")
    p0 = np.stack((P_x.flatten(), P_y.flatten(), np.zeros(P_x.size)), axis=0)
    A = rotation_matrix_deg(theta_deg) @ (np.diag([Lx, Ly, 1.0]) / 2.0)
    return A @ p0 + np.asarray(pos, dtype=np.float64).reshape(3, 1)


def element_integration_points(pos_m, size_m, angles_deg, translation_m, dx, upsampling_rate):
    """Integration points and per-point scale for one rect element (A5).
    ``translation_m`` is the array offset -mean(coords) of kwave_if.py:108 (rotation 0)."""
    Lx, Ly = float(size_m[0]), float(size_m[1])
    m_grid, nx, ny = rect_point_counts(Lx, Ly, dx, upsampling_rate)
    centre = np.asarray(pos_m, dtype=np.float64) + np.asarray(translation_m, dtype=np.float64)
    pts = make_cart_rect(centre, Lx, Ly, angles_deg, nx, ny)
    return pts, m_grid / pts.shape[1]


_STAR_CACHE: dict = {}


def star_offsets(tolerance: float):
    """tolStar stencil: offsets in [-h,h]^3, h = ceil(1/(pi*tol)), kept where |i*j*k| <= h (A5)."""
    key = float(tolerance)
    if key not in _STAR_CACHE:
        h = int(np.ceil(1.0 / (np.pi * tolerance)))
        lin = np.arange(-h, h + 1)
        i0, j0, k0 = np.meshgrid(lin, lin, lin, indexing="ij")
        keep = np.abs(i0 * j0 * k0) <= h
        # Fortran-order flattening, like matlab_mask(is0, matlab_find(instar))
        _STAR_CACHE[key] = (i0.flatten("F")[keep.flatten("F")],
                            j0.flatten("F")[keep.flatten("F")],
                            k0.flatten("F")[keep.flatten("F")], h)
    return _STAR_CACHE[key]


def tol_star(tolerance, vecs, N, dx, point):
    """Grid subscripts (0-based) of the truncated-sinc support of one integration point:
    nearest grid node by argmin |x_vec - p| (first on ties); an axis where the point is
    within dx*1e-3 of a node collapses to that node only; out-of-bounds nodes are dropped."""
    is_, js, ks, _ = star_offsets(tolerance)
    thr = dx * 1e-3
    closest = []
    for ax in range(3):
        ci = int(np.argmin(np.abs(vecs[ax] - point[ax])))
        closest.append(ci)
    sel = np.ones(is_.shape, dtype=bool)
    for ax, sub in enumerate((is_, js, ks)):
        if abs(vecs[ax][closest[ax]] - point[ax]) < thr:
            sel &= sub == 0
    i = is_[sel] + closest[0]
    j = js[sel] + closest[1]
    k = ks[sel] + closest[2]
    inb = (i >= 0) & (i < N[0]) & (j >= 0) & (j < N[1]) & (k >= 0) & (k < N[2])
    return i[inb], j[inb], k[inb]


def _sinc(x):
    out = np.ones_like(x)
    nz = x != 0
    out[nz] = np.sin(x[nz]) / x[nz]
    return out


def element_grid_weights(N, d, points, scale, tolerance, mask_only=False, single_precision=True):
    """offGridPoints for the integration points of one element on the inner grid."""
    N = tuple(int(v) for v in N)
    vecs = [x_vec(N[a], d[a]) for a in range(3)]
    if mask_only:
        grid = np.zeros(N, dtype=bool)
    else:
        grid = np.zeros(N, dtype=np.float32 if single_precision else np.float64)
    for p in range(points.shape[1]):
        pt = points[:, p]
        i, j, k = tol_star(tolerance, vecs, N, d[0], pt)
        if mask_only:
            grid[i, j, k] = True
        else:
            w = (_sinc(np.pi / d[0] * (vecs[0][i] - pt[0]))
                 * _sinc(np.pi / d[1] * (vecs[1][j] - pt[1]))
                 * _sinc(np.pi / d[2] * (vecs[2][k] - pt[2])))
            grid[i, j, k] += scale * w
    return grid


def array_source_geometry(N, d, elems_pos_m, elems_size_m, elems_angles_deg, translation_m,
                          bli_tolerance=0.05, upsampling_rate=5, single_precision=True):
    """Binary mask (get_array_binary_mask) and per-element weights restricted to it.

    Same arithmetic as calling ``element_grid_weights`` twice per element (mask, then
    weights) like kwave_if.py:75-77 does, but each element is accumulated inside its own
    bounding box so that 256^3-class grids stay cheap.

    Returns
    -------
    idx : (n_src,) int64  sorted Fortran-order linear indices of the mask on the inner grid
    W   : (n_src, n_el) float32 dense weight matrix (zero where an element does not reach)
    """
    N = tuple(int(v) for v in N)
    n_el = len(elems_pos_m)
    vecs = [x_vec(N[a], d[a]) for a in range(3)]
    wdt = np.float32 if single_precision else np.float64
    per_el = []
    for e in range(n_el):
        pts, scale = element_integration_points(elems_pos_m[e], elems_size_m[e], elems_angles_deg[e],
                                                translation_m, d[0], upsampling_rate)
        stars = [tol_star(bli_tolerance, vecs, N, d[0], pts[:, p]) for p in range(pts.shape[1])]
        nonempty = [s for s in stars if s[0].size]
        if not nonempty:
            per_el.append((np.zeros(0, np.int64), np.zeros(0, wdt)))
            continue
        lo = [min(int(s[a].min()) for s in nonempty) for a in range(3)]
        hi = [max(int(s[a].max()) for s in nonempty) for a in range(3)]
        shape = tuple(hi[a] - lo[a] + 1 for a in range(3))
        box_w = np.zeros(shape, dtype=wdt)
        box_m = np.zeros(shape, dtype=bool)
        for p, (i, j, k) in enumerate(stars):
            if not i.size:
                continue
            pt = pts[:, p]
            w = (_sinc(np.pi / d[0] * (vecs[0][i] - pt[0]))
                 * _sinc(np.pi / d[1] * (vecs[1][j] - pt[1]))
                 * _sinc(np.pi / d[2] * (vecs[2][k] - pt[2])))
            box_w[i - lo[0], j - lo[1], k - lo[2]] += scale * w
            box_m[i - lo[0], j - lo[1], k - lo[2]] = True
        bi, bj, bk = np.nonzero(box_m)
        lin = (bi + lo[0]) + N[0] * ((bj + lo[1]) + N[1] * (bk + lo[2]).astype(np.int64))
        per_el.append((lin.astype(np.int64), box_w[bi, bj, bk]))
    idx = np.unique(np.concatenate([pe[0] for pe in per_el])) if per_el else np.zeros(0, np.int64)
    W = np.zeros((idx.size, n_el), dtype=wdt)
    for e, (lin, w) in enumerate(per_el):
        W[np.searchsorted(idx, lin), e] = w
    return idx, W


def distributed_source_signal(W: np.ndarray, source_mat: np.ndarray, single_precision=True) -> np.ndarray:
    """get_distributed_source_signal: p[i,:] = sum_e W[i,e]*source_mat[e,:], accumulated
    element by element into a single-precision array like the reference dependency does."""
    n_src, n_el = W.shape
    out = np.zeros((n_src, source_mat.shape[1]), dtype=np.float32 if single_precision else np.float64)
    for e in range(n_el):
        rows = np.flatnonzero(W[:, e] != 0)
        if rows.size:
            out[rows] += W[rows, e][:, None].astype(np.float64) * source_mat[e][None, :]
    return out
