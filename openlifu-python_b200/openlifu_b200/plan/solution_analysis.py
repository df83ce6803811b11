"""Beam metrics of a simulated solution.

Mirrors the public surface of /root/reference/src/openlifu/plan/solution_analysis.py
(``SolutionAnalysis:50``, ``SolutionAnalysisOptions:228``, ``find_centroid:306``,
``get_focus_matrix:319``, ``get_gridded_transformed_coords:344``, ``get_offset_grid:365``,
``calc_dist_from_focus:384``, ``get_mask:405``, ``interp_transformed_axis:444``,
``get_beam_bounds:489``, ``get_beamwidth:537``) with the same names, arguments and results.

The implementation differs: the reference materialises an ``(N, 4)`` homogeneous coordinate
matrix per call and leans on ``DataArray.interp``; here the focus-frame coordinates are evaluated
separably from the three 1-D grid vectors (``FocusFrame.offsets``) and the line samples by a direct
trilinear gather (``trilinear_line``), so one focus costs a few broadcast passes over the grid.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Type

import numpy as np
import pandas as pd

from .. import xa
from ..util.dict_conversion import DictMixin
from ..util.units import getunitconversion, getunittype
from .param_constraint import PARAM_STATUS_SYMBOLS, ParameterConstraint

DEFAULT_ORIGIN = np.zeros(3)

# id -> [aggregation over foci, format, unit, label]   (the table layout of the reference, :18-48)
PARAM_FORMATS = {
    "mainlobe_pnp_MPa": ["max", "0.3f", "MPa", "Mainlobe Peak Negative Pressure"],
    "mainlobe_isppa_Wcm2": ["max", "0.1f", "W/cm^2", "Mainlobe I_SPPA"],
    "mainlobe_ispta_mWcm2": ["mean", "0.1f", "mW/cm^2", "Mainlobe I_SPTA"],
    "target_position_lat_mm": ["mean", "0.1f", "mm", "Target Position (Lateral)"],
    "target_position_ele_mm": ["mean", "0.1f", "mm", "Target Position (Elevation)"],
    "target_position_ax_mm": ["mean", "0.1f", "mm", "Target Position (Axial)"],
    "focal_centroid_lat_mm": ["mean", "0.1f", "mm", "Focal Centroid (Lateral)"],
    "focal_centroid_ele_mm": ["mean", "0.1f", "mm", "Focal Centroid (Elevation)"],
    "focal_centroid_ax_mm": ["mean", "0.1f", "mm", "Focal Centroid (Axial)"],
    "beamwidth_lat_3dB_mm": ["mean", "0.2f", "mm", "3dB Beamwidth (Lateral)"],
    "beamwidth_ele_3dB_mm": ["mean", "0.2f", "mm", "3dB Beamwidth (Elevational)"],
    "beamwidth_ax_3dB_mm": ["mean", "0.2f", "mm", "3dB Beamwidth (Axial)"],
    "beamwidth_lat_6dB_mm": ["mean", "0.2f", "mm", "6dB Beamwidth (Lateral)"],
    "beamwidth_ele_6dB_mm": ["mean", "0.2f", "mm", "6dB Beamwidth (Elevational)"],
    "beamwidth_ax_6dB_mm": ["mean", "0.2f", "mm", "6dB Beamwidth (Axial)"],
    "sidelobe_pnp_MPa": ["max", "0.3f", "MPa", "Sidelobe Peak Negative Pressure"],
    "sidelobe_isppa_Wcm2": ["max", "0.1f", "W/cm^2", "Sidelobe I_SPPA"],
    "sidelobe_to_mainlobe_pressure_ratio": ["mean", "0.2f", "", "Sidelobe/Mainlobe Pressure Ratio"],
    "sidelobe_to_mainlobe_intensity_ratio": ["mean", "0.2f", "", "Sidelobe/Mainlobe Intensity Ratio"],
    "global_pnp_MPa": ["max", "0.3f", "MPa", "Global Peak Negative Pressure"],
    "global_isppa_Wcm2": ["max", "0.1f", "W/cm^2", "Global I_SPPA"],
    "global_ispta_mWcm2": [None, "0.1f", "mW/cm^2", "Global I_SPTA"],
    "MI": [None, "0.2f", "", "MI"],
    "TIC": [None, "0.2f", "", "TIC"],
    "voltage_V": [None, "0.1f", "V", "Voltage"],
    "p0_MPa": ["max", "0.3f", "MPa", "Emitted Pressure"],
    "power_W": [None, "0.2f", "W", "Emitted Power"],
    "duty_cycle_pulse_train_pct": [None, "0.1f", "%", "Pulse Train Duty Cycle"],
    "duty_cycle_sequence_pct": [None, "0.1f", "%", "Sequence Duty Cycle"],
    "sequence_duration_s": [None, "0.0f", "s", "Sequence Duration"],
}


def _constraints_from(d: Dict[str, Any]) -> Dict[str, ParameterConstraint]:
    return {k: (v if isinstance(v, ParameterConstraint) else ParameterConstraint.from_dict(v))
            for k, v in (d or {}).items()}


@dataclass
class SolutionAnalysis(DictMixin):
    """Per-focus lists and scalar summaries; field names are the reference's (:52-143)."""
    mainlobe_pnp_MPa: List[float] = field(default_factory=list)
    mainlobe_isppa_Wcm2: List[float] = field(default_factory=list)
    mainlobe_ispta_mWcm2: List[float] = field(default_factory=list)
    target_position_lat_mm: List[float] = field(default_factory=list)
    target_position_ele_mm: List[float] = field(default_factory=list)
    target_position_ax_mm: List[float] = field(default_factory=list)
    focal_centroid_lat_mm: List[float] = field(default_factory=list)
    focal_centroid_ele_mm: List[float] = field(default_factory=list)
    focal_centroid_ax_mm: List[float] = field(default_factory=list)
    beamwidth_lat_3dB_mm: List[float] = field(default_factory=list)
    beamwidth_ele_3dB_mm: List[float] = field(default_factory=list)
    beamwidth_ax_3dB_mm: List[float] = field(default_factory=list)
    beamwidth_lat_6dB_mm: List[float] = field(default_factory=list)
    beamwidth_ele_6dB_mm: List[float] = field(default_factory=list)
    beamwidth_ax_6dB_mm: List[float] = field(default_factory=list)
    sidelobe_pnp_MPa: List[float] = field(default_factory=list)
    sidelobe_isppa_Wcm2: List[float] = field(default_factory=list)
    sidelobe_to_mainlobe_pressure_ratio: List[float] = field(default_factory=list)
    sidelobe_to_mainlobe_intensity_ratio: List[float] = field(default_factory=list)
    global_pnp_MPa: List[float] = field(default_factory=list)
    global_isppa_Wcm2: List[float] = field(default_factory=list)
    global_ispta_mWcm2: Optional[float] = None
    MI: Optional[float] = None
    TIC: Optional[float] = None
    voltage_V: Optional[float] = None
    p0_MPa: List[float] = field(default_factory=list)
    power_W: Optional[float] = None
    duty_cycle_pulse_train_pct: Optional[float] = None
    duty_cycle_sequence_pct: Optional[float] = None
    sequence_duration_s: Optional[float] = None
    param_constraints: Dict[str, ParameterConstraint] = field(default_factory=dict)

    def to_table(self, constraints: Dict[str, ParameterConstraint] | None = None, focus_index=None) -> pd.DataFrame:
        constraints = self.param_constraints if constraints is None else constraints
        for p in constraints:
            if p not in PARAM_FORMATS:
                raise ValueError(f"Unknown parameter constraint for '{p}'. Must be one of: {list(PARAM_FORMATS.keys())}")
        rows = []
        for param, (agg, fmt, unit, label) in PARAM_FORMATS.items():
            raw = getattr(self, param)
            if agg is None:
                by_focus, value = None, raw
            elif agg == "max":
                by_focus, value = raw, max(raw)
            elif agg == "mean":
                by_focus, value = raw, np.mean(raw)
            else:
                raise ValueError(f"Unknown aggregation method '{agg}' for parameter '{param}'.")
            if value is None:
                continue
            row = {"id": param, "Param": label, "Value": "", "Units": unit, "Status": "", "_value": value,
                   "_value_by_focus": by_focus, "_warning": False, "_error": False}
            if np.isnan(value):
                row["Value"] = "NaN"
            else:
                if focus_index is None:
                    row["Value"] = f"{value:{fmt}}"
                elif by_focus is None:
                    row["Value"] = "N/A"
                else:
                    row["Value"] = f"{by_focus[focus_index]:{fmt}}"
                if param in constraints:
                    c = constraints[param]
                    row["_warning"], row["_error"] = c.is_warning(value), c.is_error(value)
                    row["Status"] = PARAM_STATUS_SYMBOLS[c.get_status(value)]
            rows.append(row)
        return pd.DataFrame.from_records(rows)

    @classmethod
    def from_dict(cls: Type["SolutionAnalysis"], parameter_dict: Dict[str, Any]) -> "SolutionAnalysis":
        d = dict(parameter_dict)
        d["param_constraints"] = _constraints_from(d.get("param_constraints", {}))
        return cls(**d)

    @staticmethod
    def from_json(json_string: str) -> "SolutionAnalysis":
        return SolutionAnalysis.from_dict(json.loads(json_string))

    def to_json(self, compact: bool) -> str:
        if compact:
            return json.dumps(self.to_dict(), separators=(",", ":"))
        return json.dumps(self.to_dict(), indent=4)


def _positive(name, v, strict=True):
    if not isinstance(v, (int, float)) or (v <= 0 if strict else v < 0):
        raise ValueError(f"{name} must be a {'positive' if strict else 'non-negative'} number")


@dataclass
class SolutionAnalysisOptions(DictMixin):
    """Analysis knobs with the reference's names and defaults (:230-260)."""
    standoff_sound_speed: float = 1500.0
    standoff_density: float = 1000.0
    ref_sound_speed: float = 1500.0
    ref_density: float = 1000.0
    mainlobe_aspect_ratio: Tuple[float, float, float] = (1., 1., 5.)
    mainlobe_radius: float = 2.5e-3
    beamwidth_radius: float = 5e-3
    sidelobe_radius: float = 3e-3
    sidelobe_zmin: float = 1e-3
    distance_units: str = "m"
    param_constraints: Dict[str, ParameterConstraint] = field(default_factory=dict)

    def __post_init__(self):
        for label, v in (("Standoff sound speed", self.standoff_sound_speed), ("Standoff density", self.standoff_density),
                         ("Reference sound speed", self.ref_sound_speed), ("Reference density", self.ref_density)):
            if v <= 0:
                raise ValueError(f"{label} must be greater than 0")
        if not isinstance(self.mainlobe_aspect_ratio, (tuple, list)) or len(self.mainlobe_aspect_ratio) != 3:
            raise TypeError("Mainlobe aspect ratio must be a tuple or list of three floats (lat, ele, ax)")
        self.mainlobe_aspect_ratio = tuple(self.mainlobe_aspect_ratio)
        if not all(isinstance(x, (int, float)) for x in self.mainlobe_aspect_ratio):
            raise TypeError("Mainlobe aspect ratio must contain only numbers")
        _positive("Mainlobe radius", self.mainlobe_radius)
        _positive("Beamwidth radius", self.beamwidth_radius)
        _positive("Sidelobe radius", self.sidelobe_radius)
        _positive("Sidelobe minimum z", self.sidelobe_zmin, strict=False)
        if not isinstance(self.distance_units, str):
            raise TypeError("Distance units must be a string")
        if getunittype(self.distance_units) != "distance":
            raise ValueError(f"Distance units must be a length unit, got {self.distance_units}")

    @classmethod
    def from_dict(cls: Type["SolutionAnalysisOptions"], parameter_dict: Dict[str, Any]) -> "SolutionAnalysisOptions":
        d = dict(parameter_dict)
        d["param_constraints"] = _constraints_from(d.get("param_constraints", {}))
        return cls(**d)


# ------------------------------------------------------------------------------------------------
# numpy kernels
def _axes(da) -> List[np.ndarray]:
    return [np.asarray(da.coords[d].data if hasattr(da.coords[d], "data") else da.coords[d], dtype=np.float64)
            for d in da.dims]


def _values(da) -> np.ndarray:
    return np.asarray(da.data)


class FocusFrame:
    """Right-handed frame whose z axis runs from ``origin`` to ``focus`` and whose x axis stays in
    the x-z plane of the grid; ``matrix`` maps frame coordinates to grid coordinates."""

    def __init__(self, focus, origin=DEFAULT_ORIGIN):
        f = np.asarray(focus, dtype=np.float64)
        o = np.asarray(origin, dtype=np.float64)
        ez = (f - o) / np.linalg.norm(f - o)
        az = -np.arctan2(ez[0], ez[2])
        ex = np.array([np.cos(az), 0.0, np.sin(az)])
        ey = np.cross(ez, ex)
        m = np.zeros((4, 4))
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = ex, ey, ez, f
        m[3, 3] = 1.0
        self.matrix = m
        self.inverse = np.linalg.inv(m)

    def offsets(self, axes: Sequence[np.ndarray]) -> np.ndarray:
        """Frame coordinates of every grid node, shape ``(*grid, 3)``.  Affine maps are separable:
        component i = a_i[x] + b_i[y] + c_i[z] + t_i, evaluated by broadcasting the 1-D axes."""
        x, y, z = (np.asarray(a, dtype=np.float64) for a in axes)
        out = np.empty((x.size, y.size, z.size, 3))
        w = self.inverse
        for i in range(3):
            out[..., i] = ((w[i, 0] * x)[:, None, None] + (w[i, 1] * y)[None, :, None]) + (w[i, 2] * z)[None, None, :] + w[i, 3]
        return out

    def distance(self, axes: Sequence[np.ndarray], aspect_ratio=(1, 1, 1)) -> np.ndarray:
        x, y, z = (np.asarray(a, dtype=np.float64) for a in axes)
        w = self.inverse
        acc = np.zeros((x.size, y.size, z.size))
        for i in range(3):
            c = ((w[i, 0] * x)[:, None, None] + (w[i, 1] * y)[None, :, None]) + (w[i, 2] * z)[None, None, :] + w[i, 3]
            c /= aspect_ratio[i]
            acc += c * c
        return np.sqrt(acc, out=acc)

    def line(self, axis: int, offsets: np.ndarray) -> np.ndarray:
        """Grid coordinates of the points ``offsets`` along frame axis ``axis`` through the focus."""
        pts = np.zeros((4, offsets.size))
        pts[axis] = offsets
        pts[3] = 1.0
        return (self.matrix @ pts).T[:, :3]


def _bracket(axis: np.ndarray, q: np.ndarray):
    """Lower node index and fractional position of ``q`` on a monotonically increasing axis;
    points outside the axis are flagged (they sample NaN, as a linear interpolant without
    extrapolation does)."""
    n = axis.size
    if n == 1:
        inside = q == axis[0]
        return np.zeros(q.shape, dtype=np.intp), np.zeros(q.shape), inside
    i = np.clip(np.searchsorted(axis, q, side="right") - 1, 0, n - 2)
    t = (q - axis[i]) / (axis[i + 1] - axis[i])
    inside = (q >= axis[0]) & (q <= axis[-1])
    return i, t, inside


def trilinear_line(values: np.ndarray, axes: Sequence[np.ndarray], pts: np.ndarray) -> np.ndarray:
    """Trilinear samples of a 3-D array at the points ``pts`` (n, 3) in grid coordinates; NaN
    outside the grid."""
    (ix, tx, okx), (iy, ty, oky), (iz, tz, okz) = (_bracket(a, pts[:, k]) for k, a in enumerate(axes))
    v = values.astype(np.float64, copy=False)
    jx = np.minimum(ix + 1, v.shape[0] - 1)
    jy = np.minimum(iy + 1, v.shape[1] - 1)
    jz = np.minimum(iz + 1, v.shape[2] - 1)

    def lerp(a, b, t):
        return a + (b - a) * t

    # interpolate along x first, then y, then z (the order a dimension-by-dimension linear
    # interpolant over dims (x, y, z) uses)
    c00 = lerp(v[ix, iy, iz], v[jx, iy, iz], tx)
    c10 = lerp(v[ix, jy, iz], v[jx, jy, iz], tx)
    c01 = lerp(v[ix, iy, jz], v[jx, iy, jz], tx)
    c11 = lerp(v[ix, jy, jz], v[jx, jy, jz], tx)
    out = lerp(lerp(c00, c10, ty), lerp(c01, c11, ty), tz)
    out[~(okx & oky & okz)] = np.nan
    return out


# ------------------------------------------------------------------------------------------------
# reference-named entry points (xarray in, xarray / numpy out)
def find_centroid(da, cutoff: float, units=None) -> np.ndarray:
    """Value-weighted centroid of the region where ``da > cutoff``."""
    if units is not None and getunittype(units) != "distance":
        raise ValueError(f"Units must be a length unit, got {units}")
    v = _values(da).astype(np.float64)
    with np.errstate(invalid="ignore"):
        w = np.where(v > cutoff, v, 0.0)
    axes = _axes(da)
    total = w.sum()
    cen = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, a in enumerate(axes):
            shape = [1] * w.ndim
            shape[k] = a.size
            cen.append(np.sum(w * a.reshape(shape)) / total)
    cen = np.array(cen)
    if units is not None:
        cu = [getattr(da.coords[d], "attrs", {}).get("units", None) for d in da.dims]
        cen = np.array([getunitconversion(u, units) * c for u, c in zip(cu, cen)])
    return cen


def get_focus_matrix(focus, origin=(0, 0, 0)) -> np.ndarray:
    """4x4 transform from the focus frame to grid coordinates."""
    return FocusFrame(focus, origin).matrix


def _frame_from_matrix(matrix: np.ndarray) -> FocusFrame:
    fr = FocusFrame.__new__(FocusFrame)
    fr.matrix = np.asarray(matrix, dtype=np.float64)
    fr.inverse = np.linalg.inv(fr.matrix)
    return fr


def get_gridded_transformed_coords(da, matrix: np.ndarray, as_dataset=True):
    """Coordinates of every grid node of ``da`` expressed in the frame ``matrix`` maps from."""
    off = _frame_from_matrix(matrix).offsets(_axes(da))
    if as_dataset:
        return xa.Dataset({f"d_{dim}": xa.DataArray(off[..., i], coords=da.coords, dims=da.dims)
                           for i, dim in enumerate(da.dims)}, coords=da.coords)
    return off


def get_offset_grid(da, focus, origin=DEFAULT_ORIGIN, as_dataset=True):
    return get_gridded_transformed_coords(da, FocusFrame(focus, origin).matrix, as_dataset=as_dataset)


def calc_dist_from_focus(da, focus, origin=DEFAULT_ORIGIN, aspect_ratio=(1, 1, 1), as_dataarray=True):
    dist = FocusFrame(focus, origin).distance(_axes(da), aspect_ratio)
    if as_dataarray:
        return xa.DataArray(dist, coords=da.coords, dims=da.dims)
    return dist


_COMPARE = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal}


def get_mask(da, focus, distance: float, origin=DEFAULT_ORIGIN, aspect_ratio=(1, 1, 1), operator="<"):
    """Boolean ellipsoid mask around the focus (inside for '<', outside for '>')."""
    if operator not in _COMPARE:
        raise ValueError("Operator must be '<', '>', '<=', or '>='.")
    dist = FocusFrame(focus, origin).distance(_axes(da), aspect_ratio)
    return xa.DataArray(_COMPARE[operator](dist, distance), coords=da.coords, dims=da.dims)


def _line_samples(da, focus, dim, origin, min_offset, max_offset):
    frame = FocusFrame(focus, origin)
    axes = _axes(da)
    k = list(da.dims).index(dim)
    if min_offset is None or max_offset is None:
        comp = frame.offsets(axes)[..., k]
        min_offset = float(comp.min()) if min_offset is None else min_offset
        max_offset = float(comp.max()) if max_offset is None else max_offset
    n = da.sizes[dim] * 2
    offsets = np.linspace(min_offset, max_offset, n)
    return offsets, trilinear_line(_values(da), axes, frame.line(k, offsets))


def interp_transformed_axis(da, focus, dim, origin=DEFAULT_ORIGIN, min_offset: float | None = None,
                            max_offset: float | None = None):
    """``2 * size(dim)`` trilinear samples of ``da`` along the focus-frame axis matching ``dim``."""
    offsets, vals = _line_samples(da, focus, dim, origin, min_offset, max_offset)
    name = f"offset_d{dim}"
    return xa.DataArray(vals, coords={name: offsets}, dims=(name,), attrs=dict(getattr(da, "attrs", {})))


def _bounds_from_line(offsets: np.ndarray, vals: np.ndarray, cutoff: float) -> Tuple[float, float]:
    with np.errstate(invalid="ignore"):
        below = vals < float(cutoff)
    neg = np.flatnonzero(below & (offsets <= 0))
    pos = np.flatnonzero(below & (offsets >= 0))
    return (float(offsets[neg[-1]]) if neg.size else np.nan, float(offsets[pos[0]]) if pos.size else np.nan)


def get_beam_bounds(da, focus, dim, cutoff: float, origin=DEFAULT_ORIGIN, min_offset: float | None = None,
                    max_offset: float | None = None) -> Tuple[float, float]:
    """Closest offsets on either side of the focus at which the line samples fall below ``cutoff``."""
    offsets, vals = _line_samples(da, focus, dim, origin, min_offset, max_offset)
    return _bounds_from_line(offsets, vals, cutoff)


def get_beamwidth(da, focus, dim, cutoff: float | None = None, origin=DEFAULT_ORIGIN, min_offset: float | None = None,
                  max_offset: float | None = None) -> float:
    if cutoff is None:
        cutoff = float(np.nanmax(_values(da))) / 2
    neg, pos = get_beam_bounds(da, focus, dim, float(cutoff), origin=origin, min_offset=min_offset, max_offset=max_offset)
    return pos - neg
