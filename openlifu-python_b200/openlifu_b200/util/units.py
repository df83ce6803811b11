"""Unit-string handling (host glue of the hot path).

Mirrors the behaviour of /root/reference/src/openlifu/util/units.py (``getunittype:7``,
``getunitconversion:36``, ``getsiscale:96``, ``rescale_data_arr:182``, ``rescale_coords:201``,
``get_ndgrid_from_arr:225``) with a table-driven parser: a unit is <SI prefix><base>, the base
decides the quantity, ratios "a/b" convert numerator and denominator separately.
"""
from __future__ import annotations

import numpy as np

_PREFIX = {
    "pico": 1e-12, "p": 1e-12, "nano": 1e-9, "n": 1e-9,
    "micro": 1e-6, "u": 1e-6, "µ": 1e-6, "μ": 1e-6,
    "milli": 1e-3, "m": 1e-3, "centi": 1e-2, "c": 1e-2, "": 1.0,
    "kilo": 1e3, "k": 1e3, "mega": 1e6, "M": 1e6, "giga": 1e9, "G": 1e9, "tera": 1e12, "T": 1e12,
    "min": 60.0, "minute": 60.0, "hour": 3600.0, "hr": 3600.0, "day": 86400.0, "d": 86400.0,
    "rad": 1.0, "radian": 1.0, "radians": 1.0,
    "deg": 2 * 3.14159265358979323846 / 360, "degree": 2 * 3.14159265358979323846 / 360,
    "degrees": 2 * 3.14159265358979323846 / 360, "°": 2 * 3.14159265358979323846 / 360,
}

_TIME_WORDS = ("minute", "minutes", "min", "mins", "hour", "hours", "hr", "hrs", "day", "days", "d")
_ANGLE_WORDS = ("rad", "deg", "radian", "radians", "degree", "degrees", "°")


def getunittype(unit: str) -> str:
    u = unit.lower()
    if u in ("micron", "microns"):
        return "distance"
    if u in _TIME_WORDS:
        return "time"
    if u in _ANGLE_WORDS:
        return "angle"
    if "sec" in u:
        return "time"
    if "meter" in u or "micron" in u:
        return "distance"
    for suffix, kind in (("s", "time"), ("m", "distance"), (("m2", "m^2"), "area"), (("m3", "m^3"), "volume"),
                         ("hz", "frequency"), ("pa", "pressure"), ("w", "watt")):
        if u.endswith(suffix):
            return kind
    return "other"


def _first_found(unit: str, needles, fallback_char: str) -> int:
    for nd in needles:
        i = unit.find(nd)
        if i != -1:
            return i
    i = unit.rfind(fallback_char)
    return len(unit) if i == -1 else i


def getsiscale(unit: str, type: str) -> float:  # noqa: A002 - reference signature
    kind = type.lower()
    if kind in ("distance", "area", "volume"):
        if unit.find("meter") == -1 and unit.lower() == "micron":
            cut = 6
        else:
            cut = _first_found(unit, ("meters", "meter"), "m")
    elif kind == "time":
        cut = _first_found(unit, ("seconds", "second", "sec"), "s")
    elif kind == "angle":
        cut = len(unit)
    elif kind in ("frequency", "pressure"):
        cut = len(unit) - 2
    elif kind == "watt":
        cut = len(unit) - 1
    else:
        cut = len(unit) - len(kind) + 1
    prefix = unit[:cut]
    if prefix not in _PREFIX:
        raise ValueError(f"Unknown prefix {prefix}")
    scale = _PREFIX[prefix]
    if kind == "area":
        scale = scale ** 2.0
    elif kind == "volume":
        scale = scale ** 3.0
    return scale


def getunitconversion(from_unit, to_unit, unitratio=None, constant=None):
    if not from_unit:
        return 1.0
    if unitratio is not None and constant is not None:
        if "/" not in unitratio:
            raise ValueError("Conversion unit ratio must have a '/' symbol")
        num, den = unitratio.split("/")
        t_from, t_to, t_num, t_den = (getunittype(v) for v in (from_unit, to_unit, num, den))
        if t_from == t_den and t_to == t_num:
            return getunitconversion(from_unit, den) * constant * getunitconversion(num, to_unit)
        if t_from == t_num and t_to == t_den:
            return getunitconversion(from_unit, num) * 1 / constant * getunitconversion(den, to_unit)
        if t_from == t_to:
            return getunitconversion(from_unit, to_unit)
        raise ValueError(f"Unit type mismatch {t_from} -> ({t_num}/{t_den}) -> {t_to}")
    s_from, s_to = from_unit.find("/"), to_unit.find("/")
    if s_from != -1 and s_to != -1:
        return (getunitconversion(from_unit[:s_from], to_unit[:s_to])
                / getunitconversion(from_unit[s_from + 1:], to_unit[s_to + 1:]))
    if s_from != -1 or s_to != -1:
        raise ValueError(f"Unit ratio mismatch ({from_unit} vs {to_unit})")
    t_from, t_to = getunittype(from_unit), getunittype(to_unit)
    if t_from != t_to:
        raise ValueError(f"Unit type mismatch ({t_from}) vs ({t_to})")
    if t_from == "other":
        if from_unit[-1] != to_unit[-1]:
            raise ValueError(f"Cannot convert {from_unit} to {to_unit}")
        common = ""
        i = 0
        while i < min(len(from_unit), len(to_unit)) and from_unit[-i:] == to_unit[-i:]:
            common = from_unit[-i:]
            i += 1
        return getsiscale(from_unit, common) / getsiscale(to_unit, common)
    return getsiscale(from_unit, t_from) / getsiscale(to_unit, t_from)


def rescale_data_arr(data_arr, units: str):
    """Copy of ``data_arr`` with values converted to ``units`` (attrs['units'] updated)."""
    out = data_arr.copy(deep=True)
    out.data *= getunitconversion(data_arr.attrs["units"], units)
    out.attrs["units"] = units
    return out


def rescale_coords(data_arr, units: str):
    """Copy of ``data_arr`` whose unit-carrying coordinates are converted to ``units``."""
    out = data_arr.copy(deep=True)
    for key in data_arr.coords:
        attrs = out[key].attrs
        if "units" in attrs:
            scaled = getunitconversion(attrs["units"], units) * out[key].data
            out = out.assign_coords({key: (key, scaled, attrs)})
            out[key].attrs["units"] = units
    return out


def get_ndgrid_from_arr(data_arr) -> np.ndarray:
    """(…, ndim) array of coordinates, ``ij`` indexing, in the dim order of the first variable."""
    first = next(iter(data_arr.keys()))
    axes = [data_arr.coords[k].data for k in data_arr[first].dims if "units" in data_arr[k].attrs]
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1)
