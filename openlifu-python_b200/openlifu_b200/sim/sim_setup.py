"""Simulation grid definition.

Mirrors /root/reference/src/openlifu/sim/sim_setup.py (``SimSetup:22``, extent snapping
``:91-105``, ``get_coords:107``, ``get_size:152``, ``setup_sim_scene:161``).  Axis names are fixed
``('x','y','z')`` with long names (Lateral, Elevation, Axial).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

from .. import xa
from ..util.dict_conversion import DictMixin
from ..util.units import getunitconversion, getunittype

COORD_DIMS = ("x", "y", "z")
COORD_NAMES = ("Lateral", "Elevation", "Axial")


def _number(name, value, positive=True):
    if not isinstance(value, (int, float)):
        raise TypeError(f"{name} must be a number.")
    if positive and value <= 0:
        raise ValueError(f"{name} must be a positive number.")
    if not positive and value < 0:
        raise ValueError(f"{name} must be a non-negative number.")


@dataclass
class SimSetup(DictMixin):
    spacing: float = 1.0
    units: str = "mm"
    x_extent: Tuple[float, float] = (-30., 30.)
    y_extent: Tuple[float, float] = (-30., 30.)
    z_extent: Tuple[float, float] = (-4., 60.)
    dt: float = 0.
    t_end: float = 0.
    c0: float = 1500.0
    cfl: float = 0.5
    options: dict = field(default_factory=dict)

    def __post_init__(self):
        for ax in COORD_DIMS:
            ext = getattr(self, f"{ax}_extent")
            if len(ext) != 2:
                raise ValueError(f"{ax}_extent must have length 2.")
            if ext[0] >= ext[1]:
                raise ValueError(f"{ax}_extent must be in the form (min, max) with min < max.")
        _number("spacing", self.spacing)
        if not isinstance(self.units, str):
            raise TypeError("units must be a string.")
        if getunittype(self.units) != "distance":
            raise ValueError(f"units must be a length unit, got {self.units}.")
        _number("c0", self.c0)
        _number("cfl", self.cfl)
        _number("dt", self.dt, positive=False)
        _number("t_end", self.t_end, positive=False)
        # snap every extent to a whole number of cells, keeping the lower bound
        for ax in COORD_DIMS:
            ext = getattr(self, f"{ax}_extent")
            cells = np.diff(ext) / self.spacing
            snapped = tuple(np.arange(2) * np.round(cells) * self.spacing + ext[0])
            if ((0.5 - np.abs((cells % 1) - 0.5)) / np.round(cells)) > 1e-3:
                logging.warning(f"{ax}_extent {ext} does not evenly divide by spacing ({self.spacing}). "
                                f"Rounding to {snapped}.")
            setattr(self, f"{ax}_extent", snapped)

    def _extents(self):
        return [self.x_extent, self.y_extent, self.z_extent]

    def get_size(self, dims=None):
        dims = COORD_DIMS if dims is None else dims
        n = [int(np.round(np.diff(ext) / self.spacing).item()) + 1 for ext in self._extents()]
        return np.array([n[COORD_DIMS.index(d)] for d in dims]).squeeze()

    def get_extent(self, dims=None, units: str | None = None):
        dims = COORD_DIMS if dims is None else dims
        scale = getunitconversion(self.units, self.units if units is None else units)
        ext = self._extents()
        return np.array([ext[COORD_DIMS.index(d)] for d in dims]) * scale

    def get_spacing(self, units: str | None = None):
        return getunitconversion(self.units, self.units if units is None else units) * self.spacing

    def get_coords(self, dims=None, units: str | None = None):
        dims = COORD_DIMS if dims is None else dims
        units = self.units if units is None else units
        sizes = np.atleast_1d(self.get_size(dims))
        extents = self.get_extent(dims, units)
        coords = xa.Coordinates({d: np.linspace(extents[i][0], extents[i][1], int(sizes[i])) for i, d in enumerate(dims)})
        for d in dims:
            coords[d].attrs["units"] = units
            coords[d].attrs["long_name"] = COORD_NAMES[COORD_DIMS.index(d)]
        return coords

    def get_corners(self, units: str | None = None):
        scale = getunitconversion(self.units, self.units if units is None else units)
        xyz = np.array(np.meshgrid(self.x_extent, self.y_extent, self.z_extent, indexing="ij"))
        return xyz.reshape(3, -1) * scale

    def get_max_distance(self, arr, units: str | None = None):
        units = self.units if units is None else units
        corners = self.get_corners(units=units)
        pos = arr.get_positions(units=units)
        return float(np.max(np.linalg.norm(pos[:, :, None] - corners[None, :, :], axis=1)))

    def setup_sim_scene(self, seg_method, volume=None):
        """Medium parameter Dataset on the simulation grid: uniform reference material when no
        volume is given, else the segmented volume (assumed already resampled on the grid)."""
        if volume is None:
            return seg_method.ref_params(self.get_coords())
        return seg_method.seg_params(volume)

    def to_table(self):
        import pandas as pd
        rows = [("Spacing", self.spacing, self.units)]
        rows += [(f"{ax.upper()} Extent", f"{e[0]} to {e[1]}", self.units) for ax, e in zip(COORD_DIMS, self._extents())]
        rows += [("Time Step", self.dt, "s"), ("End Time", self.t_end, "s"), ("Speed of Sound", self.c0, "m/s"),
                 ("CFL", self.cfl, "")]
        return pd.DataFrame.from_records([{"Name": n, "Value": v, "Unit": u} for n, v, u in rows])

    @staticmethod
    def from_dict(d: dict, on_keyword_mismatch: str = "warn") -> "SimSetup":
        if not isinstance(d, dict):
            raise TypeError("Input must be a dictionary.")
        known = ["spacing", "units", "x_extent", "y_extent", "z_extent", "dt", "t_end", "c0", "cfl", "options"]
        extra = [k for k in d if k not in known]
        if extra:
            if on_keyword_mismatch == "raise":
                raise TypeError(f"Unexpected keyword arguments for SimSetup: {extra}")
            if on_keyword_mismatch == "warn":
                logging.warning(f"Ignoring unexpected keyword arguments for SimSetup: {extra}")
        return SimSetup(**{k: v for k, v in d.items() if k in known})
