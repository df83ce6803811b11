"""Points and focus frames (beamforming input).

API mirror of ``Point`` from /root/reference/src/openlifu/geo.py:18-74 without the vtk
rendering helpers (out of the hot path).  ``get_matrix`` builds the focus frame whose z axis
points from the origin to the point and whose x axis stays in the x-z plane.
"""
from __future__ import annotations

import copy
import json
from dataclasses import dataclass, field
from typing import Any, Tuple

import numpy as np

from .util.units import getunitconversion


def _unit(v):
    n = np.linalg.norm(v)
    return v / n if n != 0 else None


@dataclass
class Point:
    position: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 0.0]))
    id: str = "point"
    name: str = "Point"
    color: Any = (1.0, 0.0, 0.0)
    radius: float = 1.0
    dims: Tuple[str, str, str] = ("x", "y", "z")
    units: str = "mm"

    def __post_init__(self):
        if len(self.position) != len(self.dims):
            raise ValueError("Position and dims must have same length.")
        self.position = np.array(self.position).reshape(3)

    def copy(self):
        return copy.deepcopy(self)

    def get_position(self, dim=None, units: str | None = None):
        scale = getunitconversion(self.units, self.units if units is None else units)
        if dim is None:
            return self.position * scale
        return self.position[self.dims.index(dim)] * scale

    def get_matrix(self, origin: np.ndarray = np.eye(4), center_on_point: bool = True, local: bool = False):
        """4x4 focus frame (geo.py:56-74): columns (x', y', z', centre)."""
        rel = (np.linalg.inv(origin) @ np.append(self.position, 1.0))[:3]
        zhat = _unit(rel)
        if zhat is None:
            zhat = np.array([0.0, 0.0, 1.0])
        az = -np.arctan2(zhat[0], zhat[2])
        xhat = np.array([np.cos(az), 0.0, np.sin(az)])
        frame = np.eye(4)
        frame[:3, 0] = xhat
        frame[:3, 1] = np.cross(zhat, xhat)
        frame[:3, 2] = zhat
        frame[:3, 3] = rel if center_on_point else 0.0
        return frame if local else origin @ frame

    def rescale(self, units: str):
        scale = getunitconversion(self.units, units)
        self.position = self.position * scale
        self.radius = self.radius * scale
        self.units = units

    def transform(self, matrix: np.ndarray, units: str | None = None, new_dims=None):
        if units is not None:
            self.rescale(units)
        self.position = (matrix @ np.append(self.position, 1.0))[:3]
        if new_dims is not None:
            self.dims = new_dims

    def to_dict(self):
        return {"id": self.id, "name": self.name, "color": self.color, "radius": self.radius,
                "position": self.position.tolist(), "dims": self.dims, "units": self.units}

    @staticmethod
    def from_dict(d):
        d = dict(d)
        if isinstance(d.get("dims"), list):
            d["dims"] = tuple(d["dims"])
        if isinstance(d.get("color"), list):
            d["color"] = tuple(d["color"])
        return Point(**d)

    def to_json(self, compact: bool = False) -> str:
        return json.dumps(self.to_dict(), separators=(",", ":")) if compact else json.dumps(self.to_dict(), indent=4)

    @staticmethod
    def from_json(s: str) -> "Point":
        return Point.from_dict(json.loads(s))
