"""Focal patterns: target -> list of foci (mirrors /root/reference/src/openlifu/bf/focal_patterns/:
focal_pattern.py:16-88, single.py:18-25, wheel.py:41-69)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np

from ...geo import Point
from ...util.units import getunittype
from .._registry import ClassKeyed, table


@dataclass
class FocalPattern(ClassKeyed, ABC):
    _family = {}
    target_pressure: float = 1.0
    units: str = "Pa"

    def __post_init__(self):
        if self.target_pressure <= 0:
            raise ValueError("Target pressure must be greater than 0")
        if not isinstance(self.units, str):
            raise TypeError("Units must be a string")
        if getunittype(self.units) != "pressure":
            raise ValueError(f"Units must be a pressure unit, got {self.units}")

    @abstractmethod
    def get_targets(self, target: Point):
        ...

    @abstractmethod
    def num_foci(self):
        ...

    @staticmethod
    def from_dict(d):
        return ClassKeyed._from_dict(FocalPattern, d)

    @abstractmethod
    def to_table(self):
        ...


@dataclass
class SinglePoint(FocalPattern):
    def get_targets(self, target: Point):
        return [target.copy()]

    def num_foci(self):
        return 1

    def to_table(self):
        return table([{"Name": "Type", "Value": "Single Point", "Unit": ""},
                      {"Name": "Target Pressure", "Value": self.target_pressure, "Unit": self.units}])


@dataclass
class Wheel(FocalPattern):
    """Optional centre plus ``num_spokes`` foci on a circle of ``spoke_radius`` in the plane normal
    to the origin->target ray.  Reference quirk 7 is kept: the radius is added to the target
    position in the TARGET's units, while spoke points are labelled with ``distance_units``."""
    center: bool = True
    num_spokes: int = 4
    spoke_radius: float = 1.0
    distance_units: str = "mm"

    def __post_init__(self):
        if not isinstance(self.center, bool):
            raise TypeError(f"Center must be a boolean, got {type(self.center).__name__}.")
        if not isinstance(self.num_spokes, int) or self.num_spokes < 1:
            raise ValueError(f"Number of spokes must be a positive integer, got {self.num_spokes}.")
        if not isinstance(self.spoke_radius, (int, float)) or self.spoke_radius <= 0:
            raise ValueError(f"Spoke radius must be a positive number, got {self.spoke_radius}.")
        super().__post_init__()

    def get_targets(self, target: Point):
        foci = []
        if self.center:
            hub = target.copy()
            hub.id = f"{target.id} (Center)"
            foci.append(hub)
        frame = target.get_matrix(center_on_point=True)
        for theta in 2 * np.pi * np.arange(self.num_spokes) / self.num_spokes:
            rim = np.append(self.spoke_radius * np.array([np.cos(theta), np.sin(theta), 0.0]), 1.0)
            deg = np.rad2deg(theta)
            foci.append(Point(id=f"{target.id}_{deg:.0f}deg", name=f"{target.name} ({deg:.0f}°)",
                              position=(frame @ rim)[:3], units=self.distance_units, radius=target.radius))
        return foci

    def num_foci(self) -> int:
        return int(self.center) + self.num_spokes

    def to_table(self):
        return table([{"Name": "Type", "Value": "Wheel", "Unit": ""},
                      {"Name": "Target Pressure", "Value": self.target_pressure, "Unit": self.units},
                      {"Name": "Center", "Value": self.center, "Unit": ""},
                      {"Name": "Number of Spokes", "Value": self.num_spokes, "Unit": ""},
                      {"Name": "Spoke Radius", "Value": self.spoke_radius, "Unit": self.distance_units}])


__all__ = ["FocalPattern", "SinglePoint", "Wheel"]
