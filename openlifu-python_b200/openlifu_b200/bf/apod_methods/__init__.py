"""Transmit apodization methods (mirrors /root/reference/src/openlifu/bf/apod_methods/:
uniform.py:21-22, maxangle.py:33-39, piecewiselinear.py:42-49)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np

from ...util.units import getunittype
from .._registry import ClassKeyed, table


@dataclass
class ApodizationMethod(ClassKeyed, ABC):
    _family = {}

    @abstractmethod
    def calc_apodization(self, arr, target, params=None, transform=None):
        ...

    @staticmethod
    def from_dict(d):
        return ClassKeyed._from_dict(ApodizationMethod, d)

    @abstractmethod
    def to_table(self):
        ...


def _element_angles(arr, target, transform, units):
    tgt = target.get_position(units="m")
    m = np.eye(4) if transform is None else transform
    return np.array([el.angle_to_point(tgt, units="m", matrix=m, return_as=units) for el in arr.elements])


def _check_angle(name, value):
    if not isinstance(value, (int, float)):
        raise TypeError(f"{name} must be a number, got {type(value).__name__}.")
    if value < 0:
        raise ValueError(f"{name} must be non-negative, got {value}.")


@dataclass
class Uniform(ApodizationMethod):
    value: float = 1.0

    def calc_apodization(self, arr, target=None, params=None, transform=None):
        return np.full(arr.numelements(), self.value)

    def to_table(self):
        return table([{"Name": "Type", "Value": "Uniform", "Unit": ""},
                      {"Name": "Value", "Value": self.value, "Unit": ""}])


@dataclass
class MaxAngle(ApodizationMethod):
    """1 where the angle between the element normal and the ray to the target is <= max_angle."""
    max_angle: float = 30.0
    units: str = "deg"

    def __post_init__(self):
        _check_angle("Max angle", self.max_angle)
        if getunittype(self.units) != "angle":
            raise ValueError(f"Units must be an angle type, got {self.units}.")

    def calc_apodization(self, arr, target, params=None, transform=None):
        return (_element_angles(arr, target, transform, self.units) <= self.max_angle).astype(np.float64)

    def to_table(self):
        return table([{"Name": "Type", "Value": "Max Angle", "Unit": ""},
                      {"Name": "Max Angle", "Value": self.max_angle, "Unit": self.units}])


@dataclass
class PiecewiseLinear(ApodizationMethod):
    """1 below rolloff_angle, linear ramp to 0 at zero_angle."""
    zero_angle: float = 90.0
    rolloff_angle: float = 45.0
    units: str = "deg"

    def __post_init__(self):
        _check_angle("Zero angle", self.zero_angle)
        _check_angle("Rolloff angle", self.rolloff_angle)
        if self.rolloff_angle >= self.zero_angle:
            raise ValueError(f"Rolloff angle must be less than zero angle, got {self.rolloff_angle} >= {self.zero_angle}.")
        if getunittype(self.units) != "angle":
            raise ValueError(f"Units must be an angle type, got {self.units}.")

    def calc_apodization(self, arr, target, params=None, transform=None):
        ang = _element_angles(arr, target, transform, self.units)
        return np.clip((self.zero_angle - ang) / (self.zero_angle - self.rolloff_angle), 0, 1)

    def to_table(self):
        return table([{"Name": "Type", "Value": "Piecewise-Linear", "Unit": ""},
                      {"Name": "Zero Angle", "Value": self.zero_angle, "Unit": self.units},
                      {"Name": "Rolloff Angle", "Value": self.rolloff_angle, "Unit": self.units}])


__all__ = ["ApodizationMethod", "Uniform", "MaxAngle", "PiecewiseLinear"]
