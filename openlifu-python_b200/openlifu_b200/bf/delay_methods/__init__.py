"""Transmit delay methods (mirrors /root/reference/src/openlifu/bf/delay_methods/)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np

from .._registry import ClassKeyed, table


@dataclass
class DelayMethod(ClassKeyed, ABC):
    _family = {}

    @abstractmethod
    def calc_delays(self, arr, target, params=None, transform=None):
        ...

    @staticmethod
    def from_dict(d):
        return ClassKeyed._from_dict(DelayMethod, d)

    @abstractmethod
    def to_table(self):
        ...


@dataclass
class Direct(DelayMethod):
    """Time-of-flight focusing: delay_e = max(tof) - tof_e, tof_e = |target - pos_e| / c
    (direct.py:28-38).  ``c`` is ``params['sound_speed'].attrs['ref_value']`` whenever params are
    given (reference quirk 2), else ``c0``."""
    c0: float = 1480.0

    def __post_init__(self):
        if not isinstance(self.c0, (int, float)):
            raise TypeError("Speed of sound must be a number")
        if self.c0 <= 0:
            raise ValueError("Speed of sound must be greater than 0")
        self.c0 = float(self.c0)

    def calc_delays(self, arr, target, params=None, transform=None):
        c = self.c0 if params is None else params["sound_speed"].attrs["ref_value"]
        tgt = target.get_position(units="m")
        pos = arr.get_positions(transform=transform, units="m")
        tof = np.linalg.norm(tgt[None, :] - pos, axis=1) / c
        return np.max(tof) - tof

    def to_table(self):
        return table([{"Name": "Type", "Value": "Direct", "Unit": ""},
                      {"Name": "Default Sound Speed", "Value": self.c0, "Unit": "m/s"}])


__all__ = ["DelayMethod", "Direct"]
