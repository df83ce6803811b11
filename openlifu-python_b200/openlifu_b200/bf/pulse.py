"""Sinusoidal pulse (mirrors /root/reference/src/openlifu/bf/pulse.py:13-62)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..util.dict_conversion import DictMixin
from ._registry import table


@dataclass
class Pulse(DictMixin):
    frequency: float = 1.0
    amplitude: float = 1.0
    duration: float = 1.0

    def __post_init__(self):
        if self.frequency <= 0:
            raise ValueError("Frequency must be greater than 0")
        if not 0 <= self.amplitude <= 1:
            raise ValueError("Amplitude must be between 0 and 1")
        if self.duration <= 0:
            raise ValueError("Duration must be greater than 0")

    def calc_pulse(self, t):
        return self.amplitude * np.sin(2 * np.pi * self.frequency * t)

    def calc_time(self, dt: float):
        return np.arange(0, self.duration, dt)

    def to_table(self):
        return table([{"Name": "Frequency", "Value": self.frequency, "Unit": "Hz"},
                      {"Name": "Amplitude", "Value": self.amplitude, "Unit": "AU"},
                      {"Name": "Duration", "Value": self.duration, "Unit": "s"}])
