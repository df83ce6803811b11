"""Pulse-train timing (mirrors /root/reference/src/openlifu/bf/sequence.py:13-80)."""
from __future__ import annotations

from dataclasses import dataclass

from ..util.dict_conversion import DictMixin
from ._registry import table


@dataclass
class Sequence(DictMixin):
    pulse_interval: float = 1.0
    pulse_count: int = 1
    pulse_train_interval: float = 1.0
    pulse_train_count: int = 1

    def __post_init__(self):
        if self.pulse_interval <= 0:
            raise ValueError("Pulse interval must be positive")
        if self.pulse_count <= 0:
            raise ValueError("Pulse count must be positive")
        if self.pulse_train_interval < 0:
            raise ValueError("Pulse train interval must be non-negative")
        if 0 < self.pulse_train_interval < self.pulse_interval * self.pulse_count:
            raise ValueError("Pulse train interval must be greater than or equal to the total pulse interval")
        if self.pulse_train_count <= 0:
            raise ValueError("Pulse train count must be positive")

    def get_pulse_train_duration(self) -> float:
        return self.pulse_interval * self.pulse_count

    def get_sequence_duration(self) -> float:
        per_train = self.pulse_train_interval if self.pulse_train_interval != 0 else self.get_pulse_train_duration()
        return per_train * self.pulse_train_count

    def to_table(self):
        return table([{"Name": "Pulse Interval", "Value": self.pulse_interval, "Unit": "s"},
                      {"Name": "Pulse Count", "Value": self.pulse_count, "Unit": ""},
                      {"Name": "Pulse Train Interval", "Value": self.pulse_train_interval, "Unit": "s"},
                      {"Name": "Pulse Train Count", "Value": self.pulse_train_count, "Unit": ""}])
