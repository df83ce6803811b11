"""Allowed range of a target coordinate (mirrors /root/reference/src/openlifu/plan/target_constraints.py:15-71)."""
from __future__ import annotations

import logging
from dataclasses import dataclass

import pandas as pd

from ..util.dict_conversion import DictMixin
from ..util.units import getunittype


@dataclass
class TargetConstraints(DictMixin):
    dim: str = "x"
    name: str = "dim"
    units: str = "m"
    min: float = float("-inf")
    max: float = float("inf")

    def __post_init__(self):
        for label, v in (("Dimension ID", self.dim), ("Dimension name", self.name), ("Dimension units", self.units)):
            if not isinstance(v, str):
                raise TypeError(f"{label} must be a string")
        if getunittype(self.units) != "distance":
            raise ValueError(f"Units must be a length unit, got {self.units}")
        for label, v in (("Minimum", self.min), ("Maximum", self.max)):
            if not isinstance(v, (int, float)):
                raise TypeError(f"{label} value must be a number")
        if self.min > self.max:
            raise ValueError("Minimum value cannot be greater than maximum value")

    def check_bounds(self, pos: float):
        if pos < self.min or pos > self.max:
            msg = f"The position {pos} at dimension {self.name} is not within bounds [{self.min}, {self.max}]!"
            logging.error(msg=msg)
            raise ValueError(msg)

    def to_table(self) -> pd.DataFrame:
        return pd.DataFrame.from_records([{"Name": self.name, "Value": f"({self.min},{self.max})", "Unit": self.units}])
