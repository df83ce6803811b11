"""Threshold checks on analysis parameters (mirrors /root/reference/src/openlifu/plan/param_constraint.py:18-95).

A constraint states what a value must satisfy ("value < 1.9"); a value that does NOT satisfy the
warning (error) condition raises a warning (error) status.
"""
from __future__ import annotations

import operator as _op
from dataclasses import dataclass
from typing import Tuple, Union

import pandas as pd

from ..util.dict_conversion import DictMixin

PARAM_STATUS_SYMBOLS = {"ok": "✅", "warning": "❗", "error": "❌"}

Number = Union[float, int]
_SCALAR_OPS = {"<": _op.lt, "<=": _op.le, ">": _op.gt, ">=": _op.ge}
_RANGE_OPS = {
    "within": lambda v, lo, hi: lo < v < hi,
    "inside": lambda v, lo, hi: lo <= v <= hi,
    "outside": lambda v, lo, hi: v < lo or v > hi,
    "outside_inclusive": lambda v, lo, hi: v <= lo or v >= hi,
}


def _is_number(v) -> bool:
    return isinstance(v, (int, float))


@dataclass
class ParameterConstraint(DictMixin):
    operator: str = "<="
    warning_value: Union[Number, Tuple[Number, Number], None] = None
    error_value: Union[Number, Tuple[Number, Number], None] = None

    def __post_init__(self):
        if self.warning_value is None and self.error_value is None:
            raise ValueError("At least one of warning_value or error_value must be set")
        if self.operator in _RANGE_OPS:
            for label, v in (("Warning", self.warning_value), ("Error", self.error_value)):
                if v and (not isinstance(v, tuple) or len(v) != 2 or v[0] >= v[1]):
                    raise ValueError(f"{label} value must be a sorted tuple of two numbers")
        elif self.operator in _SCALAR_OPS:
            for label, v in (("Warning", self.warning_value), ("Error", self.error_value)):
                if v is not None and not _is_number(v):
                    raise ValueError(f"{label} value must be a single value")

    @staticmethod
    def compare(value, operator, threshold) -> bool:
        if operator in _SCALAR_OPS:
            return _SCALAR_OPS[operator](value, threshold)
        if operator in _RANGE_OPS:
            return _RANGE_OPS[operator](value, threshold[0], threshold[1])
        raise ValueError(f"Unsupported operator: {operator}")

    def _violates(self, value, threshold) -> bool:
        return threshold is not None and not self.compare(value, self.operator, threshold)

    def is_warning(self, value: Number) -> bool:
        return self._violates(value, self.warning_value)

    def is_error(self, value: Number) -> bool:
        return self._violates(value, self.error_value)

    def get_status(self, value: float) -> str:
        return "error" if self.is_error(value) else ("warning" if self.is_warning(value) else "ok")

    def get_status_symbol(self, value: float) -> str:
        return PARAM_STATUS_SYMBOLS[self.get_status(value)]

    def to_table(self) -> pd.DataFrame:
        if self.operator not in _SCALAR_OPS and self.operator not in _RANGE_OPS:
            raise ValueError(f"Unsupported operator: {self.operator}")
        rows = [{"Name": name, "Value": f"value {self.operator} {v}", "Unit": ""}
                for name, v in (("Warn if not", self.warning_value), ("Error if not", self.error_value)) if v is not None]
        return pd.DataFrame.from_records(rows)
