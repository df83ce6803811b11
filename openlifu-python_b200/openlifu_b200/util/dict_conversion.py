"""dataclass <-> dict mixin (mirrors /root/reference/src/openlifu/util/dict_conversion.py:8-37)."""
from __future__ import annotations

from dataclasses import asdict, fields
from typing import Any, Dict, get_origin

import numpy as np


class DictMixin:
    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @classmethod
    def from_dict(cls, parameter_dict: Dict[str, Any]):
        params = {k: v for k, v in parameter_dict.items() if k != "class"}
        obj = cls(**params)
        for f in fields(cls):
            t = f.type
            if t is np.ndarray or get_origin(t) is np.ndarray or (isinstance(t, str) and "np.ndarray" in t):
                setattr(obj, f.name, np.array(getattr(obj, f.name)))
        return obj
