"""Label volume -> per-voxel medium parameter maps.

Mirrors /root/reference/src/openlifu/seg/seg_method.py (``SegmentationMethod:18``,
``_material_indices:80``, ``_map_params:84-97``, ``seg_params:99``, ``ref_params:105``,
``_ref_segment:110``).  The parameter Dataset (five float64 maps with ``units``, ``long_name``,
``ref_value`` attrs) is the medium input of ``run_simulation``.
"""
from __future__ import annotations

import copy
import inspect
import logging
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from .. import xa
from .material import MATERIALS, PARAM_INFO, Material


@dataclass
class SegmentationMethod(ABC):
    materials: dict = field(default_factory=lambda: MATERIALS.copy())
    ref_material: str = "water"

    def __post_init__(self):
        if self.materials is None:
            self.materials = MATERIALS.copy()
        if not isinstance(self.materials, dict):
            raise TypeError(f"Materials must be a dictionary, got {type(self.materials).__name__}.")
        if not all(isinstance(m, Material) for m in self.materials.values()):
            raise TypeError("All materials must be instances of Material class.")
        if self.ref_material not in self.materials:
            raise ValueError(f"Reference material {self.ref_material} not found.")

    @abstractmethod
    def _segment(self, volume):
        ...

    @abstractmethod
    def to_table(self):
        ...

    def to_dict(self) -> dict[str, Any]:
        d = dict(self.__dict__)
        d["materials"] = {k: m.to_dict() for k, m in self.materials.items()}
        d["class"] = type(self).__name__
        return d

    @staticmethod
    def from_dict(d: dict, on_keyword_mismatch: str = "warn") -> "SegmentationMethod":
        from . import seg_methods
        if not isinstance(d, dict):
            raise TypeError(f"Expected dict for from_dict, got {type(d).__name__}")
        d = copy.deepcopy(d)
        cls = getattr(seg_methods, d.pop("class"))
        if d.get("materials") is not None:
            d["materials"] = {k: (m if isinstance(m, Material) else Material.from_dict(m))
                              for k, m in d["materials"].items()}
        accepted = {p.name for p in inspect.signature(cls).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD}
        extra = [k for k in d if k not in accepted]
        if extra:
            if on_keyword_mismatch == "raise":
                raise TypeError(f"Unexpected keyword arguments for {cls.__name__}: {extra}")
            if on_keyword_mismatch == "warn":
                logging.warning(f"Ignoring unexpected keyword arguments for {cls.__name__}: {extra}")
            for k in extra:
                d.pop(k)
        return cls(**d)

    def _material_indices(self, materials: dict | None = None):
        """Label value of a material = its position in the materials dict."""
        materials = self.materials if materials is None else materials
        return {mid: i for i, mid in enumerate(materials)}

    def _map_params(self, seg, materials: dict | None = None):
        materials = self.materials if materials is None else materials
        ref = materials[self.ref_material]
        labels = np.asarray(seg.data)
        # one gather through a per-label lookup table instead of a boolean pass per material;
        # labels outside the table keep the reference's initial value 0
        n_mat = len(materials)
        lo, hi = (int(labels.min()), int(labels.max())) if labels.size else (0, 0)
        uniform = lo == hi                                        # e.g. the reference-material volume of UniformWater
        if not uniform:
            valid = (labels >= 0) & (labels < n_mat)
            safe = np.where(valid, labels, 0).astype(np.intp)
        params = xa.Dataset()
        for pid in PARAM_INFO:
            info = Material.param_info(pid)
            lut = np.array([getattr(m, pid) for m in materials.values()], dtype=np.float64)
            if uniform:
                data = np.full(labels.shape, lut[lo] if 0 <= lo < n_mat else 0.0, dtype=np.float64)
            else:
                data = np.where(valid, lut[safe], 0.0)
            params[pid] = xa.DataArray(data, coords=seg.coords, dims=seg.dims,
                                       attrs={"units": info["units"], "long_name": info["name"],
                                              "ref_value": ref.get_param(pid)})
        params.attrs["ref_material"] = ref
        if not uniform and n_mat <= 32 and bool(valid.all()):
            # provenance for the solver adapter: while the three acoustic maps still hold what was expanded here
            # (content keys), run_simulation uploads the label volume + tables instead of the maps (lifu_set_medium_labels)
            from ..util.content import content_key
            solver_ids = ("sound_speed", "density", "attenuation")
            params.attrs["lifu_label_medium"] = {
                "labels": labels.astype(np.uint8),
                "lut": {pid: np.array([getattr(m, pid) for m in materials.values()], dtype=np.float64) for pid in solver_ids},
                "keys": tuple(content_key(params[pid].data) for pid in solver_ids)}
        return params

    def seg_params(self, volume, materials: dict | None = None):
        materials = self.materials if materials is None else materials
        return self._map_params(self._segment(volume), materials=materials)

    def ref_params(self, coords):
        return self._map_params(self._ref_segment(coords))

    def _ref_segment(self, coords):
        label = self._material_indices()[self.ref_material]
        shape = list(coords.sizes.values())
        return xa.DataArray(np.full(shape, label, dtype=int), coords=coords, dims=tuple(coords.dims))
