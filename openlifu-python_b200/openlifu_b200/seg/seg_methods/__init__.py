"""Segmentation methods (mirrors /root/reference/src/openlifu/seg/seg_methods/uniform.py:10-66).
Only uniform segmenters exist in the reference; heterogeneous maps are built by passing a label
volume through ``SegmentationMethod._map_params`` (see ``LabelVolume`` below)."""
from __future__ import annotations

import pandas as pd

from ..material import MATERIALS, Material
from ..seg_method import SegmentationMethod


class UniformSegmentation(SegmentationMethod):
    def _segment(self, volume):
        return self._ref_segment(volume.coords)

    def to_table(self) -> pd.DataFrame:
        return pd.DataFrame.from_records([{"Name": "Type", "Value": "Uniform", "Unit": ""},
                                          {"Name": "Reference Material", "Value": self.ref_material, "Unit": ""}])


class _FixedUniform(UniformSegmentation):
    _ref = "water"
    _label = "Uniform"

    def __init__(self, materials: dict[str, Material] | None = None):
        super().__init__(materials=MATERIALS.copy() if materials is None else materials, ref_material=self._ref)

    def to_table(self) -> pd.DataFrame:
        return pd.DataFrame.from_records([{"Name": "Type", "Value": self._label, "Unit": ""}])

    def to_dict(self):
        d = super().to_dict()
        d.pop("ref_material")
        return d


class UniformTissue(_FixedUniform):
    """Every voxel is tissue."""
    _ref = "tissue"
    _label = "Uniform Tissue"


class UniformWater(_FixedUniform):
    """Every voxel is water."""
    _ref = "water"
    _label = "Uniform Water"


class LabelVolume(SegmentationMethod):
    """Treat the input volume itself as the integer label map (label = index into ``materials``).
    Not in the reference; it exposes ``_map_params`` for synthetic heterogeneous phantoms
    (SURVEY.md 8d config C3)."""

    def _segment(self, volume):
        return volume

    def to_table(self) -> pd.DataFrame:
        return pd.DataFrame.from_records([{"Name": "Type", "Value": "Label Volume", "Unit": ""},
                                          {"Name": "Reference Material", "Value": self.ref_material, "Unit": ""}])


__all__ = ["UniformSegmentation", "UniformTissue", "UniformWater", "LabelVolume"]
