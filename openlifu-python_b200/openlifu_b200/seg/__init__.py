from . import seg_methods
from .material import AIR, MATERIALS, PARAM_INFO, SKULL, STANDOFF, TISSUE, WATER, Material
from .seg_method import SegmentationMethod

__all__ = ["Material", "MATERIALS", "PARAM_INFO", "WATER", "TISSUE", "SKULL", "AIR", "STANDOFF",
           "SegmentationMethod", "seg_methods"]
