"""openlifu_b200 -- B200-native drop-in for OpenLIFU's treatment-planning hot path.

Mirrors the names of the reference package (/root/reference/src/openlifu/__init__.py) for the
components on that path only: beamforming inputs (``bf``), transducer geometry (``xdc``),
medium maps (``seg``), grid setup and the solver boundary (``sim.run_simulation``), and the
planning API that consumes it (``plan``).  Everything else of OpenLIFU (db, io, nav, cloud,
virtual fit) is out of scope.
"""
from __future__ import annotations

from . import bf, geo, seg, sim, util, xa, xdc
from .bf import ApodizationMethod, DelayMethod, FocalPattern, Pulse, Sequence, apod_methods, delay_methods, focal_patterns
from .geo import Point
from .seg import AIR, MATERIALS, SKULL, STANDOFF, TISSUE, WATER, Material, SegmentationMethod, seg_methods
from .sim import SimSetup
from .xdc import Transducer

__version__ = "0.1.0"

__all__ = ["Point", "Transducer", "Material", "SegmentationMethod", "seg_methods", "MATERIALS", "WATER", "TISSUE",
           "SKULL", "AIR", "STANDOFF", "DelayMethod", "ApodizationMethod", "Pulse", "Sequence", "FocalPattern",
           "focal_patterns", "delay_methods", "apod_methods", "SimSetup", "bf", "geo", "seg", "sim", "xdc", "xa", "util"]


def __getattr__(name):   # plan imports pandas-heavy analysis code lazily
    if name in ("plan", "Protocol", "Solution"):
        import importlib
        plan = importlib.import_module(".plan", __name__)
        return plan if name == "plan" else getattr(plan, name)
    raise AttributeError(name)
