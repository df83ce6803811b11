"""Planning API that consumes the solver (mirrors /root/reference/src/openlifu/plan/__init__.py):
``Protocol.calc_solution`` -> per-focus beamform + ``run_simulation`` -> ``Solution`` -> ``SolutionAnalysis``.
"""
from __future__ import annotations

from . import solution_analysis
from .param_constraint import PARAM_STATUS_SYMBOLS, ParameterConstraint
from .protocol import OnPulseMismatchAction, Protocol
from .solution import Solution
from .solution_analysis import SolutionAnalysis, SolutionAnalysisOptions
from .target_constraints import TargetConstraints

__all__ = ["Protocol", "OnPulseMismatchAction", "Solution", "SolutionAnalysis", "SolutionAnalysisOptions",
           "ParameterConstraint", "PARAM_STATUS_SYMBOLS", "TargetConstraints", "solution_analysis"]
