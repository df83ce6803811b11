from . import kwave_if
from .kwave_if import run_simulation
from .sim_setup import SimSetup

__all__ = ["SimSetup", "run_simulation", "kwave_if"]
