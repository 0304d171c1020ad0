"""erpl_monte_carlo_sim_b200 — B200-native batch engine behind the Python API of
smcconoughey/erpl_monte_carlo_sim (reference rocket_simulation/__init__.py:12-25).

Same class names and constructors; `FlightSimulator.simulate_flight` and
`MonteCarloAnalyzer.run_monte_carlo` run the 6-DOF integration of every sample in hand-written
sm_100a CUDA through a C ABI (include/emc.h).  No CPU fallback.
"""
from .environment import StandardAtmosphere, WindModel
from .monte_carlo import MonteCarloAnalyzer
from .motor import LiquidMotor, SolidMotor
from .rocket import Rocket
from .simulator import FlightSimulator
from .utils import *  # noqa: F401,F403

__version__ = "0.1.0"

__all__ = ["Rocket", "SolidMotor", "LiquidMotor", "StandardAtmosphere", "WindModel", "FlightSimulator",
           "MonteCarloAnalyzer"]
