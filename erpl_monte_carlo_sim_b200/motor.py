"""Motor parameter holders — constructor-compatible with the reference's SolidMotor / LiquidMotor
(motor.py:8-52, 128-150) including the Monte Carlo perturbation constructors (motor.py:95-125,
171-186), which are input generation.  Thrust and mass-flow evaluation (motor.py:54-93, 152-169)
runs on the GPU.
"""
from __future__ import annotations

import numpy as np

LBF = 4.44822


class SolidMotor:
    def __init__(self, name="Solid Motor"):
        self.name = name
        self.total_impulse = 156297
        self.burn_time = 15.0
        self.propellant_mass = 63.5
        self.average_thrust = self.total_impulse / self.burn_time
        self.thrust_sea_level = 2290 * LBF
        self.thrust_vacuum = 2590 * LBF
        self.nozzle_exit_area = (self.thrust_vacuum - self.thrust_sea_level) / 101325.0
        self.thrust_curve_time = np.array([0.0, 0.2, 0.5, 1.0, 2.0, 5.0, 8.0, 12.0, 14.0, 15.0])
        self.thrust_curve_normalized = np.array([0.0, 2.2, 2.0, 1.8, 1.5, 1.2, 1.0, 0.8, 0.3, 0.0])
        self.thrust_curve_thrust = self.thrust_curve_normalized * self.average_thrust
        self.mass_flow_rate = 4.26
        self.exhaust_velocity = self.average_thrust / self.mass_flow_rate
        self.thrust_uncertainty = 0.05
        self.burn_time_uncertainty = 0.02
        self.total_impulse_uncertainty = 0.03

    def perturb_for_monte_carlo(self, random_state=None):
        rs = random_state if random_state is not None else np.random.RandomState()
        out = SolidMotor(self.name + "_perturbed")
        k = rs.normal(1.0, self.thrust_uncertainty)
        out.thrust_curve_thrust = self.thrust_curve_thrust * k
        out.average_thrust = self.average_thrust * k
        out.thrust_sea_level = self.thrust_sea_level * k
        out.thrust_vacuum = self.thrust_vacuum * k
        out.burn_time = self.burn_time * rs.normal(1.0, self.burn_time_uncertainty)
        out.total_impulse = self.total_impulse * rs.normal(1.0, self.total_impulse_uncertainty)
        out.mass_flow_rate = 4.26 * k
        out.exhaust_velocity = out.average_thrust / out.mass_flow_rate
        out.nozzle_exit_area = self.nozzle_exit_area * k
        out._thrust_multiplier = k
        return out

    # model-evaluation helpers (motor.py:54-93 / 152-169 of the reference); thrust is evaluated on the GPU by the engine
    def get_thrust(self, time, ambient_pressure=None):
        from .simulator import _scalar_or_array, evaluate_component
        p = 101325.0 if ambient_pressure is None else ambient_pressure
        a = 1.0 if is_solid(self) else self.thrust_vacuum
        return _scalar_or_array(evaluate_component(3, (time, p, a, self.nozzle_exit_area, self.burn_time), motor=self)[0], time)

    def get_mass_flow_rate(self, time):
        return 0.0 if (time < 0 or time > self.burn_time) else self.mass_flow_rate

    def get_propellant_remaining(self, time):
        if time <= 0:
            return 1.0
        if time >= self.burn_time:
            return 0.0
        return max(0.0, 1.0 - time / self.burn_time)


class LiquidMotor:
    def __init__(self, name="Liquid Motor", thrust_vacuum=2590 * LBF, thrust_sea_level=2290 * LBF,
                 mass_flow_rate=4.26, propellant_mass=63.5):
        self.name = name
        self.thrust_vacuum = thrust_vacuum
        self.thrust_sea_level = thrust_sea_level
        self.mass_flow_rate = mass_flow_rate
        self.propellant_mass = propellant_mass
        self.nozzle_exit_area = (self.thrust_vacuum - self.thrust_sea_level) / 101325.0
        self.burn_time = self.propellant_mass / self.mass_flow_rate
        self.total_impulse = self.thrust_vacuum * self.burn_time
        self.thrust_uncertainty = 0.05
        self.mass_flow_uncertainty = 0.03

    def perturb_for_monte_carlo(self, random_state=None):
        rs = random_state if random_state is not None else np.random.RandomState()
        k_thrust = rs.normal(1.0, self.thrust_uncertainty)
        k_flow = rs.normal(1.0, self.mass_flow_uncertainty)
        return LiquidMotor(self.name + "_perturbed", thrust_vacuum=self.thrust_vacuum * k_thrust,
                           thrust_sea_level=self.thrust_sea_level * k_thrust,
                           mass_flow_rate=self.mass_flow_rate * k_flow, propellant_mass=self.propellant_mass)

    # model-evaluation helpers (motor.py:54-93 / 152-169 of the reference); thrust is evaluated on the GPU by the engine
    def get_thrust(self, time, ambient_pressure=None):
        from .simulator import _scalar_or_array, evaluate_component
        p = 101325.0 if ambient_pressure is None else ambient_pressure
        a = 1.0 if is_solid(self) else self.thrust_vacuum
        return _scalar_or_array(evaluate_component(3, (time, p, a, self.nozzle_exit_area, self.burn_time), motor=self)[0], time)

    def get_mass_flow_rate(self, time):
        return 0.0 if (time < 0 or time > self.burn_time) else self.mass_flow_rate

    def get_propellant_remaining(self, time):
        if time <= 0:
            return 1.0
        if time >= self.burn_time:
            return 0.0
        return max(0.0, 1.0 - time / self.burn_time)


def is_solid(motor) -> bool:
    return hasattr(motor, "thrust_curve_time")
