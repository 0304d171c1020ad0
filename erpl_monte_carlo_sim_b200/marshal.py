"""Marshalling of the parameter objects into the C-ABI blocks of include/emc.h.

Every attribute the hot path reads is taken from the live objects at call time, because users of the
reference mutate attributes between calls (SURVEY.md §5 "Config / flags").
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .motor import is_solid
from .utils import euler_to_quaternion

RAIL_LENGTH = 18.288      # default argument of the reference's _simulate_launch_rail (simulator.py:42)


def model_dict(rocket, motor, atmosphere, sim, altitude_profile=None) -> dict:
    solid = is_solid(motor)
    f = float
    return dict(
        center_of_mass_dry=f(rocket.center_of_mass_dry), Ixx_dry=f(rocket.Ixx_dry), Iyy_dry=f(rocket.Iyy_dry),
        diameter=f(rocket.diameter), reference_area=f(rocket.reference_area),
        reference_diameter=f(rocket.reference_diameter), fin_root_chord=f(rocket.fin_root_chord),
        fin_tip_chord=f(rocket.fin_tip_chord), fin_span=f(rocket.fin_span),
        fin_sweep_angle=f(rocket.fin_sweep_angle), cp_location=f(rocket.cp_location),
        parachute_area=f(rocket.parachute_area), parachute_cd=f(rocket.parachute_cd),
        parachute_deployment_altitude=f(rocket.parachute_deployment_altitude),
        power_off_drag_factor=f(rocket.power_off_drag_factor),
        cd_mach=np.asarray(rocket.Cd_data["mach"], np.float64), cd0=np.asarray(rocket.Cd_data["cd0"], np.float64),
        cda=np.asarray(rocket.Cd_data["cda"], np.float64),
        cp_mach=np.asarray(rocket.CP_shift_data["mach"], np.float64),
        cp_shift=np.asarray(rocket.CP_shift_data["cp_shift"], np.float64),
        motor_kind=_abi.MOTOR_SOLID if solid else _abi.MOTOR_LIQUID,
        thrust_time=np.asarray(motor.thrust_curve_time, np.float64) if solid else np.zeros(0),
        thrust_curve=np.asarray(motor.thrust_curve_thrust, np.float64) if solid else np.zeros(0),
        sea_level_pressure=f(atmosphere.sea_level_pressure), sea_level_temperature=f(atmosphere.sea_level_temperature),
        temperature_lapse_rate=f(atmosphere.temperature_lapse_rate), gas_constant=f(atmosphere.gas_constant),
        gravity=f(atmosphere.gravity), troposphere_height=f(atmosphere.troposphere_height),
        stratosphere_height=f(atmosphere.stratosphere_height), stratosphere_temp=f(atmosphere.stratosphere_temp),
        gamma=f(getattr(atmosphere, "gamma", 1.4)),
        max_time=f(sim.max_time), dt_initial=f(sim.dt_initial), pitch_damping=f(sim.pitch_damping),
        yaw_damping=f(sim.yaw_damping), rail_length=f(getattr(sim, "rail_length", RAIL_LENGTH)),
        has_wind=0 if altitude_profile is None else 1,
        wind_altitudes=np.zeros(0) if altitude_profile is None else np.ascontiguousarray(altitude_profile, np.float64),
    )


def initial_state_block(n, position, velocity, attitude, angular_velocity):
    """Rows X..WZ of the scalar block from (n,3) arrays (or broadcastable length-3 vectors)."""
    blk = np.empty((_abi.IN_COUNT, n), np.float64)
    pos = np.broadcast_to(np.asarray(position, np.float64), (n, 3))
    vel = np.broadcast_to(np.asarray(velocity, np.float64), (n, 3))
    att = np.broadcast_to(np.asarray(attitude, np.float64), (n, 3))
    om = np.broadcast_to(np.asarray(angular_velocity, np.float64), (n, 3))
    blk[_abi.IN["x"]:_abi.IN["z"] + 1] = pos.T
    blk[_abi.IN["vx"]:_abi.IN["vz"] + 1] = vel.T
    blk[_abi.IN["q0"]:_abi.IN["q3"] + 1] = euler_to_quaternion(att[:, 0], att[:, 1], att[:, 2]).T
    blk[_abi.IN["wx"]:_abi.IN["wz"] + 1] = om.T
    return blk


def single_sample_block(initial_conditions, rocket, motor):
    ic = initial_conditions
    blk = initial_state_block(1, ic.get("position", [0.0, 0.0, 0.0]), ic.get("velocity", [0.0, 0.0, 0.0]),
                              ic.get("attitude", [0.0, 0.0, 0.0]), ic.get("angular_velocity", [0.0, 0.0, 0.0]))
    IN = _abi.IN
    blk[IN["dry_mass"]] = rocket.dry_mass
    blk[IN["prop_mass"]] = rocket.propellant_mass
    blk[IN["thrust_a"]] = 1.0 if is_solid(motor) else motor.thrust_vacuum
    blk[IN["nozzle_area"]] = motor.nozzle_exit_area
    blk[IN["mdot"]] = motor.mass_flow_rate
    blk[IN["burn_time"]] = motor.burn_time
    blk[IN["cd_scale"]] = 1.0
    return blk
