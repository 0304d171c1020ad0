"""FlightSimulator — drop-in for the reference's class (simulator.py:9-40, 127-293).

`simulate_flight` keeps its signature and result keys; the rail phase, the RK4 loop, the derivative
and the event logic run in the CUDA engine (csrc/) through the C ABI.  There is no CPU integration
path: without a B200 the call raises.
"""
from __future__ import annotations

import numpy as np

from . import _abi, _lib, marshal
from .utils import object_to_serializable_dict

_ENGINES = {}
_COMPONENT_ENGINES = {}


def get_engine(device: int = 0) -> "_lib.Engine":
    """One engine (emc_ctx) per device, shared by the Python front ends of this process."""
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _lib.Engine(device)
        _ENGINES[device] = eng
    return eng


def evaluate_component(comp, columns, rocket=None, motor=None, atmosphere=None, device=0):
    """Model-evaluation helpers of the parameter classes (get_properties, get_thrust, ...): evaluated by the engine's
    own device functions (emc_component_debug), never by a host re-implementation."""
    from .environment import StandardAtmosphere
    from .motor import LiquidMotor
    from .rocket import Rocket
    knobs = type("Knobs", (), dict(max_time=300.0, dt_initial=0.01, pitch_damping=20.0, yaw_damping=20.0))()
    # a context of its own: the helpers must not replace the model (or touch the resident batch) of the engine that the
    # simulator / analyzer of this process fly with (round-1 advisor finding); contexts of one device keep their own model
    eng = _COMPONENT_ENGINES.get(device)
    if eng is None:
        eng = _COMPONENT_ENGINES[device] = _lib.Engine(device)
    eng.set_model(marshal.model_dict(rocket or Rocket(), motor or LiquidMotor(), atmosphere or StandardAtmosphere(), knobs, None))
    return eng.component(comp, *columns)


def _scalar_or_array(v, like):
    return float(v[0]) if np.ndim(like) == 0 else v


class FlightSimulator:
    def __init__(self, rocket, motor, atmosphere, wind_model, device: int = 0):
        self.rocket = rocket
        self.motor = motor
        self.atmosphere = atmosphere
        self.wind_model = wind_model
        self.max_time = 300.0
        self.dt_initial = 0.01
        self.rtol = 1e-4            # unused by the reference's fixed-step integrator; kept as attributes
        self.atol = 1e-7
        self.ground_altitude = 0.0
        self.apogee_detected = False
        self.wind_profile = None
        self.altitude_profile = None
        self.pitch_damping = 20.0
        self.yaw_damping = 20.0
        self.parachute_deployed = False
        self.device = device
        self.verbose = False

    def simulate_flight(self, initial_conditions, wind_profile=None, altitude_profile=None):
        eng = get_engine(self.device)
        use_wind = wind_profile is not None and altitude_profile is not None
        if use_wind and len(wind_profile) == 0:
            use_wind = False                                   # environment.py:269-270
        self.wind_profile, self.altitude_profile = wind_profile, altitude_profile
        md = marshal.model_dict(self.rocket, self.motor, self.atmosphere, self, altitude_profile if use_wind else None)
        eng.set_model(md)
        blk = marshal.single_sample_block(initial_conditions, self.rocket, self.motor)
        wind = np.ascontiguousarray(wind_profile, np.float64) if use_wind else None
        cap = int(np.ceil(max(self.max_time, 0.0) / min(self.dt_initial, 0.005))) + 16
        out, iout, tape = eng.run_tape(blk, wind, cap=cap)
        ser = eng.extract_series(blk, wind, tape)                # simulator.py:496-552 on the device
        SR = _abi.SER
        o = out[:, 0]
        O = _abi.OUT
        rail_time = o[O["rail_exit_time"]]
        states = tape[:, 1:].T                                  # (14, n)
        quat = states[6:10]
        res = {
            "time": tape[:, 0] - rail_time,
            "position": states[0:3], "velocity": states[3:6], "quaternion": quat,
            "angular_velocity": states[10:13], "propellant_fraction": states[13],
            "mass": ser[SR["mass"]], "moments_of_inertia": ser[SR["Ixx"]:SR["Izz"] + 1],
            "altitude": states[2], "speed": ser[SR["speed"]],
            "euler_angles": ser[SR["euler_roll"]:SR["euler_yaw"] + 1],
            "center_of_mass": ser[SR["center_of_mass"]], "thrust": ser[SR["thrust"]], "drag": ser[SR["drag"]],
            "cd": ser[SR["cd"]], "cl": ser[SR["cl"]], "cm": ser[SR["cm"]],
            "cp_location_dynamic": ser[SR["cp_location_dynamic"]], "stability_margin": ser[SR["stability_margin"]],
            "angle_of_attack": ser[SR["angle_of_attack"]], "sideslip_angle": ser[SR["sideslip_angle"]],
            "mach": ser[SR["mach"]], "dynamic_pressure": ser[SR["dynamic_pressure"]],
            "cp_location": self.rocket.cp_location,
            "thrust_curve_time": getattr(self.motor, "thrust_curve_time", None),
            "thrust_curve_thrust": getattr(self.motor, "thrust_curve_thrust", None),
            "apogee_time": o[O["apogee_time"]], "apogee_altitude": o[O["apogee_altitude"]],
            "range": o[O["range"]], "flight_time": o[O["flight_time"]],
            "rail_exit_time": rail_time,
            "rail_exit_position": o[O["rail_exit_x"]:O["rail_exit_z"] + 1].copy(),
            "rail_exit_velocity": o[O["rail_exit_vx"]:O["rail_exit_vz"] + 1].copy(),
            "rail_exit_speed": float(o[O["rail_exit_speed"]]),
            "rail_exit_euler": o[O["rail_exit_roll"]:O["rail_exit_yaw"] + 1].copy(),
            "rail_exit_angle_of_attack": o[O["rail_exit_aoa"]],
            "rail_exit_sideslip": o[O["rail_exit_sideslip"]],
            "wind_at_exit": o[O["wind_at_exit_u"]:O["wind_at_exit_w"] + 1].copy(),
        }
        res.update(summary_extras(out, iout, 0))
        ic = initial_conditions
        res["initial_conditions"] = {
            "position": [float(v) for v in blk[0:3, 0]], "velocity": [float(v) for v in blk[3:6, 0]],
            "attitude": ic.get("attitude", [0.0, 0.0, 0.0]),
            "angular_velocity": [float(v) for v in blk[10:13, 0]],
        }
        res["rocket_parameters"] = object_to_serializable_dict(self.rocket)
        res["motor_parameters"] = object_to_serializable_dict(self.motor)
        res["simulation_assumptions"] = {"max_time": self.max_time, "dt_initial": self.dt_initial, "rtol": self.rtol,
                                         "atol": self.atol, "rail_length": marshal.RAIL_LENGTH}
        if wind_profile is not None and altitude_profile is not None:
            res["wind_profile"] = wind_profile
            res["altitude_profile"] = altitude_profile
        self.parachute_deployed = bool(np.isfinite(o[O["chute_time"]]))
        return res


def summary_extras(out, iout, i):
    """Engine-side per-sample diagnostics that the reference leaves to post-processing scripts
    (analyze_outlier.py:18-25): returned under their own keys next to the reference's."""
    O, I = _abi.OUT, _abi.IOUT
    return {
        "max_mach": out[O["max_mach"], i], "max_dynamic_pressure": out[O["max_q"], i],
        "max_speed": out[O["max_speed"], i], "max_abs_angular_velocity": out[O["max_abs_omega"], i],
        "min_stability_margin": out[O["min_stability"], i], "max_stability_margin": out[O["max_stability"], i],
        "max_abs_angle_of_attack": out[O["max_abs_aoa"], i], "burnout_time": out[O["burnout_time"], i],
        "parachute_deploy_time": out[O["chute_time"], i],
        "landing_position": out[O["final_x"]:O["final_z"] + 1, i].copy(),
        "landing_velocity": out[O["final_vx"]:O["final_vz"] + 1, i].copy(),
        "n_steps": int(iout[I["n_steps"], i]), "termination": _abi.TERMINATION[int(iout[I["termination"], i])],
    }
