"""
TEST INFRASTRUCTURE — not part of the product path.

Live-reference harness: imports the unmodified Python reference from /root/reference (this
container only; the GPU box does not have it) and drives it to produce golden vectors.

Nothing under erpl_monte_carlo_sim_b200/ may import this module.  Only oracle/make_golden.py and
`tests/` (container-only cross-checks, skipped when /root/reference is absent) use it.

Reference entry points driven here (all unmodified):
  rocket_simulation/simulator.py:127   FlightSimulator.simulate_flight
  rocket_simulation/simulator.py:295   FlightSimulator._rocket_dynamics
  rocket_simulation/monte_carlo.py:156 MonteCarloAnalyzer._generate_parameter_samples
  rocket_simulation/monte_carlo.py:225 MonteCarloAnalyzer._run_single_simulation
  rocket_simulation/monte_carlo.py:400 MonteCarloAnalyzer._analyze_results
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("EMC_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "rocket_simulation")

# Field order of the per-sample scalar block; must equal enum emc_in_field in include/emc.h.
IN_FIELDS = ["x", "y", "z", "vx", "vy", "vz", "q0", "q1", "q2", "q3", "wx", "wy", "wz",
             "dry_mass", "prop_mass", "thrust_a", "nozzle_area", "mdot", "burn_time", "cd_scale"]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "simulator.py"))


def _install_matplotlib_stub():
    """monte_carlo.py:6 imports matplotlib.pyplot at module top; it is absent in this image."""
    try:
        import matplotlib.pyplot  # noqa: F401
        return
    except Exception:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


_REF = None


def ref():
    """Import the reference modules (flat imports, so its directory goes on sys.path)."""
    global _REF
    if _REF is not None:
        return _REF
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_PKG}")
    _install_matplotlib_stub()
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    import environment
    import monte_carlo
    import motor
    import rocket
    import simulator
    import utils
    _REF = types.SimpleNamespace(utils=utils, rocket=rocket, motor=motor, environment=environment,
                                 simulator=simulator, monte_carlo=monte_carlo)
    return _REF


@contextlib.contextmanager
def quiet():
    """The reference prints 4 lines per simulate_flight (simulator.py:142-147)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


class _ChuteProbeMixin:
    """Records the stage time of the derivative call that latched the sticky parachute flag
    (simulator.py:366-369).  Observation only: calls the reference's own _rocket_dynamics."""

    def _rocket_dynamics(self, t, state):
        before = self.parachute_deployed
        out = super()._rocket_dynamics(t, state)
        if (not before) and self.parachute_deployed and getattr(self, "_chute_time", None) is None:
            self._chute_time = float(t)
        return out


def make_probe_simulator_class():
    R = ref()

    class ProbeSim(_ChuteProbeMixin, R.simulator.FlightSimulator):
        def simulate_flight(self, *a, **k):
            self._chute_time = None
            res = super().simulate_flight(*a, **k)
            res["_chute_time"] = np.nan if self._chute_time is None else self._chute_time
            res["_burn_time"] = float(self.motor.burn_time)
            return res

    return ProbeSim


# ----------------------------------------------------------------------------------------------
# marshalling reference objects -> the C-ABI input block (explicit, independent of the package)
# ----------------------------------------------------------------------------------------------
def model_dict(rocket, motor, atmosphere, sim, altitude_profile=None):
    """Run-constant attributes read by the hot path, as a plain dict (see emc_model in emc.h)."""
    solid = hasattr(motor, "thrust_curve_time")
    return dict(
        center_of_mass_dry=float(rocket.center_of_mass_dry),
        Ixx_dry=float(rocket.Ixx_dry), Iyy_dry=float(rocket.Iyy_dry),
        diameter=float(rocket.diameter), reference_area=float(rocket.reference_area),
        reference_diameter=float(rocket.reference_diameter),
        fin_root_chord=float(rocket.fin_root_chord), fin_tip_chord=float(rocket.fin_tip_chord),
        fin_span=float(rocket.fin_span), fin_sweep_angle=float(rocket.fin_sweep_angle),
        cp_location=float(rocket.cp_location), parachute_area=float(rocket.parachute_area),
        parachute_cd=float(rocket.parachute_cd),
        parachute_deployment_altitude=float(rocket.parachute_deployment_altitude),
        power_off_drag_factor=float(rocket.power_off_drag_factor),
        cd_mach=np.asarray(rocket.Cd_data["mach"], float), cd0=np.asarray(rocket.Cd_data["cd0"], float),
        cda=np.asarray(rocket.Cd_data["cda"], float),
        cp_mach=np.asarray(rocket.CP_shift_data["mach"], float),
        cp_shift=np.asarray(rocket.CP_shift_data["cp_shift"], float),
        motor_kind=1 if solid else 0,
        thrust_time=np.asarray(motor.thrust_curve_time, float) if solid else np.zeros(0),
        thrust_curve=np.asarray(motor.thrust_curve_thrust, float) if solid else np.zeros(0),
        sea_level_pressure=float(atmosphere.sea_level_pressure),
        sea_level_temperature=float(atmosphere.sea_level_temperature),
        temperature_lapse_rate=float(atmosphere.temperature_lapse_rate),
        gas_constant=float(atmosphere.gas_constant), gravity=float(atmosphere.gravity),
        troposphere_height=float(atmosphere.troposphere_height),
        stratosphere_height=float(atmosphere.stratosphere_height),
        stratosphere_temp=float(atmosphere.stratosphere_temp),
        max_time=float(sim.max_time), dt_initial=float(sim.dt_initial),
        pitch_damping=float(sim.pitch_damping), yaw_damping=float(sim.yaw_damping),
        rail_length=18.288,
        has_wind=0 if altitude_profile is None else 1,
        wind_altitudes=np.zeros(0) if altitude_profile is None else np.asarray(altitude_profile, float),
    )


def sample_scalars(ic, rocket, motor, cd_scale=1.0):
    """One column of the scalar block for a (possibly perturbed) rocket/motor pair.
    Solid: THRUST_A is the multiplier on the base curve (1.0 for an unperturbed motor;
    mc_sample_setup records the drawn multiplier on the perturbed motor object)."""
    R = ref()
    att = ic.get("attitude", [0.0, 0.0, 0.0])
    q = R.utils.euler_to_quaternion(att[0], att[1], att[2])
    pos = np.zeros(3); pos[:] = ic.get("position", [0.0, 0.0, 0.0])
    vel = np.zeros(3); vel[:] = ic.get("velocity", [0.0, 0.0, 0.0])
    om = np.zeros(3); om[:] = ic.get("angular_velocity", [0.0, 0.0, 0.0])
    solid = hasattr(motor, "thrust_curve_time")
    if solid:
        thrust_a = float(getattr(motor, "_emc_thrust_multiplier", 1.0))
    else:
        thrust_a = float(motor.thrust_vacuum)
    col = np.array([pos[0], pos[1], pos[2], vel[0], vel[1], vel[2], q[0], q[1], q[2], q[3],
                    om[0], om[1], om[2], rocket.dry_mass, rocket.propellant_mass, thrust_a,
                    motor.nozzle_exit_area, motor.mass_flow_rate, motor.burn_time, cd_scale], float)
    return col


def mc_sample_setup(mc, base_ic, params, planar=False):
    """Follows monte_carlo.py:225-288 with the reference's own helpers and returns
    (ic, rocket, motor, atmosphere, wind_profile, altitude_profile).  `planar` applies the W-B
    projection of SURVEY.md §8d (beta == 0 exactly)."""
    R = ref()
    p = dict(params)
    if planar:
        p = {k: (np.array(v, float) if isinstance(v, np.ndarray) else v) for k, v in p.items()}
        p["initial_velocity_offset"][1] = 0.0
        p["initial_attitude_offset"][0] = 0.0
        p["initial_attitude_offset"][2] = 0.0
        p["initial_angular_velocity_offset"][0] = 0.0
        p["initial_angular_velocity_offset"][2] = 0.0
        p["initial_position_offset"][1] = 0.0
        p["wind_direction"] = 0.0 if (p["random_seed"] % 2 == 0) else float(np.pi)
    ic = base_ic.copy()
    for key, off in (("position", "initial_position_offset"), ("velocity", "initial_velocity_offset"),
                     ("attitude", "initial_attitude_offset"),
                     ("angular_velocity", "initial_angular_velocity_offset")):
        if key in ic:
            ic[key] = np.array(ic[key]) + p[off]
        else:
            ic[key] = p[off]
    rocket = mc._perturb_rocket(p)
    seed = p["random_seed"]
    motor = mc._perturb_motor(p)
    # recover the Solid thrust multiplier the motor drew (first normal of RandomState(seed), motor.py:104)
    if hasattr(motor, "thrust_curve_time"):
        motor._emc_thrust_multiplier = np.random.RandomState(seed).normal(1.0, mc.motor.thrust_uncertainty)
    motor.propellant_mass = rocket.propellant_mass
    if hasattr(motor, "mass_flow_rate") and motor.mass_flow_rate > 0:
        motor.burn_time = motor.propellant_mass / motor.mass_flow_rate
    atmosphere = mc._perturb_atmosphere(p)
    if mc.base_wind_profile is not None and mc.base_altitude_profile is not None:
        altitude_profile = mc.base_altitude_profile
        wind_profile = mc.wind_model.perturb_wind_profile(
            altitude_profile, mc.base_wind_profile, random_state=np.random.RandomState(seed))
        wind_profile[:, 0] += p["wind_speed"] * np.cos(p["wind_direction"])
        wind_profile[:, 1] += p["wind_speed"] * np.sin(p["wind_direction"])
    else:
        altitude_profile = np.linspace(0, 25000, 100)
        wind_profile = mc.wind_model.generate_stochastic_profile(
            altitude_profile, p["wind_speed"], p["wind_direction"],
            random_state=np.random.RandomState(seed))
    if planar:
        wind_profile[:, 1] = 0.0
    return ic, rocket, motor, atmosphere, wind_profile, altitude_profile


# ----------------------------------------------------------------------------------------------
# summaries from a reference result dict (layout of enum emc_out_field / emc_iout_field)
# ----------------------------------------------------------------------------------------------
OUT_FIELDS = ["rail_exit_time", "rail_exit_x", "rail_exit_y", "rail_exit_z", "rail_exit_vx",
              "rail_exit_vy", "rail_exit_vz", "rail_exit_speed", "rail_exit_roll", "rail_exit_pitch",
              "rail_exit_yaw", "rail_exit_aoa", "rail_exit_sideslip", "wind_at_exit_u",
              "wind_at_exit_v", "wind_at_exit_w", "apogee_altitude", "apogee_time", "range",
              "flight_time", "final_x", "final_y", "final_z", "final_vx", "final_vy", "final_vz",
              "max_mach", "max_q", "max_speed", "max_abs_omega", "min_stability", "max_stability",
              "max_abs_aoa", "burnout_time", "chute_time"]
IOUT_FIELDS = ["n_steps", "termination", "apogee_index", "first_nan_step", "rail_steps"]


def summarize(res, sim, wind_profile, altitude_profile):
    """Summary vector of one reference flight.  Mach and dynamic pressure are not result keys; they
    are recomputed per stored state with the reference's own functions exactly as
    _extract_results does (simulator.py:522-532,541)."""
    R = ref()
    pos, vel = res["position"], res["velocity"]
    n = pos.shape[1]
    mach = np.empty(n)
    qdyn = np.empty(n)
    for i in range(n):
        alt = pos[2, i]
        atm = sim.atmosphere.get_properties(alt)
        if wind_profile is not None and altitude_profile is not None:
            w = sim.wind_model.get_wind_at_altitude(alt, wind_profile, altitude_profile)
        else:
            w = np.zeros(3)
        vr = vel[:, i] - w
        mach[i] = R.utils.mach_number(vr, atm["temperature"])
        qdyn[i] = 0.5 * atm["density"] * np.linalg.norm(vr) ** 2
    time = res["time"]
    burn_time = res.get("_burn_time", sim.motor.burn_time)
    bidx = int(np.argmax(time > burn_time))
    out = np.array([
        res["rail_exit_time"], *res["rail_exit_position"], *res["rail_exit_velocity"],
        res["rail_exit_speed"], *res["rail_exit_euler"], res["rail_exit_angle_of_attack"],
        res["rail_exit_sideslip"], *res["wind_at_exit"], res["apogee_altitude"], res["apogee_time"],
        res["range"], res["flight_time"], *pos[:, -1], *vel[:, -1],
        np.max(mach), np.max(qdyn), np.max(res["speed"]), np.max(np.abs(res["angular_velocity"])),
        np.min(res["stability_margin"]), np.max(res["stability_margin"]),
        np.max(np.abs(res["angle_of_attack"])), time[bidx], res.get("_chute_time", np.nan)], float)
    alt = pos[2]
    z_f, vz_f = alt[-1], vel[2, -1]
    t_last = time[-1] + res["rail_exit_time"]
    if z_f <= 0.5 and vz_f <= 0:
        term = 1
    elif z_f > 100000.0:
        term = 2
    elif not (t_last < sim.max_time):
        # the coast cap can coincide with max_time only in theory; the loop guard wins if t>=max_time
        term = 4
    else:
        term = 3
    nan_idx = np.flatnonzero(np.isnan(alt))
    iout = np.array([n - 1, term, int(np.argmax(alt)), int(nan_idx[0]) if nan_idx.size else -1,
                     int(round(res["rail_exit_time"] / sim.dt_initial))], np.int32)
    return out, iout
