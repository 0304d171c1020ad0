"""ctypes mirror of include/emc.h (structs, enums, field orders) and model packing.

Pure host-side plumbing: no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 3
MAX_CD_KNOTS = 16
MAX_CP_KNOTS = 16
MAX_THRUST_KNOTS = 32
MAX_WIND_KNOTS = 1024

MOTOR_LIQUID, MOTOR_SOLID = 0, 1

IN_FIELDS = ["x", "y", "z", "vx", "vy", "vz", "q0", "q1", "q2", "q3", "wx", "wy", "wz",
             "dry_mass", "prop_mass", "thrust_a", "nozzle_area", "mdot", "burn_time", "cd_scale"]
OUT_FIELDS = ["rail_exit_time", "rail_exit_x", "rail_exit_y", "rail_exit_z", "rail_exit_vx",
              "rail_exit_vy", "rail_exit_vz", "rail_exit_speed", "rail_exit_roll", "rail_exit_pitch",
              "rail_exit_yaw", "rail_exit_aoa", "rail_exit_sideslip", "wind_at_exit_u",
              "wind_at_exit_v", "wind_at_exit_w", "apogee_altitude", "apogee_time", "range",
              "flight_time", "final_x", "final_y", "final_z", "final_vx", "final_vy", "final_vz",
              "max_mach", "max_q", "max_speed", "max_abs_omega", "min_stability", "max_stability",
              "max_abs_aoa", "burnout_time", "chute_time"]
IOUT_FIELDS = ["n_steps", "termination", "apogee_index", "first_nan_step", "rail_steps"]
IN = {k: i for i, k in enumerate(IN_FIELDS)}
OUT = {k: i for i, k in enumerate(OUT_FIELDS)}
IOUT = {k: i for i, k in enumerate(IOUT_FIELDS)}
IN_COUNT, OUT_COUNT, IOUT_COUNT = len(IN_FIELDS), len(OUT_FIELDS), len(IOUT_FIELDS)
TAPE_WIDTH = 15
BTAPE_WIDTH = 4
SERIES_FIELDS = ["mass", "Ixx", "Iyy", "Izz", "center_of_mass", "euler_roll", "euler_pitch", "euler_yaw", "thrust", "drag",
                 "cd", "cl", "cm", "cp_location_dynamic", "stability_margin", "angle_of_attack", "sideslip_angle", "speed",
                 "mach", "dynamic_pressure"]
SER = {k: i for i, k in enumerate(SERIES_FIELDS)}
SERIES_COUNT = len(SERIES_FIELDS)

TERMINATION = {0: "none", 1: "ground_impact", 2: "excessive_altitude", 3: "coast_cap", 4: "max_time"}

STATUS = {0: "EMC_OK", -1: "EMC_ERR_INVALID", -2: "EMC_ERR_NO_DEVICE", -3: "EMC_ERR_CUDA",
          -4: "EMC_ERR_NO_MODEL", -5: "EMC_ERR_CAPACITY"}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class EmcModel(C.Structure):
    _fields_ = [
        ("center_of_mass_dry", C.c_double),
        ("Ixx_dry", C.c_double), ("Iyy_dry", C.c_double),
        ("diameter", C.c_double),
        ("reference_area", C.c_double), ("reference_diameter", C.c_double),
        ("fin_root_chord", C.c_double), ("fin_tip_chord", C.c_double), ("fin_span", C.c_double),
        ("fin_sweep_angle", C.c_double),
        ("cp_location", C.c_double),
        ("parachute_area", C.c_double), ("parachute_cd", C.c_double),
        ("parachute_deployment_altitude", C.c_double),
        ("power_off_drag_factor", C.c_double),
        ("n_cd", C.c_int32), ("n_cp", C.c_int32),
        ("cd_mach", C.c_double * MAX_CD_KNOTS), ("cd0", C.c_double * MAX_CD_KNOTS),
        ("cda", C.c_double * MAX_CD_KNOTS),
        ("cp_mach", C.c_double * MAX_CP_KNOTS), ("cp_shift", C.c_double * MAX_CP_KNOTS),
        ("motor_kind", C.c_int32), ("n_thrust", C.c_int32),
        ("thrust_time", C.c_double * MAX_THRUST_KNOTS), ("thrust_curve", C.c_double * MAX_THRUST_KNOTS),
        ("sea_level_pressure", C.c_double), ("sea_level_temperature", C.c_double),
        ("temperature_lapse_rate", C.c_double),
        ("gas_constant", C.c_double), ("gravity", C.c_double),
        ("troposphere_height", C.c_double), ("stratosphere_height", C.c_double),
        ("stratosphere_temp", C.c_double),
        ("max_time", C.c_double), ("dt_initial", C.c_double), ("pitch_damping", C.c_double),
        ("yaw_damping", C.c_double), ("rail_length", C.c_double),
        ("has_wind", C.c_int32), ("n_wind", C.c_int32),
        ("wind_altitudes", _dp),
        ("gamma", C.c_double),
    ]


class EmcInputs(C.Structure):
    _fields_ = [("scalars", C.c_void_p), ("ld", C.c_int64), ("wind", C.c_void_p),
                ("wind_sample_stride", C.c_int64)]


class EmcOutputs(C.Structure):
    _fields_ = [("out", C.c_void_p), ("iout", C.c_void_p), ("ld", C.c_int64)]


class EmcRunOpts(C.Structure):
    _fields_ = [("refill_threshold", C.c_int32), ("block_threads", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("nan_fast_forward", C.c_int32), ("cold_state_in_smem", C.c_int32),
                ("flags", C.c_int32)]


class EmcDispersion(C.Structure):
    _fields_ = [("base_pos", C.c_double * 3), ("base_vel", C.c_double * 3), ("base_att", C.c_double * 3), ("base_omega", C.c_double * 3),
                ("sigma_pos", C.c_double * 3), ("sigma_vel", C.c_double * 3), ("sigma_att", C.c_double * 3), ("sigma_omega", C.c_double * 3),
                ("mass_sigma", C.c_double), ("wind_speed_lo", C.c_double), ("wind_speed_hi", C.c_double),
                ("wind_dir_lo", C.c_double), ("wind_dir_hi", C.c_double),
                ("dry_mass", C.c_double), ("propellant_mass", C.c_double),
                ("thrust_vacuum", C.c_double), ("thrust_sea_level", C.c_double), ("mass_flow_rate", C.c_double),
                ("nozzle_exit_area", C.c_double), ("motor_propellant_mass", C.c_double), ("motor_burn_time", C.c_double),
                ("thrust_sigma", C.c_double), ("flow_sigma", C.c_double), ("burn_sigma", C.c_double),
                ("motor_kind", C.c_int32), ("wind_mode", C.c_int32), ("n_knots", C.c_int32), ("pad_", C.c_int32),
                ("shear", _dp), ("base_wind", _dp), ("rho", _dp), ("innov", _dp)]


class EmcCounters(C.Structure):
    _fields_ = [("rk4_steps", C.c_int64), ("replay_steps", C.c_int64), ("rail_steps", C.c_int64),
                ("refills", C.c_int64), ("kernel_launches", C.c_int64),
                ("rail_ms", C.c_double), ("flight_ms", C.c_double), ("tape_rows", C.c_int64), ("handovers", C.c_int64), ("parked", C.c_int64), ("strict_steps", C.c_int64),
                ("strict_ms", C.c_double), ("yielded", C.c_int64)]


_MODEL_SCALARS = ["center_of_mass_dry", "Ixx_dry", "Iyy_dry", "diameter", "reference_area",
                  "reference_diameter", "fin_root_chord", "fin_tip_chord", "fin_span", "fin_sweep_angle",
                  "cp_location", "parachute_area", "parachute_cd", "parachute_deployment_altitude",
                  "power_off_drag_factor", "sea_level_pressure", "sea_level_temperature",
                  "temperature_lapse_rate", "gas_constant", "gravity", "troposphere_height",
                  "stratosphere_height", "stratosphere_temp", "max_time", "dt_initial", "pitch_damping",
                  "yaw_damping", "rail_length"]


def _fill(arr, values, limit, what):
    v = np.asarray(values, dtype=np.float64).ravel()
    if v.size > limit:
        raise ValueError(f"{what}: {v.size} knots exceed the engine limit of {limit}")
    for i, x in enumerate(v):
        arr[i] = float(x)
    return v.size


def pack_model(md: dict):
    """dict (see marshal.model_dict) -> (EmcModel, keepalive).  keepalive owns the wind grid."""
    m = EmcModel()
    for k in _MODEL_SCALARS:
        setattr(m, k, float(md[k]))
    n_cd = _fill(m.cd_mach, md["cd_mach"], MAX_CD_KNOTS, "Cd_data['mach']")
    if _fill(m.cd0, md["cd0"], MAX_CD_KNOTS, "Cd_data['cd0']") != n_cd or \
            _fill(m.cda, md["cda"], MAX_CD_KNOTS, "Cd_data['cda']") != n_cd:
        raise ValueError("Cd_data columns differ in length")
    m.n_cd = n_cd
    n_cp = _fill(m.cp_mach, md["cp_mach"], MAX_CP_KNOTS, "CP_shift_data['mach']")
    if _fill(m.cp_shift, md["cp_shift"], MAX_CP_KNOTS, "CP_shift_data['cp_shift']") != n_cp:
        raise ValueError("CP_shift_data columns differ in length")
    m.n_cp = n_cp
    m.motor_kind = int(md["motor_kind"])
    n_t = _fill(m.thrust_time, md["thrust_time"], MAX_THRUST_KNOTS, "thrust_curve_time")
    if _fill(m.thrust_curve, md["thrust_curve"], MAX_THRUST_KNOTS, "thrust_curve_thrust") != n_t:
        raise ValueError("thrust curve columns differ in length")
    m.n_thrust = n_t
    m.has_wind = int(md["has_wind"])
    alts = np.ascontiguousarray(np.asarray(md["wind_altitudes"], dtype=np.float64).ravel())
    if alts.size > MAX_WIND_KNOTS:
        raise ValueError(f"altitude_profile: {alts.size} knots exceed the engine limit of {MAX_WIND_KNOTS}")
    m.n_wind = int(alts.size) if m.has_wind else 0
    m.wind_altitudes = alts.ctypes.data_as(_dp) if alts.size else None
    m.gamma = float(md.get("gamma", 1.4))
    return m, alts


def model_from_npz(z, prefix=""):
    """Rebuild the model dict stored by oracle/make_golden.py (keys '<prefix>model__<name>')."""
    pre = prefix + "model__"
    return {k[len(pre):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(pre)}


def inputs_struct(scalars: np.ndarray, wind, wind_shared=False):
    """scalars: float64 C-contiguous [IN_COUNT][ld]; wind: float64 C-contiguous [n][N][3] or None."""
    assert scalars.dtype == np.float64 and scalars.flags.c_contiguous and scalars.shape[0] == IN_COUNT
    s = EmcInputs()
    s.scalars = scalars.ctypes.data
    s.ld = scalars.shape[1]
    if wind is None or wind.size == 0:
        s.wind = None
        s.wind_sample_stride = 0
    else:
        assert wind.dtype == np.float64 and wind.flags.c_contiguous
        s.wind = wind.ctypes.data
        s.wind_sample_stride = 0 if wind_shared else wind.shape[-2] * 3
    return s


def outputs_alloc(n: int):
    out = np.empty((OUT_COUNT, n), np.float64)
    iout = np.empty((IOUT_COUNT, n), np.int32)
    s = EmcOutputs()
    s.out = out.ctypes.data
    s.iout = iout.ctypes.data
    s.ld = n
    return s, out, iout
