"""Atmosphere and wind parameter holders — constructor-compatible with the reference's
StandardAtmosphere (environment.py:11-24) and WindModel (environment.py:113-116).

The atmosphere profile, gravity and the wind-table interpolation (environment.py:26-108, 267-276)
are evaluated on the GPU.  What stays on the host is input generation: the CSV loader
(environment.py:202-216) and the stochastic wind-profile generators (environment.py:125-200,
218-265), here with batched variants that reproduce the reference's NumPy draws sample for sample.
"""
from __future__ import annotations

import numpy as np


class StandardAtmosphere:
    def __init__(self):
        self.sea_level_pressure = 101325.0
        self.sea_level_temperature = 288.15
        self.sea_level_density = 1.225
        self.temperature_lapse_rate = 0.0065
        self.gas_constant = 287.053
        self.gravity = 9.80665
        self.gamma = 1.4
        self.troposphere_height = 11000.0
        self.stratosphere_height = 20000.0
        self.stratosphere_temp = 216.65

    # model-evaluation helpers (environment.py:26-108 of the reference): evaluated on the GPU by the engine
    def get_properties(self, altitude):
        from .simulator import _scalar_or_array, evaluate_component
        o = evaluate_component(0, (altitude,), atmosphere=self)
        return {k: _scalar_or_array(o[i], altitude) for i, k in enumerate(("temperature", "pressure", "density", "speed_of_sound"))}

    def get_gravity(self, altitude):
        from .simulator import _scalar_or_array, evaluate_component
        return _scalar_or_array(evaluate_component(0, (altitude,), atmosphere=self)[4], altitude)


def _ar1_coefficients(wind_model, altitudes):
    """Per-knot turbulence scale, AR(1) correlation and innovation scale of environment.py:161-185 /
    242-252, evaluated knot by knot with scalar NumPy calls exactly as the reference does."""
    n = len(altitudes)
    scale = np.empty(n)
    rho = np.zeros(n)
    innov = np.empty(n)
    for i in range(n):
        scale[i] = wind_model.turbulence_intensity * np.exp(-altitudes[i] / 2000.0)
        if i == 0:
            innov[i] = scale[i]
            continue
        dz = max(altitudes[i] - altitudes[i - 1], 1e-6)
        c = np.clip(np.exp(-dz / wind_model.correlation_length), 0.1, 0.95)
        rho[i] = c
        innov[i] = scale[i] * np.sqrt(max(1 - c ** 2, 0.01))
    return scale, rho, innov


class WindModel:
    def __init__(self):
        self.power_law_exponent = 0.14
        self.turbulence_intensity = 2.0
        self.correlation_length = 100.0

    def get_wind_at_altitude(self, altitude, wind_profile, altitude_profile):
        """Three np.interp calls on the columns of the (N,3) table (environment.py:267-276): plain host table lookup."""
        wind_profile = np.asarray(wind_profile, float)
        if len(wind_profile) == 0:
            return np.array([0.0, 0.0, 0.0])
        return np.array([np.interp(altitude, altitude_profile, wind_profile[:, k]) for k in range(3)])

    def power_law_profile(self, altitude, reference_wind_speed, reference_altitude=10.0):
        return reference_wind_speed * (altitude / reference_altitude) ** self.power_law_exponent

    def load_wind_profile_from_csv(self, file_path):
        data = np.genfromtxt(file_path, delimiter=",", names=True)
        altitudes = data["altitude"]
        w = data["w"] if "w" in data.dtype.names else np.zeros_like(altitudes)
        return altitudes, np.vstack([data["u"], data["v"], w]).T

    # -- single-profile generators (same call signatures and draw order as the reference) --------
    def generate_stochastic_profile(self, altitudes, base_wind_speed, base_wind_direction=None, random_state=None):
        rs = random_state if random_state is not None else np.random.RandomState()
        if base_wind_direction is None:
            base_wind_direction = rs.uniform(0.0, 2 * np.pi)
        altitudes = np.asarray(altitudes, dtype=np.float64)
        g = rs.standard_normal(3 * len(altitudes)).reshape(-1, 3)
        return self.stochastic_profiles_batch(altitudes, np.array([base_wind_speed], float),
                                              np.array([base_wind_direction], float), g[None])[0]

    def perturb_wind_profile(self, altitudes, base_profile, random_state=None):
        rs = random_state if random_state is not None else np.random.RandomState()
        altitudes = np.asarray(altitudes, dtype=np.float64)
        g = rs.standard_normal(3 * len(altitudes)).reshape(-1, 3)
        return self.perturbed_profiles_batch(altitudes, np.asarray(base_profile, float), g[None])[0]

    # -- batched generators: gauss[s, i, k] is the (3*i + k)-th standard normal of sample s' stream ---
    def stochastic_profiles_batch(self, altitudes, speeds, directions, gauss):
        """environment.py:125-200 for many samples at once -> (S, N, 3)."""
        altitudes = np.asarray(altitudes, dtype=np.float64)
        n = len(altitudes)
        scale, rho, innov = _ar1_coefficients(self, altitudes)
        shear = np.array([(altitudes[i] / 10.0) ** self.power_law_exponent for i in range(n)])
        cos_d, sin_d = np.cos(directions), np.sin(directions)
        out = np.zeros((len(speeds), n, 3))
        mean_u_prev = mean_v_prev = None
        for i in range(n):
            level = speeds * shear[i]                       # power_law_profile(altitude, speed)
            mean_u, mean_v = level * cos_d, level * sin_d
            if i == 0:
                out[:, 0, 0] = mean_u + (0 + innov[0] * gauss[:, 0, 0])
                out[:, 0, 1] = mean_v + (0 + innov[0] * gauss[:, 0, 1])
                out[:, 0, 2] = 0 + (scale[0] * 0.3) * gauss[:, 0, 2]
            else:
                tu = rho[i] * (out[:, i - 1, 0] - mean_u_prev) + (0 + innov[i] * gauss[:, i, 0])
                tv = rho[i] * (out[:, i - 1, 1] - mean_v_prev) + (0 + innov[i] * gauss[:, i, 1])
                tw = rho[i] * out[:, i - 1, 2] + (0 + (innov[i] * 0.3) * gauss[:, i, 2])
                out[:, i, 0] = mean_u + tu
                out[:, i, 1] = mean_v + tv
                out[:, i, 2] = tw
            mean_u_prev, mean_v_prev = mean_u, mean_v
        return out

    def perturbed_profiles_batch(self, altitudes, base_profile, gauss):
        """environment.py:218-265 for many samples at once -> (S, N, 3)."""
        altitudes = np.asarray(altitudes, dtype=np.float64)
        base = np.asarray(base_profile, dtype=np.float64)
        n = len(altitudes)
        scale, rho, innov = _ar1_coefficients(self, altitudes)
        out = np.zeros((gauss.shape[0], n, 3))
        for i in range(n):
            for k in range(3):
                s = innov[i] * 0.3 if k == 2 else innov[i]
                if i == 0:
                    s = scale[0] * 0.3 if k == 2 else scale[0]
                    out[:, 0, k] = base[0, k] + (0 + s * gauss[:, 0, k])
                else:
                    turb = rho[i] * (out[:, i - 1, k] - base[i - 1, k]) + (0 + s * gauss[:, i, k])
                    out[:, i, k] = base[i, k] + turb
        return out
