"""Rocket parameter holder — constructor-compatible with the reference's Rocket (rocket.py:11-66).

It carries the attributes the engine marshals into `emc_model` / the per-sample input block; the
mass-property and aerodynamic-coefficient evaluation (rocket.py:105-218) happens on the GPU inside
the derivative kernel, not here.
"""
from __future__ import annotations

import math

_DEFAULTS = dict(
    length=7.62, diameter=0.219, nose_length=0.2, fin_span=0.2, fin_root_chord=0.20, fin_tip_chord=0.1,
    fin_count=4, fin_sweep_angle=0.0, fin_cant_angle=0.0,
    dry_mass=113.4, propellant_mass=63.5, center_of_mass_dry=5.8,
    Ixx_dry=45, Iyy_dry=971.9, Izz_dry=971.693,
)


class Rocket:
    def __init__(self, name="Sounding Rocket"):
        self.name = name
        for key, value in _DEFAULTS.items():
            setattr(self, key, value)
        self.reference_area = math.pi * (self.diameter / 2) ** 2
        self.reference_diameter = self.diameter
        self.Cd_data = {
            "mach": [0.0, 0.5, 0.8, 1.0, 1.2, 1.5, 2.0, 3.0],
            "cd0": [0.4, 0.42, 0.48, 0.65, 0.52, 0.45, 0.40, 0.38],
            "cda": [1.2, 1.25, 1.3, 1.4, 1.35, 1.25, 1.2, 1.15],
        }
        self.CP_shift_data = {
            "mach": [0.0, 0.8, 1.0, 1.2, 2.0, 3.0],
            "cp_shift": [0.0, -0.05, -0.1, -0.05, 0.0, 0.0],
        }
        self.cp_location = self._calculate_center_of_pressure()
        self.parachute_area = 15.0
        self.parachute_cd = 2.0
        self.parachute_deployment_altitude = 500
        self.power_off_drag_factor = 1.2

    def _calculate_center_of_pressure(self):
        """Barrowman estimate, evaluated once at construction (reference rocket.py:68-103)."""
        nose_cn, nose_x = 2.0, 0.666 * self.nose_length
        cr, ct, span = self.fin_root_chord, self.fin_tip_chord, self.fin_span
        taper = ct / cr if cr != 0 else 0.0
        planform = 0.5 * (cr + ct) * span
        fins_cn = 2 * self.fin_count * (1 + self.diameter / (2 * span)) * (planform / self.reference_area)
        mean_chord = (2 / 3) * cr * (1 + taper + taper ** 2) / (1 + taper)
        y_mac = span * (1 + 2 * taper) / (3 * (1 + taper))
        fins_x = (self.length - cr) + y_mac * math.tan(self.fin_sweep_angle) + 0.25 * mean_chord
        total = nose_cn + 0.0 + fins_cn
        if total > 0:
            return (nose_cn * nose_x + 0.0 * 0.0 + fins_cn * fins_x) / total
        return self.length / 2

    # model-evaluation helpers (rocket.py:105-218 of the reference): evaluated on the GPU by the engine
    def get_mass_properties(self, propellant_fraction_remaining):
        from .simulator import _scalar_or_array, evaluate_component
        o = evaluate_component(1, (propellant_fraction_remaining, self.dry_mass, self.propellant_mass), rocket=self)
        return {k: _scalar_or_array(o[i], propellant_fraction_remaining) for i, k in enumerate(("mass", "center_of_mass", "Ixx", "Iyy", "Izz"))}

    def get_aerodynamic_coefficients(self, mach, alpha, beta=0.0, mass_props=None, power_on=True):
        from .simulator import _scalar_or_array, evaluate_component
        cg = self.center_of_mass_dry if mass_props is None else mass_props["center_of_mass"]
        o = evaluate_component(2, (mach, alpha, beta, cg, 1.0 if power_on else 0.0, 1.0), rocket=self)
        v = lambda i: _scalar_or_array(o[i], mach)
        return {"cd": v(0), "cl": v(1), "cm": v(2), "cp": v(5), "cn": v(6), "cy": v(3), "croll": 0.0, "cpitch": v(2), "cyaw": v(4)}

    def get_dynamic_cp(self, mach, alpha=0.0):
        return self.get_aerodynamic_coefficients(mach, alpha)["cp"]

    def get_stability_margin(self, propellant_fraction_remaining):
        return (self.cp_location - self.get_mass_properties(propellant_fraction_remaining)["center_of_mass"]) / self.reference_diameter
