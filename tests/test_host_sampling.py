"""CPU: the host side of the boundary.  The package's parameter classes, its host-seeded dispersion
draws and its marshalling must reproduce, bit for bit, the inputs the unmodified reference builds for
the same seeds (captured in tests/golden by oracle/make_golden.py), and its statistics must equal
MonteCarloAnalyzer._analyze_results of the reference (monte_carlo.py:400-473)."""
import os

import numpy as np
import pytest

import util
from erpl_monte_carlo_sim_b200 import (FlightSimulator, LiquidMotor, MonteCarloAnalyzer, Rocket, SolidMotor,
                                       StandardAtmosphere, WindModel, _abi, marshal)

CSV_ALT = np.array([0.0, 5000.0, 10000.0, 15000.0, 20000.0, 25000.0])
CSV_WIND = np.array([[2.0, 0, 0], [5, 1, 0], [8, 2, 0], [10, 2, 0], [12, 3, 0], [15, 3, 0]], float)   # sample_wind.csv:2-7


def _analyzer(motor_kind, csv):
    mc = MonteCarloAnalyzer(Rocket(), SolidMotor() if motor_kind == 1 else LiquidMotor(), StandardAtmosphere(), WindModel())
    if csv:
        mc.base_altitude_profile, mc.base_wind_profile = CSV_ALT, CSV_WIND
    return mc


def _base_ic(z):
    b = z["base_ic"]
    return {"position": b[0].tolist(), "velocity": b[1].tolist(), "attitude": b[2].tolist(), "angular_velocity": b[3].tolist()}


@pytest.mark.parametrize("name,csv", [("mc_liquid_default", False), ("mc_solid_csv", True), ("mc_readme_literal", False)])
def test_dispersion_inputs_bit_exact(name, csv):
    z = util.golden(name)
    md_ref = _abi.model_from_npz(z)
    mc = _analyzer(int(md_ref["motor_kind"]), csv)
    n = z["scalars"].shape[1]
    disp = mc.draw_parameters(n)
    pv = np.concatenate([disp.pos, disp.vel, disp.att, disp.omega, disp.mass_multiplier[:, None],
                         disp.thrust_multiplier[:, None], disp.wind_speed[:, None], disp.wind_direction[:, None],
                         disp.density_multiplier[:, None]], axis=1)
    np.testing.assert_array_equal(pv, z["params"])
    blk, wind, alts = mc.build_inputs(_base_ic(z), disp)
    np.testing.assert_array_equal(alts, md_ref["wind_altitudes"])
    np.testing.assert_array_equal(wind, z["wind"])
    np.testing.assert_array_equal(blk, z["scalars"])


def test_parameter_dicts_match_struct_of_arrays():
    mc = _analyzer(0, False)
    dicts = mc._generate_parameter_samples(5)
    z = util.golden("mc_liquid_default")
    assert set(dicts[0]) == {"initial_position_offset", "initial_velocity_offset", "initial_attitude_offset",
                             "initial_angular_velocity_offset", "mass_multiplier", "thrust_multiplier", "wind_speed",
                             "wind_direction", "density_multiplier", "random_seed"}
    assert dicts[3]["mass_multiplier"] == z["params"][3, 12] and dicts[3]["random_seed"] == 3
    # optimized path: one stream seeded 42, draws in the reference's order (monte_carlo.py:181-201)
    opt = mc._generate_parameter_samples_vectorized(3)
    rs = np.random.RandomState(42)
    first = rs.normal(0, [0.0, 0.0, 0.0]); vel = rs.normal(0, [0.1] * 3)
    np.testing.assert_array_equal(opt[0]["initial_velocity_offset"], vel)


def test_model_marshalling_matches_reference_objects():
    for name, solid, csv in (("mc_liquid_default", False, False), ("mc_solid_csv", True, True)):
        z = util.golden(name)
        ref = _abi.model_from_npz(z)
        mc = _analyzer(1 if solid else 0, csv)
        md = marshal.model_dict(mc.rocket, mc.motor, mc.atmosphere, mc._model_simulator(), mc._altitude_grid())
        assert set(md) == set(ref) | {"gamma"} and md["gamma"] == 1.4       # ABI 2 field; the goldens predate it (default 1.4)
        for k in ref:
            np.testing.assert_array_equal(np.asarray(md[k], float), np.asarray(ref[k], float), err_msg=k)
        _abi.pack_model(md)


def test_single_flight_block_matches_reference():
    z = util.golden("flights_single")
    ic = {"position": [0.0, 0.0, 10.0], "velocity": [0, 0, 0.0], "attitude": [0.0, -np.pi / 2 + 0.02, 0.0],
          "angular_velocity": [0.0, 0.0, 0.0]}
    blk = marshal.single_sample_block(ic, Rocket(), LiquidMotor())
    np.testing.assert_array_equal(blk, z["c1b_example_liquid_csv__scalars"])
    blk = marshal.single_sample_block({"attitude": [0.0, 0.0, 0.0]}, Rocket(), SolidMotor())
    np.testing.assert_array_equal(blk, z["testfixes_solid__scalars"])


def test_wind_generators_scalar_api_equals_batch():
    wm = WindModel()
    alts = np.linspace(0, 25000, 100)
    a = wm.generate_stochastic_profile(alts, 3.3, 1.1, np.random.RandomState(5))
    z = util.golden("mc_liquid_default")
    p = z["params"][5]
    b = wm.generate_stochastic_profile(alts, p[14], p[15], np.random.RandomState(5))
    np.testing.assert_array_equal(b, z["wind"][5])
    assert a.shape == (100, 3)
    c = wm.perturb_wind_profile(CSV_ALT, CSV_WIND, np.random.RandomState(9))
    zc = util.golden("mc_solid_csv")
    pc = zc["params"][9]
    c[:, 0] += pc[14] * np.cos(pc[15]); c[:, 1] += pc[14] * np.sin(pc[15])
    np.testing.assert_array_equal(c, zc["wind"][9])


def test_csv_loader(tmp_path):
    f = tmp_path / "w.csv"
    f.write_text("altitude,u,v,w\n0,2,0,0\n5000,5,1,0\n")
    alt, w = WindModel().load_wind_profile_from_csv(str(f))
    np.testing.assert_array_equal(alt, [0.0, 5000.0]); np.testing.assert_array_equal(w, [[2, 0, 0], [5, 1, 0]])
    f.write_text("altitude,u,v\n0,2,0\n5000,5,1\n")
    alt, w = WindModel().load_wind_profile_from_csv(str(f))
    assert w.shape == (2, 3) and np.all(w[:, 2] == 0)


def test_constructor_surface():
    r = Rocket("x")
    assert r.cp_location == util.golden("components")["cp_location"].item()
    assert r.reference_area == np.pi * (0.219 / 2) ** 2
    lm = LiquidMotor()
    assert lm.burn_time == 63.5 / 4.26 and lm.nozzle_exit_area == (2590 * 4.44822 - 2290 * 4.44822) / 101325.0
    sm = SolidMotor()
    assert sm.average_thrust == 156297 / 15.0 and sm.thrust_curve_thrust[1] == 2.2 * sm.average_thrust
    p = sm.perturb_for_monte_carlo(np.random.RandomState(3))
    k = np.random.RandomState(3).normal(1.0, 0.05)
    assert p.mass_flow_rate == 4.26 * k and p.nozzle_exit_area == sm.nozzle_exit_area * k
    sim = FlightSimulator(r, lm, StandardAtmosphere(), WindModel())
    assert (sim.max_time, sim.dt_initial, sim.pitch_damping, sim.yaw_damping) == (300.0, 0.01, 20.0, 20.0)
    mc = MonteCarloAnalyzer(r, lm, StandardAtmosphere(), WindModel())
    assert mc.uncertainty_params["wind_speed_range"] == [0.0, 5.0] and mc.n_cores == os.cpu_count()


def test_analysis_matches_reference():
    z = util.golden("analysis")
    mc = _analyzer(0, False)
    results = []
    for k in range(len(z["in_apogee"])):
        if z["in_failed"][k]:
            results.append(None)
            continue
        results.append({"apogee_altitude": z["in_apogee"][k], "range": z["in_range"][k], "flight_time": z["in_flight_time"][k],
                        "simulation_id": k if k < 3 else k - 1, "parameters": {"mass_multiplier": 1.0}})
    an = mc._analyze_results(results)
    assert (an["n_samples"], an["n_failed"], an["n_outliers"]) == (int(z["n_samples"]), int(z["n_failed"]), int(z["n_outliers"]))
    assert [r["simulation_id"] for r in an["results"]] == z["valid_ids"].tolist()
    assert [r["simulation_id"] for r in an["outliers"]] == z["outlier_ids"].tolist()
    for key in ("apogee_altitude", "range", "flight_time"):
        s = an[key]
        got = np.array([s["mean"], s["std"], s["min"], s["max"], *s["percentiles"]])
        np.testing.assert_allclose(got, z[key], rtol=1e-14)
    ok = ~z["in_failed"]
    mask = mc.outlier_mask(z["in_apogee"][ok], z["in_range"][ok], z["in_flight_time"][ok])
    assert mask.sum() == int(z["n_outliers"])
    with pytest.raises(ValueError, match="No physically reasonable"):
        mc._analyze_results([{"apogee_altitude": 0.0, "range": 1.0, "flight_time": 1.0}])
    with pytest.raises(ValueError, match="No valid simulation results"):
        mc._analyze_results([None])
