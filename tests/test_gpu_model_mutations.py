"""GPU (-m gpu): users of the reference mutate attributes between calls (tables, atmosphere constants, integrator knobs,
wind grids).  Randomised models — different knot counts, non-uniform wind grids up to the engine limit, other time steps,
recovery and damping settings — flown by the CUDA engine and by the C oracle on the same inputs."""
import numpy as np
import pytest

import oracle_lib as O
import util
from erpl_monte_carlo_sim_b200 import _abi

pytestmark = pytest.mark.gpu


def _mutated_model(base, rng, case):
    md = dict(base)
    n_cd = int(rng.choice([2, 3, 8, 16])); n_cp = int(rng.choice([2, 6, 16]))
    md["cd_mach"] = np.concatenate([[0.0], np.sort(rng.uniform(0.2, 4.0, n_cd - 1))])
    md["cd0"] = rng.uniform(0.3, 0.7, n_cd); md["cda"] = rng.uniform(1.0, 1.5, n_cd)
    md["cp_mach"] = np.concatenate([[0.0], np.sort(rng.uniform(0.3, 3.5, n_cp - 1))])
    md["cp_shift"] = rng.uniform(-0.1, 0.05, n_cp)
    if int(md["motor_kind"]) == 1:
        n_t = int(rng.choice([2, 10, 32]))
        md["thrust_time"] = np.concatenate([[0.0], np.sort(rng.uniform(0.1, 14.9, n_t - 2)), [15.0]])
        md["thrust_curve"] = np.concatenate([[0.0], rng.uniform(5000, 20000, n_t - 2), [0.0]])
    md["power_off_drag_factor"] = rng.uniform(1.0, 1.5)
    md["parachute_deployment_altitude"] = rng.choice([300.0, 500.0, 1500.0])
    md["parachute_area"] = rng.uniform(5, 20); md["parachute_cd"] = rng.uniform(1.0, 2.2)
    md["pitch_damping"] = rng.uniform(5, 40); md["yaw_damping"] = rng.uniform(5, 40)
    md["dt_initial"] = float(rng.choice([0.01, 0.004, 0.0025, 0.02]))
    md["max_time"] = float(rng.choice([300.0, 40.0, 12.345]))
    md["rail_length"] = float(rng.choice([18.288, 6.0, 0.0]))
    md["fin_sweep_angle"] = rng.uniform(0, 0.5); md["fin_span"] = rng.uniform(0.15, 0.3)
    md["sea_level_temperature"] = rng.uniform(280, 295); md["temperature_lapse_rate"] = rng.uniform(0.006, 0.007)
    md["troposphere_height"] = rng.choice([11000.0, 10500.0]); md["stratosphere_temp"] = md["sea_level_temperature"] - md["temperature_lapse_rate"] * md["troposphere_height"]
    return md


@pytest.mark.parametrize("golden_name,case", [("mc_planar_solid", 0), ("mc_planar_liquid", 1), ("mc_solid_csv", 2), ("mc_liquid_default", 3)])
def test_gpu_matches_oracle_on_mutated_models(engine, golden_name, case):
    z = util.golden(golden_name)
    rng = np.random.RandomState(100 + case)
    base = _abi.model_from_npz(z)
    n = 48
    pick = rng.randint(0, z["scalars"].shape[1], n)
    sc = np.ascontiguousarray(z["scalars"][:, pick])
    for trial in range(3):
        md = _mutated_model(base, rng, case)
        # wind grid: non-uniform, 2 .. 1024 knots, per-sample tables
        n_w = int(rng.choice([2, 7, 100, 1024]))
        alts = np.concatenate([[0.0], np.sort(rng.uniform(10.0, 60000.0, n_w - 1))])
        md["wind_altitudes"] = alts; md["has_wind"] = 1
        wind = rng.normal(0, 6.0, (n, n_w, 3))
        if golden_name.startswith("mc_planar"):
            wind[:, :, 1] = 0.0
        wind = np.ascontiguousarray(wind)
        engine.set_model(md)
        out, iout = engine.run_batch(sc, wind)
        ref, iref = O.batch(md, sc, wind)
        np.testing.assert_array_equal(iout, iref, err_msg=f"{golden_name} trial {trial}: integer outputs")
        sens = util.oracle_sensitivity(md, sc, wind)
        util.assert_summary_close(out, ref, what=f"{golden_name} trial {trial}", sens=sens)


def test_gpu_no_wind_and_shared_table_models(engine):
    z = util.golden("mc_planar_solid")
    md = dict(_abi.model_from_npz(z)); md["has_wind"] = 0; md["wind_altitudes"] = np.zeros(0)
    engine.set_model(md)
    out, iout = engine.run_batch(z["scalars"])
    ref, iref = O.batch(md, z["scalars"], None)
    np.testing.assert_array_equal(iout, iref)
    util.assert_summary_close(out, ref, what="no wind", sens=util.oracle_sensitivity(md, z["scalars"], None))
