"""
TEST INFRASTRUCTURE — golden-vector generator.  Run in the build container only:

    python oracle/make_golden.py [--jobs 8] [--only NAME]

Imports the UNMODIFIED Python reference from /root/reference (oracle/ref_harness.py), runs it on the
configurations of SURVEY.md §8d and writes the inputs (in the C-ABI layout of include/emc.h) together
with the reference's outputs to tests/golden/*.npz.  The fixtures travel to the GPU box; the
reference does not.  Numbers only — no reference code is copied.

Files written
  components.npz         atmosphere / gravity / mass properties / aero coefficients / thrust / interp / quaternion KATs
  derivative_<cfg>.npz   _rocket_dynamics on random states in every branch (3 configurations)
  flights_single.npz     C1a, test_fixes nominal, C1b (example.py), C1c Liquid, C1c Solid  (+ decimated series)
  mc_liquid_default.npz  C2-vertical: Liquid, 100-pt synthetic wind, default dispersions, seeds 0..63
  mc_solid_csv.npz       C3: Solid + sample_wind.csv + perturbations, seeds 0..63
  mc_planar_liquid.npz / mc_planar_solid.npz   W-B launch->landing set (beta == 0), seeds 0..7
  mc_readme_literal.npz  C2 literal (pitch 0.02): one-step flights, seeds 0..15
  mc_solid_csv_blowup.npz  C3 seeds from 300000 on whose blow-up reaches z = -inf inside an RK4 stage (round-2 regression set)
  analysis.npz           MonteCarloAnalyzer._analyze_results on a mixed valid/outlier result list
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
VERTICAL = [0.0, -np.pi / 2 + 0.02, 0.0]

SERIES_KEYS = ["time", "propellant_fraction", "mass", "altitude", "speed", "center_of_mass", "thrust",
               "drag", "cd", "cl", "cm", "cp_location_dynamic", "stability_margin", "angle_of_attack",
               "sideslip_angle"]
SERIES_KEYS_2D = ["position", "velocity", "quaternion", "angular_velocity", "moments_of_inertia",
                  "euler_angles"]


def flat_model(md):
    return {"model__" + k: np.asarray(v) for k, v in md.items()}


def make_objects(motor_kind, csv_wind):
    R = H.ref()
    rocket = R.rocket.Rocket("Sounding Rocket")
    motor = R.motor.SolidMotor() if motor_kind == "solid" else R.motor.LiquidMotor("Liquid Motor")
    atm = R.environment.StandardAtmosphere()
    wm = R.environment.WindModel()
    alt = wind = None
    if csv_wind:
        alt, wind = wm.load_wind_profile_from_csv(os.path.join(H.REF_PKG, "sample_wind.csv"))
    return rocket, motor, atm, wm, alt, wind


# ------------------------------------------------------------------------------------------
# MC jobs (run in worker processes)
# ------------------------------------------------------------------------------------------
def _mc_job(job):
    cfg, seed = job
    R = H.ref()
    Probe = H.make_probe_simulator_class()
    R.monte_carlo.FlightSimulator = Probe            # observation-only subclass, see ref_harness
    rocket, motor, atm, wm, alt, wind = make_objects(cfg["motor"], cfg["csv"])
    with H.quiet():
        mc = R.monte_carlo.MonteCarloAnalyzer(rocket, motor, atm, wm)
        if cfg["csv"]:
            mc.base_altitude_profile = alt
            mc.base_wind_profile = wind
        params = mc._generate_parameter_samples(seed + 1)[seed]
        base_ic = {k: list(v) for k, v in cfg["ic"].items()}
        ic, prk, pmo, patm, wprof, aprof = H.mc_sample_setup(mc, base_ic, params, planar=cfg["planar"])
        t0 = time.time()
        if cfg["planar"]:
            sim = Probe(prk, pmo, patm, wm)
            res = sim.simulate_flight(ic, wprof, aprof)
        else:
            res = mc._run_single_simulation(base_ic, params, seed)    # the reference's own per-sample path
            sim = Probe(prk, pmo, patm, wm)
        wall = time.time() - t0
        out, iout = H.summarize(res, sim, wprof, aprof)
    col = H.sample_scalars(ic, prk, pmo)
    pvec = np.concatenate([np.ravel(params[k]) for k in
                           ("initial_position_offset", "initial_velocity_offset", "initial_attitude_offset",
                            "initial_angular_velocity_offset", "mass_multiplier", "thrust_multiplier",
                            "wind_speed", "wind_direction", "density_multiplier")]).astype(float)
    slim = {k: res[k] for k in ("apogee_altitude", "range", "flight_time")}
    return seed, col, np.asarray(wprof, float), out, iout, pvec, wall, slim


def run_mc(name, cfg, seeds, pool):
    R = H.ref()
    t0 = time.time()
    res = pool.map(_mc_job, [(cfg, s) for s in seeds], chunksize=1)
    res.sort(key=lambda r: r[0])
    rocket, motor, atm, wm, alt, wind = make_objects(cfg["motor"], cfg["csv"])
    with H.quiet():
        sim = R.simulator.FlightSimulator(rocket, motor, atm, wm)
    aprof = alt if cfg["csv"] else np.linspace(0, 25000, 100)
    md = H.model_dict(rocket, motor, atm, sim, aprof)
    data = dict(
        seeds=np.array([r[0] for r in res], np.int64),
        scalars=np.stack([r[1] for r in res], axis=1),
        wind=np.stack([r[2] for r in res], axis=0),
        out=np.stack([r[3] for r in res], axis=1),
        iout=np.stack([r[4] for r in res], axis=1),
        params=np.stack([r[5] for r in res], axis=0),
        ref_wall_s=np.array([r[6] for r in res]),
        base_ic=np.array([cfg["ic"]["position"], cfg["ic"]["velocity"], cfg["ic"]["attitude"],
                          cfg["ic"]["angular_velocity"]], float),
        planar=np.array(int(cfg["planar"])),
        in_fields=np.array(H.IN_FIELDS), out_fields=np.array(H.OUT_FIELDS), iout_fields=np.array(H.IOUT_FIELDS),
        **flat_model(md))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **data)
    steps = data["iout"][0]
    print(f"[{name}] n={len(seeds)} steps min/mean/max {steps.min()}/{steps.mean():.0f}/{steps.max()} "
          f"ref CPU {data['ref_wall_s'].sum():.0f}s wall {time.time() - t0:.0f}s", flush=True)
    return [r[7] for r in res], data


# ------------------------------------------------------------------------------------------
def gen_components():
    R = H.ref()
    rng = np.random.RandomState(1234)
    rocket, motor_l, atm, wm, _, _ = make_objects("liquid", False)
    motor_s = R.motor.SolidMotor()
    z = np.concatenate([np.array([-500.0, 0.0, 5000.0, 11000.0, 11000.0001, 15000.0, 20000.0, 20000.0001, 25000.0,
                                  25000.0001, 30000.0, 32000.0, 32000.0001, 40000.0, 50000.0, 49380.0, 99999.0, 2e5]),
                        rng.uniform(-1000, 120000, 200)])
    atm_out = np.array([[a["temperature"], a["pressure"], a["density"]] for a in map(atm.get_properties, z)])
    grav = np.array([atm.get_gravity(v) for v in z])
    pf = np.concatenate([[0.0, 1.0, 0.5, 1e-9], rng.uniform(0, 1, 60)])
    dm = rng.normal(1.0, 0.02, pf.size)
    mp_out = []
    for f, k in zip(pf, dm):
        rocket.dry_mass, rocket.propellant_mass = 113.4 * k, 63.5 * k
        d = rocket.get_mass_properties(f)
        mp_out.append([d["mass"], d["center_of_mass"], d["Ixx"], d["Iyy"]])
    rocket.dry_mass, rocket.propellant_mass = 113.4, 63.5
    n = 400
    mach = np.concatenate([[0.0, 0.5, 0.8, 1.0, 1.2, 1.5, 2.0, 3.0, 3.5, 0.9, 1.1, 2.5], rng.uniform(0, 7, n)])
    alpha = np.concatenate([[0, 0, 0, 0, 0, 0, 0, 0, 0, .05, .3, -.6], rng.normal(0, 0.4, n)])
    alpha[20:40] = np.radians(15.0) * rng.choice([-1, 1], 20) * (1 + rng.uniform(-1e-3, 1e-3, 20))
    alpha[40:50] = rng.uniform(-3.2, 3.2, 10)
    beta = np.concatenate([[0] * 9, [-.02, .1, .2], rng.normal(0, 0.2, n)])
    cg = rng.uniform(5.5, 5.8, mach.size)
    pon = rng.randint(0, 2, mach.size)
    aero = []
    for M, a, b, c, p in zip(mach, alpha, beta, cg, pon):
        d = rocket.get_aerodynamic_coefficients(M, a, b, {"center_of_mass": c}, power_on=bool(p))
        aero.append([d["cd"], d["cl"], d["cm"], d["cy"], d["cyaw"], d["cp"]])
    tt = np.concatenate([[-1.0, 0.0, 0.1, 0.2, 1.0, 14.0, 14.9, 15.0, 15.0001, 20.0], rng.uniform(0, 16, 50)])
    pp = rng.uniform(0, 101325, tt.size)
    thr_l = np.array([motor_l.get_thrust(t, p) for t, p in zip(tt, pp)])
    thr_s = np.array([motor_s.get_thrust(t, p) for t, p in zip(tt, pp)])
    eul = rng.uniform(-3.1, 3.1, (64, 3))
    eul[0] = VERTICAL
    eul[1] = [0, 0.02, 0]
    quat = np.array([R.utils.euler_to_quaternion(*e) for e in eul])
    qraw = quat * (1 + rng.uniform(-1e-3, 1e-3, (64, 1)))
    eul_back = np.array([R.utils.quaternion_to_euler(q) for q in qraw])
    rot = np.array([R.utils.quaternion_to_rotation_matrix(q) for q in qraw])
    xs = np.concatenate([[-1.0, 0.0, 0.5, 0.8, 3.0, 3.0001, np.inf, -np.inf, np.nan], rng.uniform(-0.5, 3.5, 100)])
    itp = np.array([R.utils.interpolate_1d(x, rocket.Cd_data["mach"], rocket.Cd_data["cd0"]) for x in xs])
    with H.quiet():
        sim = R.simulator.FlightSimulator(rocket, motor_l, atm, wm)
        sim_s = R.simulator.FlightSimulator(rocket, motor_s, atm, wm)
    np.savez_compressed(
        os.path.join(GOLDEN, "components.npz"),
        atm_z=z, atm_out=atm_out, gravity=grav, mp_pf=pf, mp_mult=dm, mp_out=np.array(mp_out),
        aero_mach=mach, aero_alpha=alpha, aero_beta=beta, aero_cg=cg, aero_power_on=pon, aero_out=np.array(aero),
        thr_t=tt, thr_p=pp, thr_liquid=thr_l, thr_solid=thr_s,
        liquid_scalars=np.array([motor_l.thrust_vacuum, motor_l.nozzle_exit_area, motor_l.mass_flow_rate, motor_l.burn_time]),
        solid_scalars=np.array([1.0, motor_s.nozzle_exit_area, motor_s.mass_flow_rate, motor_s.burn_time]),
        euler=eul, quat=quat, quat_raw=qraw, euler_back=eul_back, rot=rot, interp_x=xs, interp_out=itp,
        cp_location=np.array(rocket.cp_location),
        **{("liquid_" + k): v for k, v in flat_model(H.model_dict(rocket, motor_l, atm, sim)).items()},
        **{("solid_" + k): v for k, v in flat_model(H.model_dict(rocket, motor_s, atm, sim_s)).items()})
    print("[components] done", flush=True)


def gen_derivative(name, motor_kind, csv, with_wind, n=320):
    R = H.ref()
    rng = np.random.RandomState(99 if motor_kind == "liquid" else 77)
    rocket, motor, atm, wm, alt, wind = make_objects(motor_kind, csv)
    with H.quiet():
        mc = R.monte_carlo.MonteCarloAnalyzer(rocket, motor, atm, wm)
        if csv:
            mc.base_altitude_profile, mc.base_wind_profile = alt, wind
        params = mc._generate_parameter_samples(n)
    base_ic = dict(position=[0, 0, 0.0], velocity=[0, 0, 0.0], attitude=VERTICAL, angular_velocity=[0, 0, 0.0])
    cols, winds, states, ts, chin, chout, sdot = [], [], [], [], [], [], []
    aprof = None
    for i in range(n):
        with H.quiet():
            ic, prk, pmo, patm, wprof, aprof = H.mc_sample_setup(mc, base_ic, params[i])
            sim = R.simulator.FlightSimulator(prk, pmo, patm, wm)
        if with_wind:
            sim.wind_profile, sim.altitude_profile = wprof, aprof
        st = np.zeros(14)
        mode = i % 16
        st[0:2] = rng.normal(0, 3000, 2)
        st[2] = rng.choice([rng.uniform(-200, 1200), rng.uniform(0, 11000), rng.uniform(11000, 20000),
                            rng.uniform(20000, 25000), rng.uniform(25000, 32000), rng.uniform(32000, 60000),
                            rng.uniform(60000, 120000)])
        speed = 10 ** rng.uniform(-1, 3.4)
        d = rng.normal(0, 1, 3); d /= np.linalg.norm(d)
        st[3:6] = d * speed
        e = np.array(VERTICAL) + rng.normal(0, 0.5, 3)
        st[6:10] = R.utils.euler_to_quaternion(*e) * (1 + rng.uniform(-1e-3, 1e-3))
        st[10:13] = rng.normal(0, 0.5, 3)
        st[13] = rng.choice([1.0, rng.uniform(0, 1), rng.uniform(0, 1e-3), 0.0, -1e-3])
        t = rng.uniform(0, 25)
        ch = 0
        if mode == 1:      # aligned with the body axis: small alpha/beta
            Rm = R.utils.quaternion_to_rotation_matrix(st[6:10])
            st[3:6] = Rm[:, 0] * speed + rng.normal(0, 1e-3 * speed, 3)
        elif mode == 2:    # parachute already out
            ch = 1
        elif mode == 3:    # chute trigger: low and descending
            st[2] = rng.uniform(-50, 500); st[5] = -abs(st[5]) - 0.1
        elif mode == 4:    # zero relative velocity only possible without wind; otherwise tiny
            st[3:6] = 0.0
        elif mode == 5:    # alpha dead zone: body x/z components below 1e-6
            Rm = R.utils.quaternion_to_rotation_matrix(st[6:10])
            w = sim.wind_model.get_wind_at_altitude(st[2], wprof, aprof) if with_wind else np.zeros(3)
            st[3:6] = w + Rm[:, 1] * rng.uniform(0.5, 5.0) + Rm[:, 0] * 1e-7
        elif mode == 6:    # burnout taper window
            t = float(pmo.burn_time) - rng.uniform(0, 0.02); st[13] = rng.uniform(0, 5e-3)
        elif mode == 7:    # just past burn time
            t = float(pmo.burn_time) + rng.uniform(0, 0.01); st[13] = rng.uniform(0, 0.1)
        elif mode == 8:    # heavy stall
            Rm = R.utils.quaternion_to_rotation_matrix(st[6:10])
            a = rng.uniform(0.2, 3.0) * rng.choice([-1, 1])
            st[3:6] = (Rm[:, 0] * np.cos(a) + Rm[:, 2] * np.sin(a)) * speed
        sim.parachute_deployed = bool(ch)
        with H.quiet():
            out = sim._rocket_dynamics(t, st.copy())
        cols.append(H.sample_scalars(ic, prk, pmo)); winds.append(np.asarray(wprof, float))
        states.append(st); ts.append(t); chin.append(ch); chout.append(int(sim.parachute_deployed)); sdot.append(out)
    with H.quiet():
        sim0 = R.simulator.FlightSimulator(rocket, motor, atm, wm)
    md = H.model_dict(rocket, motor, atm, sim0, aprof if with_wind else None)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), scalars=np.stack(cols, 1),
                        wind=np.stack(winds, 0) if with_wind else np.zeros((0, 0, 3)),
                        t=np.array(ts), state=np.array(states), chute_in=np.array(chin, np.int32),
                        chute_out=np.array(chout, np.int32), state_dot=np.array(sdot), **flat_model(md))
    print(f"[{name}] done", flush=True)


def _single_job(job):
    name, motor_kind, csv, ic = job
    R = H.ref()
    Probe = H.make_probe_simulator_class()
    rocket, motor, atm, wm, alt, wind = make_objects(motor_kind, csv)
    with H.quiet():
        sim = Probe(rocket, motor, atm, wm)
        t0 = time.time()
        res = sim.simulate_flight({k: list(v) for k, v in ic.items()}, wind, alt) if csv else \
            sim.simulate_flight({k: list(v) for k, v in ic.items()})
        wall = time.time() - t0
        out, iout = H.summarize(res, sim, wind, alt)
    col = H.sample_scalars(ic, rocket, motor)
    n = res["time"].size
    idx = np.unique(np.concatenate([np.arange(0, n, 37), np.arange(max(0, n - 5), n), np.arange(0, min(n, 5))]))
    series = {"series__" + k: np.asarray(res[k])[idx] for k in SERIES_KEYS}
    series.update({"series__" + k: np.asarray(res[k])[:, idx] for k in SERIES_KEYS_2D})
    tape = np.concatenate([(res["time"] + res["rail_exit_time"])[None, :], res["position"], res["velocity"],
                           res["quaternion"], res["angular_velocity"], res["propellant_fraction"][None, :]], 0).T
    series["series__tape"] = tape[idx]
    series["series__idx"] = idx
    md = H.model_dict(rocket, motor, atm, sim, alt)
    return name, col, (np.asarray(wind, float) if csv else None), out, iout, series, md, wall


def gen_singles(pool):
    z3 = [0.0, 0.0, 0.0]
    jobs = [
        ("c1a_readme_liquid", "liquid", False, dict(position=z3, velocity=z3, attitude=[0.0, 0.02, 0.0], angular_velocity=z3)),
        ("testfixes_solid", "solid", False, dict(position=z3, velocity=z3, attitude=z3, angular_velocity=z3)),
        ("c1b_example_liquid_csv", "liquid", True, dict(position=[0.0, 0.0, 10.0], velocity=z3, attitude=VERTICAL, angular_velocity=z3)),
        ("c1c_planar_liquid", "liquid", False, dict(position=z3, velocity=z3, attitude=VERTICAL, angular_velocity=z3)),
        ("c1c_planar_solid", "solid", False, dict(position=z3, velocity=z3, attitude=VERTICAL, angular_velocity=z3)),
    ]
    res = pool.map(_single_job, jobs, chunksize=1)
    data = {"names": np.array([r[0] for r in res])}
    for name, col, wind, out, iout, series, md, wall in res:
        data[name + "__scalars"] = col[:, None]
        if wind is not None:
            data[name + "__wind"] = wind[None]
        data[name + "__out"] = out[:, None]
        data[name + "__iout"] = iout[:, None]
        data[name + "__ref_wall_s"] = np.array(wall)
        for k, v in series.items():
            data[name + "__" + k] = v
        for k, v in flat_model(md).items():
            data[name + "__" + k] = v
        print(f"[single:{name}] states={iout[0] + 1} apogee={out[16]!r} range={out[18]!r} ft={out[19]!r} wall={wall:.1f}s", flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "flights_single.npz"), **data)


def gen_analysis(valid_slim, outlier_slim):
    """_analyze_results (monte_carlo.py:400-473) on slim result dicts (only the keys it reads)."""
    R = H.ref()
    rocket, motor, atm, wm, _, _ = make_objects("liquid", False)
    with H.quiet():
        mc = R.monte_carlo.MonteCarloAnalyzer(rocket, motor, atm, wm)
        results = []
        for k, r in enumerate(list(valid_slim) + list(outlier_slim)):
            d = dict(r); d["simulation_id"] = k; d["parameters"] = {"mass_multiplier": 1.0 + 0.01 * k}
            results.append(d)
        results.insert(3, None)          # one failed simulation
        an = mc._analyze_results(results)
    data = dict(in_apogee=np.array([np.nan if r is None else r["apogee_altitude"] for r in results]),
                in_range=np.array([np.nan if r is None else r["range"] for r in results]),
                in_flight_time=np.array([np.nan if r is None else r["flight_time"] for r in results]),
                in_failed=np.array([r is None for r in results]),
                n_samples=np.array(an["n_samples"]), n_failed=np.array(an["n_failed"]), n_outliers=np.array(an["n_outliers"]),
                valid_ids=np.array([r["simulation_id"] for r in an["results"]]),
                outlier_ids=np.array([r["simulation_id"] for r in an["outliers"]]))
    for key in ("apogee_altitude", "range", "flight_time"):
        s = an[key]
        data[key] = np.array([s["mean"], s["std"], s["min"], s["max"], *s["percentiles"]])
    np.savez_compressed(os.path.join(GOLDEN, "analysis.npz"), **data)
    print(f"[analysis] n_samples={an['n_samples']} n_outliers={an['n_outliers']} n_failed={an['n_failed']}", flush=True)


# seeds of the round-2 blow-up regression set (see main)
BLOWUP_SEEDS = [300565, 302472, 305702, 305805, 307702, 309836, 310157, 310290, 310955, 312340, 313147, 314634, 314852, 315939, 316406, 317392, 317574, 317885, 318351, 318468, 319094, 320154, 320637, 320954, 320985, 321037, 321503, 321505, 324352, 326049, 326298, 328898, 329273, 330702, 333907, 334880, 335859, 336121, 336832, 338056, 338878, 339516, 340040, 340186, 342310, 342370, 344207, 346030, 346672, 346675]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    H.ref()
    want = (lambda n: (not a.only) or a.only in n)
    z3 = [0.0, 0.0, 0.0]
    ic_vert = dict(position=z3, velocity=z3, attitude=VERTICAL, angular_velocity=z3)
    ic_vert10 = dict(position=[0.0, 0.0, 10.0], velocity=z3, attitude=VERTICAL, angular_velocity=z3)
    ic_readme = dict(position=z3, velocity=z3, attitude=[0.0, 0.02, 0.0], angular_velocity=z3)
    if want("components"):
        gen_components()
    if want("derivative"):
        gen_derivative("derivative_liquid_wind100", "liquid", False, True)
        gen_derivative("derivative_solid_csv", "solid", True, True)
        gen_derivative("derivative_liquid_nowind", "liquid", False, False)
    with mp.get_context("fork").Pool(a.jobs) as pool:
        if want("single"):
            gen_singles(pool)
        planar = wa = None
        if want("mc_planar") or want("analysis"):
            planar, _ = run_mc("mc_planar_liquid", dict(motor="liquid", csv=False, ic=ic_vert, planar=True), range(8), pool)
            run_mc("mc_planar_solid", dict(motor="solid", csv=False, ic=ic_vert, planar=True), range(8), pool)
        if want("mc_liquid_default") or want("analysis"):
            wa, _ = run_mc("mc_liquid_default", dict(motor="liquid", csv=False, ic=ic_vert, planar=False), range(64), pool)
        if want("mc_solid_csv"):
            run_mc("mc_solid_csv", dict(motor="solid", csv=True, ic=ic_vert10, planar=False), range(64), pool)
        if want("mc_blowup"):
            # round 2: the 50 samples of a 50 000-sample C3 sweep (seeds 300000..) on which the first strict continuation still
            # differed from the C oracle — all of them blow-ups that reach z = -inf in an RK4 stage (np.interp returns the end
            # value of the wind table there, not NaN) and then grind to max_time as NaN runs
            run_mc("mc_solid_csv_blowup", dict(motor="solid", csv=True, ic=ic_vert10, planar=False), BLOWUP_SEEDS, pool)
        if want("mc_readme"):
            run_mc("mc_readme_literal", dict(motor="liquid", csv=False, ic=ic_readme, planar=False), range(16), pool)
        if want("analysis") and planar is not None and wa is not None:
            gen_analysis(planar, wa[:12])


if __name__ == "__main__":
    main()
