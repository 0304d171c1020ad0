"""
TEST / BENCH INFRASTRUCTURE — the CPU baseline leg that times the UNMODIFIED Python reference itself.

Runs `MonteCarloAnalyzer.run_monte_carlo` of smcconoughey/erpl_monte_carlo_sim (rocket_simulation/monte_carlo.py:52-90:
parameter draws, ProcessPoolExecutor fan-out of one FlightSimulator.simulate_flight per sample, statistics) on the host
cores of this machine, on the bench workload C3 (SolidMotor + sample_wind.csv wind + default dispersions, vertical
launch from z = 10 m; example.py:34-39,57-64), and prints ONE JSON line.

The reference is taken from baseline/_ref (the offline `pip install --target` that __graft_entry__.build() makes in the
build container; it travels to the GPU box with the repo snapshot).  It is imported in a process of its own — bench.py
starts this file with `python oracle/pyref_bench.py` BEFORE it touches CUDA — with a two-module matplotlib stub
(monte_carlo.py:6 imports pyplot at module top) and stdout silenced (4 prints per flight).

Never imported by the product package.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEEN = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=0, help="0 = 2 x cores")
    ap.add_argument("--processes", type=int, default=0, help="0 = os.cpu_count()")
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    pkg = os.path.join(a.ref, "rocket_simulation")
    if not os.path.isfile(os.path.join(pkg, "monte_carlo.py")):
        print(json.dumps({"unavailable": f"no reference install at {pkg}"}))
        return
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, pkg)                       # the reference uses flat intra-package imports (simulator.py:7)
    import numpy as np
    from environment import StandardAtmosphere, WindModel
    from monte_carlo import MonteCarloAnalyzer
    from motor import SolidMotor
    from rocket import Rocket

    os.environ["PYTHONWARNINGS"] = "ignore"       # overflow warnings of the blown-up flights, workers included
    import warnings
    warnings.simplefilter("ignore")
    cores = a.processes or os.cpu_count() or 1
    n = a.samples or 2 * cores
    wind_model = WindModel()
    with contextlib.redirect_stdout(io.StringIO()):
        class Capturing(MonteCarloAnalyzer):          # keeps the per-sample result dicts also when the analysis raises
            def _analyze_results(self, results):
                SEEN.extend(r for r in results if r is not None)
                return MonteCarloAnalyzer._analyze_results(self, results)
        Capturing.__qualname__ = "Capturing"
        globals()["Capturing"] = Capturing             # picklable by reference for the (forked) pool workers
        mc = Capturing(Rocket(), SolidMotor(), StandardAtmosphere(), wind_model)
        alt, wind = wind_model.load_wind_profile_from_csv(os.path.join(pkg, "sample_wind.csv"))
    mc.base_altitude_profile, mc.base_wind_profile = alt, wind
    ic = {"position": [0.0, 0.0, 10.0], "velocity": [0.0, 0.0, 0.0], "attitude": [0.0, -np.pi / 2 + 0.02, 0.0],
          "angular_velocity": [0.0, 0.0, 0.0]}
    seen = SEEN
    devnull = open(os.devnull, "w")
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull.fileno(), 1)                  # worker processes inherit fd 1: silence them too
    t0 = time.perf_counter()
    err = None
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            analysis = mc.run_monte_carlo(ic, n_samples=n, n_processes=cores)
    except ValueError as e:                       # "No physically reasonable simulation results ..." (monte_carlo.py:411-412)
        analysis, err = None, str(e)
    dt = time.perf_counter() - t0
    os.dup2(saved, 1)
    steps = int(sum(len(r["time"]) - 1 for r in seen)) if seen else None
    print(json.dumps({"value": n / dt, "unit": "trajectories/s", "cores": cores, "samples": n, "seconds": dt,
                      "rk4_steps": steps, "steps_per_s": (steps / dt) if steps else None,
                      "n_valid": analysis["n_samples"] if analysis else 0, "note": err,
                      "what": "unmodified Python reference, MonteCarloAnalyzer.run_monte_carlo (monte_carlo.py:52-90), C3 workload"}))


if __name__ == "__main__":
    main()
